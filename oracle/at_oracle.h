/* oracle/at_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's per-frame localization algorithm, generalised to
 * (M microphones, N = 2^n_bits samples, +-L lags).  At (3, 1024, 46) it must reproduce the
 * reference's own objects (oracle/_ref/libat_ref.so) bit for bit; tests/test_oracle_vs_ref.py
 * pins that, and tests/golden/ holds vectors generated from those objects.  For every other
 * shape (8 mics, 4096 samples, 48 kHz ...) the reference has no implementation: parity is
 * "unpinned" there and this file is the definition.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libat_b200.so) never does.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- frame preparation (rolling_buffer.c:43-71, buffer.c:4-18) ---- */
int64_t ato_dc_remove(const int16_t *ring, int n_bits, int head, int16_t *out);
void ato_shift8(int16_t *x, int n);
void ato_window(int16_t *x, int n_bits, const int16_t *table, int table_bits);

/* ---- correlation, peak, re-weighting, temporal average (correlations.c) ---- */
void ato_xcorr(const int16_t *a, const int16_t *b, int n, int L, int64_t *c, int32_t *best);
void ato_gauss(int64_t *c, int L, int best);
void ato_average(int64_t *est, int32_t *est_best, uint64_t *est_time,
                 const int64_t *fresh, int L, uint64_t now_us);

/* ---- ring + onset gate (rolling_buffer.c:3-41, :73-85; sample_compute.h:21, :75-91) ---- */
typedef struct {
    int32_t head;
    int32_t full;
    int32_t n_bits;
    int64_t in_pow, in_tot, out_pow, out_tot;
    int16_t *buf;
} ato_ring;
void ato_ring_init(ato_ring *r, int16_t *storage, int n_bits);
void ato_ring_push(ato_ring *r, int16_t s);
int64_t ato_ring_incoming(const ato_ring *r);
int64_t ato_ring_outgoing(const ato_ring *r);
/* stream[n][M] bytes; returns ticks consumed when the gate fires, -1 if never. heads[M] out. */
long ato_capture(const uint8_t *stream, size_t n, int n_mics, int n_bits,
                 int16_t *ring_storage /*[M][N]*/, int32_t *head_out);

/* ---- geometry and lag LUT (microphones.c:9-61; vga_heatmap.h:11-13, :50-92) ---- */
void ato_mics_triangle(float d_ab, float d_bc, float d_ca, int mirror, int rotate, float *xy /*[3][2]*/);
void ato_lut_build(const float *mic_xy, int n_mics, int L, float rate_hz, float speed,
                   int half_w, int half_h, float px_per_m, float height,
                   uint8_t *idx /*[pairs][2*half_h+1][2*half_w+1]*/);

/* the same table for arbitrary 3-D candidate positions points[n][3] (no reference counterpart: the general form of
 * vga_heatmap.h:63-90 with the candidate given instead of derived from a pixel) */
void ato_lut_build_points(const float *mic_xy, int n_mics, int L, float rate_hz, float speed,
                          const float *points, int n_points, uint8_t *idx /*[pairs][n_points]*/);

/* ---- likelihood map (vga_heatmap.h:96-126).  classes may be NULL. ---- */
void ato_heatmap(const int64_t *corr /*[pairs][2L+1]*/, const uint8_t *idx, int n_pairs,
                 int n_cells, int L, int64_t *highest, int32_t *first_cell, uint8_t *classes);

/* ---- whole path for a batch of chronological uint8 frames adc[F][M][N] ---- */
typedef struct {
    int32_t n_mics, n_bits, max_shift;
    const int16_t *window;   /* Q15 table */
    int32_t window_bits;     /* log2(len(window)) >= n_bits */
    const uint8_t *lut;      /* [pairs][cells] or NULL */
    int32_t n_cells;
} ato_config;
void ato_localize(const ato_config *cfg, const uint8_t *adc, const int32_t *heads /*NULL => 0*/,
                  size_t n_frames,
                  int32_t *lags /*[F][P]*/, int64_t *corr /*[F][P][2L+1] post-Gaussian, or NULL*/,
                  int64_t *raw /*[F][P][2L+1] pre-Gaussian, or NULL*/,
                  int32_t *cell /*[F] or NULL*/, int64_t *highest /*[F] or NULL*/, int nthreads);

#ifdef __cplusplus
}
#endif
