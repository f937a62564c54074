// at_internal.h -- declarations shared by the CUDA translation units of libat_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#define AT_MAX_MICS_I 8
#define AT_MAX_PAIRS (AT_MAX_MICS_I * (AT_MAX_MICS_I - 1) / 2)

// Everything one launch of the fused localization kernel needs.  Device pointers.
struct AtFusedParams {
    // inputs
    const uint8_t *adc;      // [F][M][N] ring order (or NULL when sig16 is used)
    const int16_t *sig16;    // [F][M][N] already prepared frames (drop-in correlations_init path)
    const int32_t *heads;    // [F] or NULL
    unsigned long long n_frames;
    // tcgen05 kernel of the reference shape only: process the frames named by frame_list[0 .. *list_count) instead of
    // 0 .. n_frames-1 (the count is read on the device), and / or append the frames it could not settle to redo_list
    const uint32_t *frame_list; const uint32_t *list_count;
    uint32_t *redo_list; uint32_t *redo_count;
    // outputs (any may be NULL)
    int32_t *lags;
    void *corr; int32_t corr_struct;
    long long *raw;
    int32_t *cell; long long *highest; float *xy;
    uint8_t *gate; uint8_t *classes;
    int16_t *windowed; long long *power;
    // tables
    const int16_t *window;   // [N] Q15, already decimated to the frame length
    const float *gauss;      // [2L+1]: exp(-d*d/36) for d = 0..2L, computed with the host libm
    const uint8_t *lut;      // [P][cells]
    const uint8_t *cand_idx; // [P][n_cand] distinct lag-index tuples of the LUT
    const int32_t *cand_cell;// [n_cand] first row-major cell of each tuple, ascending
    const int4 *cand_cxy;    // [n_cand] {that cell, its x and y (float bits), 0}: cell and coordinates of a tuple in one load
    const uint4 *cand_row32; // [n_cand][2] the same tuples, one 32-byte row per candidate (pairs beyond P zero): arrays with many pairs
    // the same tuples sorted by (index of pair 0, index of pair 1) for the bounded search:
    const uint8_t *cs_idx;   // [P][n_cand]
    const int32_t *cs_cell;  // [n_cand] first row-major cell of the tuple
    const int32_t *cs_grid;  // [NL*NL + 1] start offset of the tuples with (i0, i1); last = n_cand
    const float2 *cell_xy;   // [n_cells] plane coordinates of each cell (vga_heatmap.h:52-53), built on the device
    // 3-pair arrays only: direct table over (i0, i1, i2) -> {first row-major cell of that LUT tuple or -1, x, y, 0}
    const int4 *peak_tab;    // [NL][NL][NL]
    unsigned long long *stats; // optional [5]: frames resolved by the first box / a wider box / the full scan / the direct tuple look-up; [4]: lags certified without the l.l product
    int32_t n_cand, n_cells, half_w, half_h;
    int32_t opaque_four;     // always 4; see at_fused_imma.cu
    int32_t debug_skip;      // profiling knob (env AT_DEBUG_SKIP): bit0 skip prep, bit1 skip MMA loop, bit2 skip epilogue
    float px_per_m;
    unsigned long long now_us;
    unsigned long long *prof; // VARIANT=prof builds only: [4 roles][8 sections] cycle counters (MMA issue, prep, epilogue quarter 0, quarters 1-3)
};

struct AtShape { int n_mics, n_bits, max_shift; };

// at_fused_imad.cu -- returns cudaErrorInvalidValue for a shape with no instantiation
cudaError_t at_launch_fused_imad(const AtShape &shape, const AtFusedParams &p, int sm_count, cudaStream_t st);
bool at_fused_imad_supports(const AtShape &shape);
// at_fused_imma.cu
cudaError_t at_launch_fused_imma(const AtShape &shape, const AtFusedParams &p, int sm_count, cudaStream_t st);
bool at_fused_imma_supports(const AtShape &shape);
// at_fused_imma_cta.cu -- CTA-per-frame tensor kernel for general arrays (4 / 8 mics, 1024 / 4096 samples)
cudaError_t at_launch_fused_imma_cta(const AtShape &shape, const AtFusedParams &p, int sm_count, cudaStream_t st);
bool at_fused_imma_cta_supports(const AtShape &shape);

// at_fused_umma.cu -- tcgen05 (UMMA) polyphase kernel, 3 mics x 1024 samples
// redo: device scratch of 4 * (n_frames + 1) bytes whose first word is zero (certified pass + exact pass over its list), or NULL
cudaError_t at_launch_fused_umma(const AtShape &shape, const AtFusedParams &p, uint32_t *redo, int sm_count, cudaStream_t st);
bool at_fused_umma_window_ok(const int16_t *window, int n);
cudaError_t at_run_microbench_umma(int which, int sm_count, double *gops, double *mhz, cudaStream_t st);
bool at_fused_umma_supports(const AtShape &shape);
// at_fused_umma_m.cu -- tcgen05 (UMMA) kernel for 8-microphone arrays, 1024 / 4096 samples
cudaError_t at_launch_fused_umma_m(const AtShape &shape, const AtFusedParams &p, int sm_count, cudaStream_t st);
bool at_fused_umma_m_supports(const AtShape &shape);

// at_aux.cu -- small kernels
cudaError_t at_launch_mics_triangle(float d_ab, float d_bc, float d_ca, int mirror, float *d_xy, cudaStream_t st);
cudaError_t at_launch_lut_build(const float *d_mic_xy, int n_mics, int L, float rate_hz, float speed,
                                int half_w, int half_h, float px_per_m, float height, uint8_t *d_lut,
                                cudaStream_t st);
cudaError_t at_launch_cell_xy(int half_w, int half_h, float px_per_m, float2 *d_xy, cudaStream_t st);
// AT_LUT_POINTS: the same table for arbitrary 3-D candidate positions d_points [n_points][3]; also fills d_xy with their (x, y)
cudaError_t at_launch_lut_points(const float *d_mic_xy, int n_mics, int L, float rate_hz, float speed, const float *d_points,
                                 int n_points, uint8_t *d_lut, float2 *d_xy, cudaStream_t st);
cudaError_t at_launch_write_out(const int16_t *d_ring, int head, int n_bits, int16_t *d_out, long long *d_power,
                                cudaStream_t st);
cudaError_t at_launch_shift8(int16_t *d_x, int n, cudaStream_t st);
cudaError_t at_launch_window(int16_t *d_x, int n, const int16_t *d_window, cudaStream_t st);
cudaError_t at_launch_average(long long *d_est, int32_t *d_est_best, unsigned long long *d_est_time,
                              const long long *d_fresh, const uint8_t *d_gate, size_t n_arrays, int n_pairs,
                              int L, unsigned long long now_us, const float *d_decay /*NULL: compute*/,
                              cudaStream_t st);
cudaError_t at_launch_heatmap(const long long *d_corr, size_t n_arrays, int n_pairs, int L,
                              const uint8_t *d_lut, const uint8_t *d_cand_idx, const int32_t *d_cand_cell,
                              int n_cand, int n_cells, const float2 *d_cell_xy,
                              int32_t *d_cell, long long *d_highest, float *d_xy, uint8_t *d_classes,
                              cudaStream_t st);
cudaError_t at_launch_admissible_lags(const long long *d_curves, size_t n_frames, int n_pairs, int L, const int32_t *d_lmax,
                                      int32_t *d_lags, cudaStream_t st);
cudaError_t at_launch_synth(unsigned long long seed, unsigned flags, size_t first, size_t n_frames, int n_mics,
                            int n_bits, int n_cells, const int32_t *d_delay_q8, uint8_t *d_adc, int32_t *d_heads,
                            int32_t *d_cell, cudaStream_t st);
cudaError_t at_launch_stream_push(int n_mics, int n_bits, size_t n_arrays, size_t n_ticks, const uint8_t *d_samples,
                                  uint8_t *d_hist, long long *d_count, int32_t *d_fired, uint8_t *d_frames, int32_t *d_heads,
                                  cudaStream_t st);
// at_gccphat.cu -- hand-written FFT / GCC-PHAT variant (crossover study, not a reference algorithm)
void at_gccphat_twiddles(int n_bits, float2 *h_tw);      // exp(-2 pi i n / 2N), n < 2N
cudaError_t at_launch_gccphat(int n_mics, int n_bits, int L, const uint8_t *d_adc, const int32_t *d_heads,
                              const int16_t *d_window, size_t n_frames, const float2 *d_tw, void *d_spec, int fg_log2, void *d_nyq,
                              int32_t *d_lags, float *d_peak, cudaStream_t st);
// at_gccphat_dft.cu -- the inverse side as one tcgen05 contraction with a constant cos / -sin operand
void at_gccphat_dft_tiles(int n_bits, int L, uint16_t *h_out);     // (N / 32) * 8192 fp16 entries
int at_gccphat_dft_ps_log2(int n_mics);                            // pair slots per frame = 2^this; frames per group = 256 >> this
cudaError_t at_launch_gccphat_dft(int n_mics, int n_bits, int L, size_t n_frames, const void *d_a_tiles, const void *d_spec_tiled,
                                  const void *d_nyq, int32_t *d_lags, float *d_peak, cudaStream_t st);
cudaError_t at_run_microbench(int which, int sm_count, double *gops, double *mhz, cudaStream_t st);

void at_count_launch(unsigned n = 1);
