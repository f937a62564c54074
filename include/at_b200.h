/* at_b200.h -- C ABI of libat_b200.so: the B200-native (sm_100a CUDA) implementation of
 * Audio-Triangulation's per-frame localization path.
 *
 * Two groups of entry points:
 *
 *  1. DROP-IN SYMBOLS.  Exactly the functions sample_compute.h calls on the compute path,
 *     with the reference's names, struct layouts and in-place semantics, so the harness (or
 *     the reference's own sample_compute.h) links against this library instead of the
 *     reference objects without source changes.  Each runs the CUDA kernels on a batch of one.
 *     They return void like the reference; a CUDA failure aborts the process with a message
 *     (there is NO CPU fallback).
 *
 *  2. BATCHED EXTENSION (at_*).  The same path for many independent frames that live in large
 *     device-resident arrays; this is what makes a GPU worthwhile.  All return 0 on success or a
 *     negative AT_E* code; at_last_error() gives the text.
 *
 * Plain C, plain pointers and sizes; no CUDA/torch types in any signature (streams are passed
 * as void* holding a cudaStream_t; NULL is CUDA's legacy default stream, as in the runtime API).
 * Citations "ref:" are relative to the reference repository's src/ directory.
 */
#ifndef AT_B200_H
#define AT_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * Reference data types (define AT_B200_NO_REFERENCE_TYPES when the reference's own headers
 * are included in the same translation unit).
 * ---------------------------------------------------------------------------------------- */
#ifndef AT_B200_NO_REFERENCE_TYPES
typedef int64_t power_t;          /* ref: components/constants.h:6 */
typedef int16_t sample_t;         /* ref: components/constants.h:7 */
typedef uint64_t absolute_time_t; /* Pico SDK type as used at ref: components/correlations.h:15 */

#define SAMPLE_RATE_HZ 50000                               /* ref: components/constants.h:10 */
#define MAX_SHIFT_SAMPLES (SAMPLE_RATE_HZ * 32 / 34300)    /* ref: components/constants.h:12 (=46) */
#define BUFFER_SIZE_BITS 10                                /* ref: components/buffer.h:5 */
#define BUFFER_SIZE (1 << BUFFER_SIZE_BITS)                /* ref: components/buffer.h:6 */
#define CORRELATION_BUFFER_SIZE (2 * MAX_SHIFT_SAMPLES + 1) /* ref: components/correlations.h:8 (=93) */

struct buffer_t {                 /* ref: components/buffer.h:8-12, 2056 B */
    sample_t buffer[BUFFER_SIZE];
    power_t power;
};

struct rolling_buffer_t {         /* ref: components/rolling_buffer.h:13-25, 2096 B */
    int head;
    power_t incoming_power;
    power_t incoming_total;
    power_t outgoing_power;
    power_t outgoing_total;
    bool is_full;
    sample_t buffer[BUFFER_SIZE];
};

struct correlations_t {           /* ref: components/correlations.h:10-16, 760 B */
    power_t correlations[CORRELATION_BUFFER_SIZE];
    int best_shift;
    absolute_time_t last_update;
};

typedef struct { float x, y; } point2d_t; /* ref: components/point.h:3-7 */
#endif /* AT_B200_NO_REFERENCE_TYPES */

/* ------------------------------------------------------------------------------------------
 * 1. Drop-in symbols
 * ---------------------------------------------------------------------------------------- */

/* replaces ref: components/rolling_buffer.h:27 (rolling_buffer.c:3-14).  Zeroes the caller's ring. */
void rolling_buffer_init(struct rolling_buffer_t *buf);
/* replaces ref: components/rolling_buffer.h:28 (rolling_buffer.c:16-41).  O(1) bookkeeping on the
 * caller-owned host struct at capture rate (one sample per 20 us); stays on the host by design --
 * the device form of the same recurrence for many arrays at once is at_stream_push(). */
void rolling_buffer_push(struct rolling_buffer_t *buf, sample_t sample);
/* replaces ref: components/rolling_buffer.h:29 (rolling_buffer.c:43-71).  GPU: un-rotate, DC removal, power. */
void rolling_buffer_write_out(const struct rolling_buffer_t *buf, struct buffer_t *dst);
/* replace ref: components/rolling_buffer.h:31-32 (rolling_buffer.c:73-85). */
power_t rolling_buffer_get_incoming_power(const struct rolling_buffer_t *buf);
power_t rolling_buffer_get_outgoing_power(const struct rolling_buffer_t *buf);
/* replaces ref: components/buffer.h:15 (buffer.c:13-18).  GPU: x <<= 8 with int16 wrap. */
void buffer_normalize_range(struct buffer_t *buf);
/* replaces ref: components/buffer.h:14 (buffer.c:4-11).  GPU: Q15 DPSS window. */
void buffer_window(struct buffer_t *buf);
/* replaces ref: components/correlations.h:18-21 (correlations.c:4-36).  GPU: 93-lag integer
 * cross-correlation, first-max arg-max, Gaussian re-weighting, time stamp. */
void correlations_init(struct correlations_t *corr, const struct buffer_t *buf_a, const struct buffer_t *buf_b);
/* replaces ref: components/correlations.h:23-25 (correlations.c:38-63).  GPU: EMA + re-arg-max. */
void correlations_average(struct correlations_t *estimate, struct correlations_t *new_data);
/* replaces ref: components/microphones.h:6-10 (microphones.c:5-61). */
extern point2d_t mic_a_location, mic_b_location, mic_c_location;
void microphones_init(void);

/* The reference reads the clock through the SDK's get_absolute_time() (correlations.c:35, :40).
 * If the program that loads this library defines that symbol it is used; otherwise the library's
 * own clock is: at_set_time_us(t) pins it to t, at_set_time_us(UINT64_MAX) returns to
 * CLOCK_MONOTONIC microseconds. */
void at_set_time_us(uint64_t now_us);
uint64_t at_get_time_us(void);

/* ------------------------------------------------------------------------------------------
 * 2. Batched extension
 * ---------------------------------------------------------------------------------------- */
#define AT_OK 0
#define AT_EINVAL (-1)   /* bad argument / unsupported shape */
#define AT_ECUDA (-2)    /* CUDA runtime or kernel failure */
#define AT_ENOGPU (-3)   /* no usable sm_100 device */
#define AT_ENOMEM (-4)

#define AT_MAX_MICS 8

/* kernel selection for the fused prep + xcorr + arg-max stage */
#define AT_KERNEL_AUTO 0
#define AT_KERNEL_IMAD 1  /* IMAD.WIDE register-tiled direct form (integer pipe) */
#define AT_KERNEL_IMMA 2  /* byte-split Toeplitz x Hankel int8 tensor-core form (exact) */
/* 3 was an ldmatrix variant of AT_KERNEL_IMMA; removed in round 2 (never faster) */
#define AT_KERNEL_UMMA 4  /* polyphase Hankel form on tcgen05 (UMMA, TMEM accumulators), exact; 3 mics x 1024 samples
                             (AUTO picks it: the reference shape), 8 mics x 1024 / 4096 samples (AUTO picks it for 8 x 4096) */

/* layout of the optional correlation-curve output */
#define AT_CORR_PACKED 0  /* int64 [F][pairs][2L+1] */
#define AT_CORR_STRUCT 1  /* struct correlations_t [F][pairs]; requires 2L+1 == 93 */

typedef struct at_config {
    int32_t device;        /* CUDA ordinal */
    int32_t n_mics;        /* 2..AT_MAX_MICS; reference: 3 */
    int32_t n_bits;        /* log2(frame length), 8..12; reference: 10 */
    int32_t max_shift;     /* L; reference: 46 (components/constants.h:12); <= 127 */
    int32_t kernel;        /* AT_KERNEL_* */
    float sample_rate_hz;  /* reference: 50000 */
    float speed_of_sound;  /* reference: 343 (components/constants.h:14) */
    /* likelihood map (ref: components/vga/vga.h:27-35) */
    int32_t half_w, half_h; /* reference: 50, 50 -> 101 x 101 cells */
    float px_per_m;        /* reference: 24 */
    float height_m;        /* reference: 1.2 */
    /* microphone coordinates in metres; n_mics == 3 && use_reference_triangle -> microphones_init() geometry */
    int32_t use_reference_triangle;
    float mic_xy[AT_MAX_MICS][2];
    /* Candidate source positions of the lag look-up table.  AT_LUT_PLANE (default): the reference's set, the half_w x
     * half_h grid of the z = height_m plane projected onto the sphere of radius height_m (vga_heatmap.h:50-60).
     * AT_LUT_POINTS: `n_points` arbitrary 3-D positions (x, y, z in metres, microphones in the z = 0 plane), e.g. a
     * hemisphere of directions (at_hemisphere_points) or a volume grid -- the general, 3-D form of the same table: expected
     * lag of pair (i, j) = roundf((|p - m_j| - |p - m_i|) / c * fs), clamped to +-L.  `cell` outputs then index the points,
     * `xy` holds their (x, y), `classes` has n_points entries. */
    int32_t lut_mode;
    int32_t n_points;
    const float *points_xyz;   /* [n_points][3], host memory, read by at_create only */
} at_config;
#define AT_LUT_PLANE 0
#define AT_LUT_POINTS 1
/* n_az x n_el directions on the upper hemisphere of radius `radius_m` (azimuth 2 pi a / n_az; elevation from the horizon
 * towards the zenith, (e + 0.5) / n_el * pi / 2), row-major [n_el][n_az][3].  A convenience for AT_LUT_POINTS. */
void at_hemisphere_points(int n_az, int n_el, float radius_m, float *xyz);

typedef struct at_context at_context;

/* Fills `cfg` with the reference configuration (3 mics, 1024 samples, +-46 lags, 50 kHz, triangle). */
void at_config_reference(at_config *cfg);
int at_create(const at_config *cfg, at_context **out);
void at_destroy(at_context *ctx);
const char *at_last_error(void);
/* Number of CUDA kernels this library has launched in the calling process (for bench accounting). */
uint64_t at_kernel_launches(void);

/* Geometry products (host copies): mic coordinates [n_mics][2]; lag LUT uint8 [pairs][cells]
 * (ref: components/vga/vga_heatmap.h:50-92), built by a CUDA kernel at at_create(). */
int at_get_mics(const at_context *ctx, float *xy);
int at_get_lut(const at_context *ctx, uint8_t *lut);
int at_shape(const at_context *ctx, int32_t *n_mics, int32_t *n_samples, int32_t *n_pairs,
             int32_t *n_lags, int32_t *n_cells);

/* Outputs of the batched path; every pointer may be NULL (that product is then not written).
 * Device API: device pointers.  Host API: host pointers. */
typedef struct at_outputs {
    int32_t *lags;      /* [F][pairs]  best_shift per pair (ref: correlations.c:20-23) */
    void *corr;         /* post-Gaussian curves, layout per corr_layout (ref: correlations.c:26-33) */
    int32_t corr_layout;
    int64_t *raw;       /* [F][pairs][2L+1] pre-Gaussian curves (ref: correlations.c:9-18) */
    int32_t *cell;      /* [F] first row-major cell reaching the likelihood maximum (ref: vga_heatmap.h:96-108) */
    int64_t *highest;   /* [F] that maximum (highest_L) */
    float *xy;          /* [F][2] plane coordinates of `cell` in metres (ref: vga_heatmap.h:52-53) */
    uint8_t *gate;      /* [F] 1 iff sum of squared lags > 4 (ref: sample_compute.h:124-134) */
    uint8_t *classes;   /* [F][cells] colour class per cell (ref: vga_heatmap.h:111-126) */
    int16_t *windowed;  /* [F][mics][N] frames after DC removal, <<8 and window (debug/parity) */
    int64_t *power;     /* [F][mics] buffer_t.power after DC removal (ref: rolling_buffer.c:68-70) */
    uint64_t *stats;    /* device API only, [5] counters ADDED to: frames whose likelihood maximum was settled by
                           the first bounded box / a widened box / the full tuple scan / the direct look-up of the
                           tuple (best_ab, best_ac, best_bc) in the LUT; [4] = frames (a subset of [3]) whose lags
                           were certified without the l.l digit product (diagnostics) */
} at_outputs;

/* The whole per-frame path, sample_compute.h:104-122 (+ the likelihood arg-max), for n_frames
 * independent frames.  adc: uint8 [F][mics][N] in ring order; heads: int32 [F] ring head per
 * frame (NULL = all 0 = chronological).  Asynchronous on `stream`. */
int at_localize_device(at_context *ctx, const uint8_t *d_adc, const int32_t *d_heads, size_t n_frames,
                       const at_outputs *d_out, void *stream);
/* Same from host memory: chunks the batch over three device slots, host->device copies on one stream, kernels on
 * the slot's stream, device->host copies on a third (both copy directions stay busy); returns when all results
 * are in host memory.  Pinned host buffers are copied directly. */
int at_localize_host(at_context *ctx, const uint8_t *h_adc, const int32_t *h_heads, size_t n_frames,
                     const at_outputs *h_out);
/* Frame-sharded over several contexts (one per GPU) driven from one host thread: context g gets
 * the contiguous frame range [g*F/G, (g+1)*F/G); results land in the caller's host arrays. */
int at_localize_host_sharded(at_context **ctxs, int n_ctx, const uint8_t *h_adc, const int32_t *h_heads,
                             size_t n_frames, const at_outputs *h_out);
int at_synchronize(at_context *ctx);
/* Multi-GPU result placement (one process per GPU).  The rank that collects the results allocates its arrays with
 * at_shared_alloc and hands the 64-byte handle to the other ranks (any channel: torch.distributed object broadcast, a
 * pipe, a file); they map the arrays with at_shared_open and pass the returned pointers as at_outputs of their own
 * at_localize_device calls.  Every rank's kernel then stores its slice of the results straight into the collector's
 * memory over NVLink: a frame-sharded run needs no gather step and no collective in the data path.
 * at_peer_enable: same idea inside one process (several contexts, one per GPU).  All idempotent / balanced by _close. */
typedef struct at_ipc_handle { unsigned char bytes[64]; } at_ipc_handle;
int at_shared_alloc(at_context *ctx, size_t bytes, void **d_ptr, at_ipc_handle *handle);
int at_shared_open(at_context *ctx, const at_ipc_handle *handle, void **d_ptr);
int at_shared_close(at_context *ctx, void *d_ptr, int opened /* 1: from at_shared_open, 0: from at_shared_alloc */);
int at_peer_enable(at_context *ctx, int peer_device);
/* Asynchronous device-to-device copy on `stream` (copy engine; either pointer may be an at_shared_open mapping). */
int at_copy_async(at_context *ctx, void *d_dst, const void *d_src, size_t bytes, void *stream);
/* Page-locked host memory for capture / result buffers: at_localize_host copies such buffers at the link rate (55 GB/s on
 * PCIe 5 x16) instead of staging pageable memory (replaces the static frame buffers of ref: sample_compute.h:24-38 in a
 * host harness).  at_host_free releases it. */
int at_host_alloc(at_context *ctx, size_t bytes, void **h_ptr);
int at_host_free(at_context *ctx, void *h_ptr);

/* Temporal stage for `n_arrays` independent arrays (ref: sample_compute.h:124-139,
 * correlations.c:38-63): where gate[i] != 0, estimate <- EMA(estimate, fresh) with the array's own
 * last_update, re-arg-max, stamp now_us.  est/fresh: struct-of-arrays, int64 [A][pairs][2L+1];
 * est_best int32 [A][pairs]; est_time uint64 [A][pairs]. Device pointers.  The decay factor 1 - exp(-dt/0.5) is
 * evaluated on the host with the reference's own libm call (the time stamps are read back, which synchronises
 * `stream` once), so the averaged curves are the reference's bit for bit. */
int at_average_device(at_context *ctx, int64_t *d_est, int32_t *d_est_best, uint64_t *d_est_time,
                      const int64_t *d_fresh, const uint8_t *d_gate, size_t n_arrays, uint64_t now_us,
                      void *stream);

/* Likelihood map of arbitrary curves (e.g. the EMA estimates): ref: vga_heatmap.h:96-126. */
int at_heatmap_device(at_context *ctx, const int64_t *d_corr /*[A][pairs][2L+1]*/, size_t n_arrays,
                      int32_t *d_cell, int64_t *d_highest, float *d_xy, uint8_t *d_classes, void *stream);

/* Physically admissible lag window of each pair: a source cannot delay one microphone against the other by more than
 * their distance, so |lag| <= ceil(d_pair * sample_rate / speed_of_sound), clipped to max_shift.  The reference scans
 * +-max_shift for every pair (components/constants.h:12 derives it from the longest side); this is the per-pair
 * refinement of SURVEY section 8(f) item 3, offered as a separate step so that the reference's results stay untouched.
 * out: int32 [pairs] in pair order (0,1), (0,2), ..., (1,2), ... */
int at_pair_max_shift(at_context *ctx, int32_t *out);
/* First-max arg-max (components/correlations.c:20-23: strict '>', ascending lag) of raw or post-Gaussian curves
 * restricted to each pair's admissible window.  d_curves: int64 [F][pairs][2L+1] (the `raw` or packed `corr` output of
 * at_localize_device); d_lags: int32 [F][pairs].  Device pointers, asynchronous on `stream`. */
int at_admissible_lags_device(at_context *ctx, const int64_t *d_curves, size_t n_frames, int32_t *d_lags, void *stream);

/* GCC-PHAT / FFT variant of the TDOA stage (hand-written, no cuFFT: register-resident radix-8 FFTs forward, one tcgen05
 * fp16 contraction over the admissible lags backward -- or inverse FFTs with AT_GCC_INVERSE=fft), for the direct-vs-FFT
 * crossover study of long frames / wide lag ranges.  NOT a reference algorithm (the reference correlates directly,
 * components/correlations.c:9-24): PHAT whitening changes the statistic, only arg-max lags are comparable.
 * Same integer frame preparation, then float32.  d_peak (optional): normalised peak value per pair. */
int at_gccphat_device(at_context *ctx, const uint8_t *d_adc, const int32_t *d_heads, size_t n_frames,
                      int32_t *d_lags /*[F][pairs]*/, float *d_peak /*[F][pairs] or NULL*/, void *stream);

/* Streaming front end (ref: components/rolling_buffer.c:16-41, :73-85; capture loop sample_compute.h:55-99) for
 * n_arrays independent arrays whose ring state lives on the device.  at_stream_push() consumes n_ticks sample triples
 * per array (n_ticks a multiple of 16, <= frame length, so at most one onset per call) and reports per array the
 * 1-based tick at which the onset gate fired (-1: not in this block) together with the captured ring (ring order) and
 * its head, ready for at_localize_device().  After an onset the array's capture restarts with the remaining ticks,
 * exactly as the reference re-initialises its rings (sample_compute.h:55-57).  Reference shape only. */
typedef struct at_stream at_stream;
int at_stream_create(at_context *ctx, size_t n_arrays, at_stream **out);
void at_stream_destroy(at_stream *s);
int at_stream_reset(at_stream *s, void *stream);
int at_stream_push(at_stream *s, const uint8_t *d_samples /*[A][ticks][mics]*/, size_t n_ticks,
                   int32_t *d_fired_tick /*[A]*/, uint8_t *d_frames /*[A][mics][N], may be NULL*/,
                   int32_t *d_heads /*[A], may be NULL*/, void *stream);

/* Synthetic multi-channel frames (host harness input; replaces the Pico ADC/DMA capture,
 * ref: components/dma_sampler.c).  Integer-only counter-based generator: the host and device
 * versions emit identical bytes.  true_cell (optional): the source's heat-map cell. */
#define AT_SYNTH_INTEGER_DELAYS 1u  /* round per-mic delays to whole samples */
#define AT_SYNTH_RANDOM_HEADS 2u    /* store frames rotated by a random ring head */
#define AT_SYNTH_KATS 4u            /* frames 0..3 of the sequence are the known-answer frames */
#define AT_SYNTH_MAX_NOISE 8u       /* every frame at the model's lowest signal-to-noise ratio */
#define AT_SYNTH_WHITE 16u          /* no source: independent uniform bytes per channel (worst case of every data-dependent shortcut) */
int at_synth_host(const at_context *ctx, uint64_t seed, uint32_t flags, size_t first_frame, size_t n_frames,
                  uint8_t *adc, int32_t *heads, int32_t *true_cell);
int at_synth_device(at_context *ctx, uint64_t seed, uint32_t flags, size_t first_frame, size_t n_frames,
                    uint8_t *d_adc, int32_t *d_heads, int32_t *d_true_cell, void *stream);

/* Pipe-rate microbenchmarks used for the roofline denominators (IMAD.WIDE, IDP.2A, IMMA int8,
 * shared-memory loads).  which: AT_UBENCH_*; returns giga-operations (MAC for arithmetic, bytes
 * for LDS) per second in *gops, measured with CUDA events on `device`. */
#define AT_UBENCH_IMAD_WIDE 0
#define AT_UBENCH_IMAD 1
#define AT_UBENCH_DP2A 2
#define AT_UBENCH_DP4A 3
#define AT_UBENCH_IMMA_S8 4
#define AT_UBENCH_LDS 5
#define AT_UBENCH_DFMA 6
#define AT_UBENCH_UMMA_I8 7     /* dense tcgen05.mma kind::i8 M128 x N256 x K32: *gops = G int8-MAC/s */
#define AT_UBENCH_UMMA_FRAME 8  /* the eight Hankel MMAs of one reference-shape frame, back to back: *gops = G frames/s */
int at_microbench(at_context *ctx, int which, double *gops, double *sm_mhz_est);

#ifdef __cplusplus
}
#endif
#endif /* AT_B200_H */
