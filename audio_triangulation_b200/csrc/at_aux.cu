// at_aux.cu -- the small kernels around the fused path: geometry + lag LUT, the per-stage
// kernels behind the drop-in symbols, the temporal average, the stand-alone likelihood map,
// the synthetic frame generator and the pipe-rate microbenchmarks.
#include <limits.h>

#include "at_fused_common.cuh"
#include "at_synth.h"

namespace atk {

// ------------------------------------------------------------------ geometry
// ref: components/microphones.c:9-33 (the reference build has MIRROR on, ROTATE off).
// float32 with explicit round-to-nearest ops so nvcc cannot contract into FMA.
__global__ void mics_triangle_kernel(float dAB, float dBC, float dCA, int mirror, float *xy)
{
    if (threadIdx.x || blockIdx.x) return;
    const float num = __fsub_rn(__fadd_rn(__fmul_rn(dAB, dAB), __fmul_rn(dCA, dCA)), __fmul_rn(dBC, dBC));
    const float xC = __fdiv_rn(num, __fmul_rn(2.0f, dAB));
    const float yC = __fsqrt_rn(fmaxf(0.0f, __fsub_rn(__fmul_rn(dCA, dCA), __fmul_rn(xC, xC))));
    const float px[3] = {0.0f, dAB, xC};
    const float py[3] = {0.0f, 0.0f, __fmul_rn(yC, mirror ? -1.0f : 1.0f)};
    const float cx = __fdiv_rn(__fadd_rn(__fadd_rn(px[0], px[1]), px[2]), 3.0f);
    const float cy = __fdiv_rn(__fadd_rn(__fadd_rn(py[0], py[1]), py[2]), 3.0f);
    for (int m = 0; m < 3; m++) {
        xy[2 * m] = __fsub_rn(px[m], cx);
        xy[2 * m + 1] = __fsub_rn(py[m], cy);
    }
}

// ref: components/vga/vga_heatmap.h:11-13
__device__ __forceinline__ float norm3(float x, float y, float z)
{
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

// ref: components/vga/vga_heatmap.h:50-92, one thread per cell.
__global__ void lut_build_kernel(const float *mic_xy, int n_mics, int L, float rate_hz, float speed, int half_w,
                                 int half_h, float px_per_m, float height, uint8_t *lut)
{
    const int W = 2 * half_w + 1, H = 2 * half_h + 1, cells = W * H;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    const int y = c / W, x = c % W;
    float xm = __fdiv_rn((float)(x - half_w), px_per_m);
    float ym = __fdiv_rn((float)(half_h - y), px_per_m);
    float zm = height;
    const float k = __fdiv_rn(height, norm3(zm, xm, ym));
    xm = __fmul_rn(xm, k); ym = __fmul_rn(ym, k); zm = __fmul_rn(zm, k);
    float dist[AT_MAX_MICS_I];
    for (int m = 0; m < n_mics; m++)
        dist[m] = norm3(zm, __fsub_rn(xm, mic_xy[2 * m]), __fsub_rn(ym, mic_xy[2 * m + 1]));
    int p = 0;
    for (int i = 0; i < n_mics; i++)
        for (int j = i + 1; j < n_mics; j++, p++) {
            const float dt = __fdiv_rn(__fsub_rn(dist[j], dist[i]), speed);
            int s = (int)roundf(__fmul_rn(dt, rate_hz));
            s = s < -L ? -L : (s > L ? L : s);
            lut[(size_t)p * cells + c] = (uint8_t)(s + L);
        }
}

// The general (3-D) form of the table: candidate positions are given, everything after vga_heatmap.h:60 is the reference's
// arithmetic unchanged (distances :63-65, time differences :68-70, roundf :72-74, clamp :76-87, index :88-90).
__global__ void lut_points_kernel(const float *mic_xy, int n_mics, int L, float rate_hz, float speed, const float *points,
                                  int n_points, uint8_t *lut, float2 *xy)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_points) return;
    const float xm = points[3 * c], ym = points[3 * c + 1], zm = points[3 * c + 2];
    float dist[AT_MAX_MICS_I];
    for (int m = 0; m < n_mics; m++)
        dist[m] = norm3(zm, __fsub_rn(xm, mic_xy[2 * m]), __fsub_rn(ym, mic_xy[2 * m + 1]));
    int p = 0;
    for (int i = 0; i < n_mics; i++)
        for (int j = i + 1; j < n_mics; j++, p++) {
            const float dt = __fdiv_rn(__fsub_rn(dist[j], dist[i]), speed);
            int s = (int)roundf(__fmul_rn(dt, rate_hz));
            s = s < -L ? -L : (s > L ? L : s);
            lut[(size_t)p * n_points + c] = (uint8_t)(s + L);
        }
    xy[c] = make_float2(xm, ym);
}

// ref: components/vga/vga_heatmap.h:52-53 -- plane coordinates of every cell (IEEE division, as the host does)
__global__ void cell_xy_kernel(int half_w, int half_h, float px_per_m, float2 *xy)
{
    const int W = 2 * half_w + 1, cells = W * (2 * half_h + 1);
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    xy[c] = make_float2(__fdiv_rn((float)(c % W - half_w), px_per_m), __fdiv_rn((float)(half_h - c / W), px_per_m));
}

// ------------------------------------------------------------------ drop-in stage kernels
// ref: components/rolling_buffer.c:43-71 for an arbitrary int16 ring (single block).
__global__ void write_out_kernel(const int16_t *ring, int head, int n_bits, int16_t *out, long long *power)
{
    __shared__ long long red[32];
    __shared__ int mean_s;
    const int n = 1 << n_bits, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    long long tot = 0;
    for (int i = tid; i < n; i += blockDim.x) tot += ring[i];
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) red[warp] = tot;
    __syncthreads();
    if (tid == 0) {
        long long t = 0;
        for (int w = 0; w < nw; w++) t += red[w];
        mean_s = (int)(short)(t >> n_bits);
    }
    __syncthreads();
    const int mean = mean_s;
    long long pw = 0;
    for (int i = tid; i < n; i += blockDim.x) {
        const int v = (int)(short)((int)ring[(head + i) & (n - 1)] - mean);
        out[i] = (int16_t)v;
        pw += (long long)v * v;
    }
    for (int o = 16; o > 0; o >>= 1) pw += __shfl_xor_sync(0xffffffffu, pw, o);
    __syncthreads();
    if (lane == 0) red[warp] = pw;
    __syncthreads();
    if (tid == 0) {
        long long t = 0;
        for (int w = 0; w < nw; w++) t += red[w];
        *power = t;
    }
}

// ref: components/buffer.c:15-16
__global__ void shift8_kernel(int16_t *x, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = (int16_t)((unsigned)(int)x[i] << 8);
}

// ref: components/buffer.c:6-10
__global__ void window_kernel(int16_t *x, int n, const int16_t *w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = (int16_t)(((int)x[i] * (int)w[i]) >> 15);
}

// ------------------------------------------------------------------ temporal average
// ref: components/correlations.c:38-63.  One warp per (array, pair).
__global__ void average_kernel(long long *est, int32_t *est_best, unsigned long long *est_time,
                               const long long *fresh, const uint8_t *gate, size_t n_arrays, int n_pairs, int L,
                               unsigned long long now_us, const float *decay_in)
{
    const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_arrays * (size_t)n_pairs) return;
    const size_t a = w / n_pairs;
    if (gate && !gate[a]) return;
    const int NL = 2 * L + 1;
    float decay;
    if (decay_in) decay = decay_in[w];
    else {
        const float dt = __fdiv_rn(__ull2float_rn(now_us - est_time[w]), 1e6f);       // :42
        decay = (float)(1.0 - exp((double)__fdiv_rn(-dt, 0.5f)));                     // :43
    }
    Best b = {LLONG_MIN, 0x7fffffff};
    for (int i = lane; i < NL; i += 32) {
        const long long e = est[w * NL + i], nw = fresh[w * NL + i];
        const float step = __fmul_rn(__ll2float_rn(nw - e), decay);                   // :49
        const long long r = __float2ll_rz(__fadd_rn(__ll2float_rn(e), step));
        est[w * NL + i] = r;
        if (r > b.v) { b.v = r; b.i = i; }
    }
    b = warp_best(b);                                                                 // :52-60
    if (lane == 0) { est_best[w] = b.i - L; est_time[w] = now_us; }                   // :62
}

// ------------------------------------------------------------------ arg-max inside each pair's admissible lag window
// One warp per (frame, pair): first-max arg-max (correlations.c:20-23) over |s| <= lmax[pair].
__global__ void admissible_lags_kernel(const long long *curves, size_t n_items, int n_pairs, int L, const int32_t *lmax,
                                       int32_t *lags)
{
    const size_t item = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= n_items) return;
    const int lane = threadIdx.x & 31, NL = 2 * L + 1, lp = lmax[item % n_pairs];
    const long long *c = curves + item * NL;
    Best b = {LLONG_MIN, 0x7fffffff};
    for (int li = L - lp + lane; li <= L + lp; li += 32) {
        const long long v = c[li];
        if (v > b.v) { b.v = v; b.i = li; }
    }
    b = warp_best(b);
    if (lane == 0) lags[item] = b.i - L;
}

// ------------------------------------------------------------------ stand-alone likelihood map
// ref: components/vga/vga_heatmap.h:96-126 for arbitrary curves; one block per array.
__global__ void heatmap_kernel(const long long *corr, int n_pairs, int L, const uint8_t *lut, const uint8_t *cand_idx,
                               const int32_t *cand_cell, int n_cand, int n_cells, const float2 *cell_xy,
                               int32_t *cell, long long *highest, float *xy, uint8_t *classes)
{
    extern __shared__ long long curve[];   // [P][NL]
    __shared__ long long red_v[32];
    __shared__ int red_i[32];
    const int NL = 2 * L + 1, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const size_t a = blockIdx.x;
    for (int i = tid; i < n_pairs * NL; i += blockDim.x) curve[i] = corr[a * n_pairs * NL + i];
    __syncthreads();
    Best b = {LLONG_MIN, 0x7fffffff};
    for (int c = tid; c < n_cand; c += blockDim.x) {
        long long like = 0;
        for (int p = 0; p < n_pairs; p++) like += curve[p * NL + cand_idx[(size_t)p * n_cand + c]];
        if (like > b.v) { b.v = like; b.i = c; }
    }
    b = warp_best(b);
    if (lane == 0) { red_v[warp] = b.v; red_i[warp] = b.i; }
    __syncthreads();
    if (warp == 0) {
        Best r = {LLONG_MIN, 0x7fffffff};
        if (lane < nw) { r.v = red_v[lane]; r.i = red_i[lane]; }
        r = warp_best(r);
        if (lane == 0) {
            const int ci = cand_cell[r.i];
            red_v[0] = r.v;
            if (cell) cell[a] = ci;
            if (highest) highest[a] = r.v;
            if (xy) {
                xy[2 * a] = cell_xy[ci].x;        // vga_heatmap.h:52-53, tabulated per cell (or the candidate point's x, y)
                xy[2 * a + 1] = cell_xy[ci].y;
            }
        }
    }
    if (!classes) return;
    __syncthreads();
    const long long top = red_v[0];
    const long long tw = (top * 63) >> 6, tg = (top * 31) >> 5, tr = (top * 15) >> 4, tb = (top * 7) >> 3;
    for (int c = tid; c < n_cells; c += blockDim.x) {
        long long like = 0;
        for (int p = 0; p < n_pairs; p++) like += curve[p * NL + lut[(size_t)p * n_cells + c]];
        classes[a * n_cells + c] = like >= tw ? 15 : like >= tg ? 3 : like >= tr ? 8 : like >= tb ? 5 : 0;
    }
}

// ------------------------------------------------------------------ streaming front end
// ref: components/rolling_buffer.c:16-41, :73-85 and the capture loop of sample_compute.h:55-99, for many
// independent arrays.  The reference's running sums over the newer / older half of the ring are plain window
// sums of the sample stream, so a block of ticks is processed in parallel: one CTA per array, thread j owns the
// 16 positions [16j, 16j+16) of the sequence  seq = [ last N samples (zero-filled since the last reset) | new block ],
// per-thread segment sums -> block prefix -> every thread slides the two half-window sums over its ticks
// (the windows start 512 / 1024 positions = 32 / 64 segments back, i.e. on other threads' segment boundaries)
// -> first tick with  out > (2 << 2(n_bits-1)) + in  and all rings full -> block minimum.
// State per array: hist uint8 [M][N] (chronological), count = pushes since the last reset.
template <int NMICS, int NBITS>
__global__ void __launch_bounds__(128) stream_push_kernel(const uint8_t *samples /*[A][ticks][M]*/, int n_ticks, uint8_t *hist,
                                                          long long *count, int32_t *fired_tick, uint8_t *frames,
                                                          int32_t *heads)
{
    constexpr int N = 1 << NBITS, H = N / 2, SEG = (2 * N) / 128;       // SEG = 16 at N = 1024
    static_assert(SEG == 16, "one 16-byte segment per thread");
    __shared__ __align__(16) uint8_t seq[NMICS][2 * N];
    __shared__ int pre1[NMICS][129];
    __shared__ int pre2[NMICS][129];
    __shared__ int first_s;
    const int tid = threadIdx.x;
    const size_t a = blockIdx.x;
    const long long c0 = count[a];
    const int fill0 = c0 < N ? (int)c0 : N;
    if (tid == 0) first_s = 0x7fffffff;
    // stage seq: history (16-byte copies) and the new block (de-interleave [tick][mic])
    for (int m = 0; m < NMICS; m++) {
        if (tid < N / 16) reinterpret_cast<uint4 *>(&seq[m][0])[tid] = reinterpret_cast<const uint4 *>(hist + (a * NMICS + m) * N)[tid];
    }
    for (int i = tid; i < n_ticks * NMICS; i += 128) seq[i % NMICS][N + i / NMICS] = samples[a * (size_t)n_ticks * NMICS + i];
    for (int i = tid; i < (N - n_ticks) * NMICS; i += 128) seq[i % NMICS][N + n_ticks + i / NMICS] = 0;
    __syncthreads();
    // segment sums and block-wide exclusive prefix (warp scan + cross-warp offsets)
    int s1[NMICS], s2[NMICS];
#pragma unroll
    for (int m = 0; m < NMICS; m++) {
        const uint4 v = reinterpret_cast<const uint4 *>(&seq[m][0])[tid];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        unsigned t1 = 0, t2 = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { t1 = __dp4a(w[k], 0x01010101u, t1); t2 = __dp4a(w[k], w[k], t2); }
        s1[m] = (int)t1; s2[m] = (int)t2;
    }
#pragma unroll
    for (int m = 0; m < NMICS; m++) {
        int i1 = s1[m], i2 = s2[m];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u1 = __shfl_up_sync(0xffffffffu, i1, o), u2 = __shfl_up_sync(0xffffffffu, i2, o);
            if ((tid & 31) >= o) { i1 += u1; i2 += u2; }
        }
        pre1[m][tid + 1] = i1; pre2[m][tid + 1] = i2;        // inclusive within the warp, fixed up below
    }
    __syncthreads();
    if (tid < NMICS) {
        const int m = tid;
        int o1 = 0, o2 = 0;
        pre1[m][0] = 0; pre2[m][0] = 0;
        for (int w = 0; w < 4; w++) {                           // add the totals of the preceding warps
            const int e1 = pre1[m][32 * w + 32], e2 = pre2[m][32 * w + 32];
            if (w) for (int k = 1; k <= 32; k++) { pre1[m][32 * w + k] += o1; pre2[m][32 * w + k] += o2; }
            o1 += e1; o2 += e2;
        }
    }
    __syncthreads();
    // slide over this thread's ticks: position pos = 16 tid + e  <->  tick t = pos - N
    if (16 * tid + 15 >= N && 16 * tid < N + n_ticks) {
        long long cur1[NMICS], cur2[NMICS], mid1[NMICS], mid2[NMICS], old1[NMICS], old2[NMICS];
#pragma unroll
        for (int m = 0; m < NMICS; m++) {   // prefix sums S(pos) excluding pos, at pos, pos-512, pos-1024 (segment starts)
            cur1[m] = pre1[m][tid]; cur2[m] = pre2[m][tid];
            mid1[m] = pre1[m][tid - 32]; mid2[m] = pre2[m][tid - 32];
            old1[m] = tid >= 64 ? pre1[m][tid - 64] : 0; old2[m] = tid >= 64 ? pre2[m][tid - 64] : 0;
        }
        const long long thr = 2LL << (2 * (NBITS - 1));         // sample_compute.h:21
        for (int e = 0; e < 16; e++) {
            const int pos = 16 * tid + e, t = pos - N;
            long long out = 0, in = 0;
#pragma unroll
            for (int m = 0; m < NMICS; m++) {                   // include position pos / pos-512 / pos-1024
                const int x = seq[m][pos], y = seq[m][pos - H], z = pos >= N ? seq[m][pos - N] : 0;
                cur1[m] += x; cur2[m] += x * x; mid1[m] += y; mid2[m] += y * y;
                if (pos >= N) { old1[m] += z; old2[m] += z * z; }
                const long long it = cur1[m] - mid1[m], ip = cur2[m] - mid2[m];          // newest N/2 samples
                const long long ot = mid1[m] - old1[m], op = mid2[m] - old2[m];          // the N/2 before them
                in += ip * H - it * it;                                                  // rolling_buffer.c:73-78
                out += op * H - ot * ot;                                                 // rolling_buffer.c:80-85
            }
            if (t >= 0 && t < n_ticks && fill0 + t + 1 >= N && out > thr + in) { atomicMin(&first_s, t); break; }
        }
    }
    __syncthreads();
    const int first = first_s;                                    // 0-based tick of the first firing, or INT_MAX
    if (first == 0x7fffffff) {
        // no onset: keep the last N samples, chronological
        for (int m = 0; m < NMICS; m++)
            for (int i = tid; i < N; i += 128) hist[(a * NMICS + m) * N + i] = seq[m][n_ticks + i];
        if (tid == 0) { count[a] = c0 + n_ticks; fired_tick[a] = -1; }
        return;
    }
    // onset: emit the ring as the reference holds it (ring order + head), then restart the capture
    const int head = (int)((c0 + first + 1) & (N - 1));
    for (int m = 0; m < NMICS; m++)
        for (int i = tid; i < N; i += 128) {
            const uint8_t v = seq[m][first + 1 + i];              // chronological sample i of the captured frame
            if (frames) frames[(a * NMICS + m) * N + ((head + i) & (N - 1))] = v;
            const int rest = n_ticks - (first + 1);               // ticks after the onset start the next capture
            hist[(a * NMICS + m) * N + i] = i >= N - rest ? seq[m][N + first + 1 + (i - (N - rest))] : 0;
        }
    if (tid == 0) { count[a] = n_ticks - (first + 1); fired_tick[a] = first + 1; if (heads) heads[a] = head; }
}

// ------------------------------------------------------------------ synthetic frames
// One thread per 4 consecutive ring slots of one (frame, mic); bytes identical to at_synth_host.
__global__ void synth_kernel(unsigned long long seed, unsigned flags, unsigned long long first,
                             unsigned long long n_frames, int n_mics, int n_bits, int n_cells,
                             const int32_t *delay_q8, uint8_t *adc, int32_t *heads, int32_t *cells)
{
    const int n = 1 << n_bits, quads = n >> 2;
    const unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long per_frame = (unsigned long long)n_mics * quads;
    if (gid >= n_frames * per_frame) return;
    const unsigned long long fl = gid / per_frame, f = first + fl;
    const int rem = (int)(gid % per_frame), mic = rem / quads, j0 = (rem % quads) * 4;
    const at_synth_frame fp = at_synth_frame_params(seed, flags, f, n_cells, n_bits);
    int32_t dq = delay_q8[(size_t)fp.cell * n_mics + mic];
    if (flags & AT_SYNTH_F_INTEGER_DELAYS) dq = (dq + 128) & ~255;
    uint32_t word = 0;
    for (int k = 0; k < 4; k++) {
        const int i = (j0 + k - fp.head) & (n - 1);
        word |= (uint32_t)at_synth_sample(seed, f, fp, mic, i, dq) << (8 * k);
    }
    reinterpret_cast<uint32_t *>(adc)[gid] = word;
    if (rem == 0) {
        if (heads) heads[fl] = fp.head;
        if (cells) cells[fl] = fp.cell;
    }
}

// ------------------------------------------------------------------ pipe-rate microbenchmarks
// Every instruction takes an operand produced by the previous iteration (a different accumulator),
// so ptxas can neither hoist the product nor fold the chain; 8 independent chains per thread.
template <int WHICH>
__global__ void __launch_bounds__(256) ubench_kernel(int iters, int seed, long long *sink, long long *cycles)
{
    const int tid = threadIdx.x;
    long long t0 = clock64();
    if (WHICH == 0) {          // mad.wide.s32 (int32 x int32 + int64)
        long long a[8];
        const int y = seed * 3 + 1;
        for (int j = 0; j < 8; j++) a[j] = j + tid;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int x = (int)a[(j + 1) & 7];
                asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(a[j]) : "r"(x), "r"(y));
            }
        }
        long long s = 0;
        for (int j = 0; j < 8; j++) s += a[j];
        if (s == 0x1234567) sink[0] = s;
    } else if (WHICH == 1 || WHICH == 2 || WHICH == 3) {   // mad.lo.s32 / dp2a / dp4a
        int a[8];
        const int y = seed * 3 + 1;
        for (int j = 0; j < 8; j++) a[j] = j + tid;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int x = a[(j + 1) & 7];
                if (WHICH == 1) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(a[j]) : "r"(x), "r"(y));
                if (WHICH == 2) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(a[j]) : "r"(x), "r"(y));
                if (WHICH == 3) asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(a[j]) : "r"(x), "r"(y));
            }
        }
        int s = 0;
        for (int j = 0; j < 8; j++) s += a[j];
        if (s == 0x1234567) sink[0] = s;
    } else if (WHICH == 4) {   // IMMA m16n8k32 s8: 4 independent accumulator tiles per warp
        int c[4][4];
        unsigned a0 = seed + tid, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = seed * 11 + tid, b1 = b0 * 13;
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 4; k++) c[j][k] = 0;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < 4; j++)
                asm volatile(
                    "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                    : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3])
                    : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
        int s = 0;
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 4; k++) s += c[j][k];
        if (s == 0x1234567) sink[0] = s;
    } else if (WHICH == 5) {   // LDS.128 streaming, conflict-free
        __shared__ uint4 buf[1024];
        for (int i = tid; i < 1024; i += 256) buf[i] = make_uint4(i, seed, tid, 1);
        __syncthreads();
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint4 v;
                const uint32_t addr = smem_u32(&buf[(tid + 256 * j) & 1023]);
                asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
                acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
            }
        }
        if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x1234567) sink[0] = acc.x;
    } else {                   // DFMA
        double a[8];
        const double y = 1e-9 * seed;
        for (int j = 0; j < 8; j++) a[j] = j + 1e-3 * tid;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < 8; j++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[j]) : "d"(a[(j + 1) & 7]), "d"(y));
        }
        double s = 0;
        for (int j = 0; j < 8; j++) s += a[j];
        if (s == 12345.678) sink[0] = (long long)s;
    }
    long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

} // namespace atk

using namespace atk;

cudaError_t at_launch_mics_triangle(float d_ab, float d_bc, float d_ca, int mirror, float *d_xy, cudaStream_t st)
{
    mics_triangle_kernel<<<1, 32, 0, st>>>(d_ab, d_bc, d_ca, mirror, d_xy);
    at_count_launch();
    return cudaGetLastError();
}

cudaError_t at_launch_lut_build(const float *d_mic_xy, int n_mics, int L, float rate_hz, float speed, int half_w,
                                int half_h, float px_per_m, float height, uint8_t *d_lut, cudaStream_t st)
{
    const int cells = (2 * half_w + 1) * (2 * half_h + 1);
    lut_build_kernel<<<(cells + 127) / 128, 128, 0, st>>>(d_mic_xy, n_mics, L, rate_hz, speed, half_w, half_h,
                                                          px_per_m, height, d_lut);
    at_count_launch();
    return cudaGetLastError();
}

cudaError_t at_launch_lut_points(const float *d_mic_xy, int n_mics, int L, float rate_hz, float speed, const float *d_points,
                                 int n_points, uint8_t *d_lut, float2 *d_xy, cudaStream_t st)
{
    lut_points_kernel<<<(n_points + 127) / 128, 128, 0, st>>>(d_mic_xy, n_mics, L, rate_hz, speed, d_points, n_points, d_lut, d_xy);
    at_count_launch();
    return cudaGetLastError();
}

cudaError_t at_launch_cell_xy(int half_w, int half_h, float px_per_m, float2 *d_xy, cudaStream_t st)
{
    const int cells = (2 * half_w + 1) * (2 * half_h + 1);
    cell_xy_kernel<<<(cells + 127) / 128, 128, 0, st>>>(half_w, half_h, px_per_m, d_xy);
    at_count_launch();
    return cudaGetLastError();
}

cudaError_t at_launch_write_out(const int16_t *d_ring, int head, int n_bits, int16_t *d_out, long long *d_power,
                                cudaStream_t st)
{
    write_out_kernel<<<1, 256, 0, st>>>(d_ring, head, n_bits, d_out, d_power);
    at_count_launch();
    return cudaGetLastError();
}

cudaError_t at_launch_shift8(int16_t *d_x, int n, cudaStream_t st)
{
    shift8_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_x, n);
    at_count_launch();
    return cudaGetLastError();
}

cudaError_t at_launch_window(int16_t *d_x, int n, const int16_t *d_window, cudaStream_t st)
{
    window_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_x, n, d_window);
    at_count_launch();
    return cudaGetLastError();
}

cudaError_t at_launch_average(long long *d_est, int32_t *d_est_best, unsigned long long *d_est_time,
                              const long long *d_fresh, const uint8_t *d_gate, size_t n_arrays, int n_pairs, int L,
                              unsigned long long now_us, const float *d_decay, cudaStream_t st)
{
    const size_t warps = n_arrays * (size_t)n_pairs;
    if (!warps) return cudaSuccess;
    const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
    average_kernel<<<blocks, 256, 0, st>>>(d_est, d_est_best, d_est_time, d_fresh, d_gate, n_arrays, n_pairs, L,
                                           now_us, d_decay);
    at_count_launch();
    return cudaGetLastError();
}

cudaError_t at_launch_admissible_lags(const long long *d_curves, size_t n_frames, int n_pairs, int L, const int32_t *d_lmax,
                                      int32_t *d_lags, cudaStream_t st)
{
    const size_t n_items = n_frames * (size_t)n_pairs;
    if (!n_items) return cudaSuccess;
    admissible_lags_kernel<<<(unsigned)((n_items + 7) / 8), 256, 0, st>>>(d_curves, n_items, n_pairs, L, d_lmax, d_lags);
    at_count_launch();
    return cudaGetLastError();
}

cudaError_t at_launch_heatmap(const long long *d_corr, size_t n_arrays, int n_pairs, int L, const uint8_t *d_lut,
                              const uint8_t *d_cand_idx, const int32_t *d_cand_cell, int n_cand, int n_cells,
                              const float2 *d_cell_xy, int32_t *d_cell, long long *d_highest,
                              float *d_xy, uint8_t *d_classes, cudaStream_t st)
{
    if (!n_arrays) return cudaSuccess;
    const int smem = n_pairs * (2 * L + 1) * (int)sizeof(long long);
    heatmap_kernel<<<(unsigned)n_arrays, 256, smem, st>>>(d_corr, n_pairs, L, d_lut, d_cand_idx, d_cand_cell, n_cand,
                                                          n_cells, d_cell_xy, d_cell, d_highest, d_xy, d_classes);
    at_count_launch();
    return cudaGetLastError();
}

cudaError_t at_launch_synth(unsigned long long seed, unsigned flags, size_t first, size_t n_frames, int n_mics,
                            int n_bits, int n_cells, const int32_t *d_delay_q8, uint8_t *d_adc, int32_t *d_heads,
                            int32_t *d_cell, cudaStream_t st)
{
    const unsigned long long total = (unsigned long long)n_frames * n_mics * ((1u << n_bits) >> 2);
    if (!total) return cudaSuccess;
    const unsigned long long blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidValue;
    synth_kernel<<<(unsigned)blocks, 256, 0, st>>>(seed, flags, first, n_frames, n_mics, n_bits, n_cells, d_delay_q8,
                                                   d_adc, d_heads, d_cell);
    at_count_launch();
    return cudaGetLastError();
}

cudaError_t at_launch_stream_push(int n_mics, int n_bits, size_t n_arrays, size_t n_ticks, const uint8_t *d_samples,
                                  uint8_t *d_hist, long long *d_count, int32_t *d_fired, uint8_t *d_frames, int32_t *d_heads,
                                  cudaStream_t st)
{
    if (!n_arrays) return cudaSuccess;
    if (n_mics != 3 || n_bits != 10) return cudaErrorInvalidValue;
    stream_push_kernel<3, 10><<<(unsigned)n_arrays, 128, 0, st>>>(d_samples, (int)n_ticks, d_hist, d_count, d_fired, d_frames, d_heads);
    at_count_launch();
    return cudaGetLastError();
}

template <int WHICH>
static cudaError_t run_ubench(int sm_count, double ops_per_thread_iter, double *gops, double *mhz, cudaStream_t st)
{
    long long *d = nullptr;
    cudaError_t e = cudaMalloc(&d, 2 * sizeof(long long));
    if (e != cudaSuccess) return e;
    const int iters = 4096, blocks = sm_count * 4;   // 4 x 256 threads per SM: all blocks co-resident
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0); cudaEventCreate(&ev1);
    ubench_kernel<WHICH><<<blocks, 256, 0, st>>>(64, 1, d, d + 1);       // warm-up
    cudaEventRecord(ev0, st);
    ubench_kernel<WHICH><<<blocks, 256, 0, st>>>(iters, 1, d, d + 1);
    cudaEventRecord(ev1, st);
    at_count_launch(2);
    e = cudaEventSynchronize(ev1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    long long cyc[2] = {0, 0};
    cudaMemcpy(cyc, d, sizeof cyc, cudaMemcpyDeviceToHost);
    cudaFree(d);
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    if (e != cudaSuccess) return e;
    const double total = ops_per_thread_iter * iters * 256.0 * blocks;
    *gops = total / (ms * 1e-3) / 1e9;
    // block 0 ran for cyc[1] cycles; all blocks are co-resident, so that spans the whole launch
    if (mhz) *mhz = (double)cyc[1] / (ms * 1e-3) / 1e6;
    return cudaGetLastError();
}

cudaError_t at_run_microbench(int which, int sm_count, double *gops, double *mhz, cudaStream_t st)
{
    switch (which) {
    case 0: return run_ubench<0>(sm_count, 8, gops, mhz, st);                         // MAC
    case 1: return run_ubench<1>(sm_count, 8, gops, mhz, st);
    case 2: return run_ubench<2>(sm_count, 8 * 2, gops, mhz, st);                     // 2 MAC per IDP.2A
    case 3: return run_ubench<3>(sm_count, 8 * 4, gops, mhz, st);                     // 4 MAC per IDP.4A
    case 4: return run_ubench<4>(sm_count, 4.0 * 16 * 8 * 32 / 32.0, gops, mhz, st);  // MAC per lane per iteration
    case 5: return run_ubench<5>(sm_count, 4 * 16, gops, mhz, st);                    // bytes
    case 6: return run_ubench<6>(sm_count, 8, gops, mhz, st);
    case 7: return at_run_microbench_umma(0, sm_count, gops, mhz, st);                // dense tcgen05 int8 MMAs
    case 8: return at_run_microbench_umma(1, sm_count, gops, mhz, st);                // G frames/s of a frame's MMA sequence
    default: return cudaErrorInvalidValue;
    }
}
