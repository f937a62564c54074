// at_gccphat.cu -- hand-written (cuFFT-free) GCC-PHAT variant of the TDOA stage, for the crossover study of
// BASELINE config 4.  NOT a reference algorithm: the reference correlates directly in integer arithmetic
// (components/correlations.c:9-24); PHAT whitening changes the statistic, so only the arg-max lags can be compared
// (agreement rate), never the curves.  Frame preparation is the same integer path as everywhere else (DC removal,
// <<8, Q15 window), then float32:
//
//   forward_kernel  one CTA per (frame, channel pair): two real channels packed as one complex sequence, zero-padded
//                   to 2N, radix-2 decimation-in-time FFT in shared memory (precomputed twiddles), spectra separated
//                   by Hermitian symmetry and written to a global scratch [frames][mics][N+1] float2.
//   pair_kernel     one CTA per (frame, pair): G = conj(X) Y / |conj(X) Y|, Hermitian extension, inverse FFT of size
//                   2N, real part at lags -L..L, first-max arg-max.
#include "at_fused_common.cuh"

namespace atk {

template <int NB2>   // log2 of the FFT size
__device__ __forceinline__ void fft_inplace(float2 *data, const float2 *tw, bool inverse, int tid, int nthreads)
{
    constexpr int N2 = 1 << NB2, HALF = N2 >> 1;
    for (int s = 0; s < NB2; s++) {
        const int m = 1 << s;
        for (int b = tid; b < HALF; b += nthreads) {
            const int j = b & (m - 1), i = ((b >> s) << (s + 1)) + j;
            float2 w = tw[j << (NB2 - 1 - s)];              // exp(-2 pi i j / (2m))
            if (inverse) w.y = -w.y;
            const float2 u = data[i], v0 = data[i + m];
            const float2 v = make_float2(v0.x * w.x - v0.y * w.y, v0.x * w.y + v0.y * w.x);
            data[i] = make_float2(u.x + v.x, u.y + v.y);
            data[i + m] = make_float2(u.x - v.x, u.y - v.y);
        }
        __syncthreads();
    }
}

template <int NB2>
__device__ __forceinline__ int bitrev(int n) { return (int)(__brev((unsigned)n) >> (32 - NB2)); }

template <int NBITS>
__global__ void __launch_bounds__(256) gcc_forward_kernel(const uint8_t *adc, const int32_t *heads, const int16_t *window,
                                                          int n_mics, float2 *spec /*[F][M][N+1]*/)
{
    constexpr int N = 1 << NBITS, NB2 = NBITS + 1, N2 = 2 * N;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float2 *data = reinterpret_cast<float2 *>(smem_raw);          // [N2]
    float2 *tw = data + N2;                                        // [N]
    __shared__ int red[2][8];
    __shared__ int mean_s[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t f = blockIdx.x;
    const int ca = 2 * blockIdx.y, cb = ca + 1 < n_mics ? ca + 1 : -1;
    const int head = heads ? (heads[f] & (N - 1)) : 0;
    for (int k = tid; k < N; k += 256) { float sn, cs; sincospif((float)k / (float)N, &sn, &cs); tw[k] = make_float2(cs, -sn); }
    // channel sums -> floor mean (rolling_buffer.c:48-64)
    int sa = 0, sb = 0;
    for (int i = tid; i < N; i += 256) {
        sa += adc[(f * n_mics + ca) * N + i];
        if (cb >= 0) sb += adc[(f * n_mics + cb) * N + i];
    }
    for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
    if (lane == 0) { red[0][warp] = sa; red[1][warp] = sb; }
    __syncthreads();
    if (tid < 2) { int t = 0; for (int w = 0; w < 8; w++) t += red[tid][w]; mean_s[tid] = (int)(short)(t >> NBITS); }
    __syncthreads();
    // prepared samples (integer path of at_fused_common.cuh) as floats, bit-reversed placement, zero padding
    for (int n = tid; n < N2; n += 256) {
        float2 v = make_float2(0.f, 0.f);
        if (n < N) {
            const int j = (head + n) & (N - 1);
            v.x = (float)prep_sample(adc[(f * n_mics + ca) * N + j], mean_s[0], window[n]);
            if (cb >= 0) v.y = (float)prep_sample(adc[(f * n_mics + cb) * N + j], mean_s[1], window[n]);
        }
        data[bitrev<NB2>(n)] = v;
    }
    __syncthreads();
    fft_inplace<NB2>(data, tw, false, tid, 256);
    // Z = A + iB with A, B spectra of the two real channels: A[k] = (Z[k] + conj Z[-k]) / 2, B[k] = (Z[k] - conj Z[-k]) / 2i
    for (int k = tid; k <= N; k += 256) {
        const float2 z = data[k], zc = data[(N2 - k) & (N2 - 1)];
        spec[(f * n_mics + ca) * (size_t)(N + 1) + k] = make_float2(0.5f * (z.x + zc.x), 0.5f * (z.y - zc.y));
        if (cb >= 0) spec[(f * n_mics + cb) * (size_t)(N + 1) + k] = make_float2(0.5f * (z.y + zc.y), -0.5f * (z.x - zc.x));
    }
}

template <int NBITS>
__global__ void __launch_bounds__(256) gcc_pair_kernel(const float2 *spec, int n_mics, int L, int32_t *lags, float *peak)
{
    constexpr int N = 1 << NBITS, NB2 = NBITS + 1, N2 = 2 * N;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float2 *data = reinterpret_cast<float2 *>(smem_raw);
    float2 *tw = data + N2;
    const int tid = threadIdx.x, lane = tid & 31;
    const size_t f = blockIdx.x;
    const int P = n_mics * (n_mics - 1) / 2, pr = blockIdx.y;
    int ma = 0, rem = pr;
    while (rem >= n_mics - 1 - ma) { rem -= n_mics - 1 - ma; ma++; }
    const int mb = ma + 1 + rem;
    for (int k = tid; k < N; k += 256) { float sn, cs; sincospif((float)k / (float)N, &sn, &cs); tw[k] = make_float2(cs, -sn); }
    const float2 *X = spec + (f * n_mics + ma) * (size_t)(N + 1), *Y = spec + (f * n_mics + mb) * (size_t)(N + 1);
    // corr[s] = sum_i x[i] y[i+s]  <->  conj(X) Y ; PHAT: unit magnitude
    for (int k = tid; k <= N; k += 256) {
        const float2 x = X[k], y = Y[k];
        float2 gk = make_float2(x.x * y.x + x.y * y.y, x.x * y.y - x.y * y.x);
        const float mag = sqrtf(gk.x * gk.x + gk.y * gk.y);
        const float inv = mag > 1e-20f ? 1.0f / mag : 0.0f;
        gk.x *= inv; gk.y *= inv;
        data[bitrev<NB2>(k)] = gk;
        if (k > 0 && k < N) data[bitrev<NB2>(N2 - k)] = make_float2(gk.x, -gk.y);
    }
    __syncthreads();
    fft_inplace<NB2>(data, tw, true, tid, 256);
    if (tid < 32) {     // first-max arg-max over s = -L..L (ascending), as correlations.c:20-23 does
        float bv = -INFINITY; int bs = 0x7fffffff;
        for (int li = lane; li < 2 * L + 1; li += 32) {
            const int s = li - L;
            const float v = data[(s + N2) & (N2 - 1)].x;
            if (v > bv) { bv = v; bs = s; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int os = __shfl_xor_sync(0xffffffffu, bs, o);
            if (ov > bv || (ov == bv && os < bs)) { bv = ov; bs = os; }
        }
        if (lane == 0) { lags[f * P + pr] = bs; if (peak) peak[f * P + pr] = bv / (float)N2; }
    }
}

} // namespace atk

using namespace atk;

cudaError_t at_launch_gccphat(int n_mics, int n_bits, int L, const uint8_t *d_adc, const int32_t *d_heads,
                              const int16_t *d_window, size_t n_frames, float2 *d_spec, int32_t *d_lags, float *d_peak,
                              cudaStream_t st)
{
    if (!n_frames) return cudaSuccess;
    const int P = n_mics * (n_mics - 1) / 2;
    const dim3 g1((unsigned)n_frames, (unsigned)((n_mics + 1) / 2)), g2((unsigned)n_frames, (unsigned)P);
    cudaError_t e;
#define AT_GCC(NB)                                                                                              \
    {                                                                                                           \
        const int smem = (int)(sizeof(float2) * ((2 << NB) + (1 << NB)));                                       \
        if ((e = cudaFuncSetAttribute(gcc_forward_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e; \
        if ((e = cudaFuncSetAttribute(gcc_pair_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;    \
        gcc_forward_kernel<NB><<<g1, 256, smem, st>>>(d_adc, d_heads, d_window, n_mics, d_spec);                \
        gcc_pair_kernel<NB><<<g2, 256, smem, st>>>(d_spec, n_mics, L, d_lags, d_peak);                          \
    }
    if (n_bits == 10) AT_GCC(10)
    else if (n_bits == 12) AT_GCC(12)
    else return cudaErrorInvalidValue;
#undef AT_GCC
    at_count_launch(2);
    return cudaGetLastError();
}
