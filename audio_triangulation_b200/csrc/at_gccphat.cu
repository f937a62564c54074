// at_gccphat.cu -- hand-written (cuFFT-free) GCC-PHAT variant of the TDOA stage, for the crossover study of
// BASELINE config 4.  NOT a reference algorithm: the reference correlates directly in integer arithmetic
// (components/correlations.c:9-24); PHAT whitening changes the statistic, so only the arg-max lags can be compared
// (agreement rate), never the curves.  Frame preparation is the same integer path as everywhere else (DC removal,
// <<8, Q15 window), then float32.
//
// Everything is built from one routine: an N-point complex FFT (N = frame length) held by a group of N/8 threads,
// Stockham autosort, radix-8 butterflies in registers (N = 4096: 8.8.8.8; N = 1024: 8.8.4.4), split re/im arrays in
// shared memory with an index padding that keeps the strided writes of the first passes off each other's banks.  A CTA
// runs two such FFTs side by side (the even and the odd output bins of the 2N-point transform of a zero-padded frame,
// or, backwards, the even and odd input bins):
//
//   gcc_forward_kernel   one CTA per (frame, channel couple): z = a + i b of two real channels, zero-padded to 2N:
//                        Z[2k] = FFT_N(z)[k], Z[2k+1] = FFT_N(z w^n)[k], w = exp(-i pi / N); the two spectra are
//                        separated by Hermitian symmetry and WHITENED per channel, U = X / |X| (PHAT of a pair is
//                        conj(U_a) U_b), written to a scratch [frames][mics][N+2] half2 (kept small enough to
//                        stay in L2 until the pair kernel has read it).
//   gcc_pair_kernel      one CTA per (frame, pair couple): G_p + i G_q of two pairs (both inverse transforms are real),
//                        Hermitian extension, y[s] = IFFT_N(G even)[s] + exp(+i pi s / N) IFFT_N(G odd)[s]; only
//                        2L+1 outputs are wanted, so the last pass shrinks to one short sum per wanted lag; first-max
//                        arg-max.
#include <cuda_fp16.h>
#include <math.h>

#include <vector>

#include "at_fused_common.cuh"

namespace atk {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }      // a * (-i)
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// forward DFTs of size 4 and 8 (exp(-2 pi i / R)), natural order in and out
__device__ __forceinline__ void dft4(float2 &v0, float2 &v1, float2 &v2, float2 &v3)
{
    const float2 c0 = cadd(v0, v2), c2 = csub(v0, v2), c1 = cadd(v1, v3), c3 = cmul_mi(csub(v1, v3));
    v0 = cadd(c0, c1); v2 = csub(c0, c1); v1 = cadd(c2, c3); v3 = csub(c2, c3);
}
__device__ __forceinline__ void dft8(float2 (&v)[8])
{
    constexpr float H = 0.70710678118654752440f;
    float2 a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]), a1 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]);
    float2 a2 = cadd(v[2], v[6]), a6 = csub(v[2], v[6]), a3 = cadd(v[3], v[7]), a7 = csub(v[3], v[7]);
    a5 = make_float2(H * (a5.x + a5.y), H * (a5.y - a5.x));            // * (1 - i) / sqrt 2
    a6 = cmul_mi(a6);                                                  // * (-i)
    a7 = make_float2(H * (a7.y - a7.x), -H * (a7.x + a7.y));           // * (-1 - i) / sqrt 2
    dft4(a0, a1, a2, a3);                                              // even outputs 0 2 4 6
    dft4(a4, a5, a6, a7);                                              // odd outputs  1 3 5 7
    v[0] = a0; v[1] = a4; v[2] = a1; v[3] = a5; v[4] = a2; v[5] = a6; v[6] = a3; v[7] = a7;
}

// shared-memory index of element i: one pad word per 32 and eight per 64 -- the radix-8 scatter of the first pass
// (thread t writes 8 t + r) then hits 32 different banks, that of the second (64 (t / 8) + t % 8 + 8 r) at most two
__device__ __forceinline__ int fft_pad(int i) { return i + (i >> 5) + ((i >> 6) << 3); }
template <int N>
struct FftGeo {
    static constexpr int T = N / 8;                                    // threads per FFT group, 8 points each
    static constexpr int WORDS = N + (N >> 5) + ((N >> 6) << 3);       // padded floats per array
    static constexpr int NPASS = 4;
    static_assert(N == 4096 || N == 1024, "pass lists exist for these sizes");
    __host__ __device__ static constexpr int radix(int p) { return N == 4096 ? 8 : (p < 2 ? 8 : 4); }
};

// One Stockham pass over NF independent FFTs held as v[fft][8] by thread t of the group (Govindaraju et al. 2008):
// butterfly j takes in[j + r N / R], multiplies by exp(-2 pi i (j mod Ns) r / (Ns R)), transforms, and puts result r at
// (j / Ns) Ns R + (j mod Ns) + r Ns.  A radix-4 pass treats the thread's 8 values as two butterflies (j = t, t + T).
// tw = exp(-2 pi i n / (2 N)), n < 2 N.
template <int N, int R>
__device__ __forceinline__ void fft_twiddle_dft(float2 (&v)[8], int t, int Ns, const float2 *__restrict__ tw)
{
    constexpr int T = N / 8;
    if constexpr (R == 8) {
        const int k = t & (Ns - 1);
        if (Ns > 1) {
            const float2 w1 = __ldg(&tw[2 * k * (N / (Ns * 8))]);      // Ns is a power of two: the division is a shift
            float2 w = w1;
#pragma unroll
            for (int r = 1; r < 8; r++) { v[r] = cmul(v[r], w); if (r < 7) w = cmul(w, w1); }
        }
        dft8(v);
    } else {
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const int j = t + b * T, k = j & (Ns - 1);
            const float2 w1 = __ldg(&tw[2 * k * (N / (Ns * 4))]);
            const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
            v[4 * b + 1] = cmul(v[4 * b + 1], w1); v[4 * b + 2] = cmul(v[4 * b + 2], w2); v[4 * b + 3] = cmul(v[4 * b + 3], w3);
            dft4(v[4 * b], v[4 * b + 1], v[4 * b + 2], v[4 * b + 3]);
        }
    }
}
// pad(a + c) = pad(a) + pad(c) when c is a multiple of 64 (no carry into the pad terms): the element addresses of a pass
// are one fft_pad per thread plus compile-time offsets
__host__ __device__ constexpr int fft_pad_c(int c) { return c + (c >> 5) + ((c >> 6) << 3); }
template <int N, int R>
__device__ __forceinline__ void fft_store(const float2 (&v)[8], int t, int Ns, float *re, float *im)
{
    constexpr int T = N / 8;
    if constexpr (R == 8) {
        const int k = t & (Ns - 1), j0 = (t - k) * 8 + k, a0 = fft_pad(j0);
#pragma unroll
        for (int r = 0; r < 8; r++) {
            // Ns = 1: j0 = 8 t, + r stays inside the 32-word block.  Ns = 8: j0 = 64 (t / 8) + t % 8, + 8 r stays inside the
            // 64-word block and crosses its 32-word half at r = 4.  Ns >= 64: multiples of 64.
            const int a = Ns == 1 ? a0 + r : (Ns == 8 ? a0 + 8 * r + (r >= 4 ? 1 : 0) : a0 + r * fft_pad_c(Ns));
            re[a] = v[r].x; im[a] = v[r].y;
        }
    } else {
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const int j = t + b * T, k = j & (Ns - 1), j0 = (j - k) * 4 + k, a0 = fft_pad(j0);      // radix-4 passes have Ns >= 64
#pragma unroll
            for (int r = 0; r < 4; r++) { const int a = a0 + r * fft_pad_c(Ns); re[a] = v[4 * b + r].x; im[a] = v[4 * b + r].y; }
        }
    }
}
template <int N, int R>
__device__ __forceinline__ void fft_load(float2 (&v)[8], int t, const float *re, const float *im)
{
    constexpr int T = N / 8;
    static_assert(T % 64 == 0, "offsets of a thread's elements are multiples of 64");
    const int a0 = fft_pad(t);
    if constexpr (R == 8) {
#pragma unroll
        for (int r = 0; r < 8; r++) { const int a = a0 + fft_pad_c(r * T); v[r] = make_float2(re[a], im[a]); }
    } else {
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
            for (int r = 0; r < 4; r++) { const int a = a0 + fft_pad_c(b * T + r * 2 * T); v[4 * b + r] = make_float2(re[a], im[a]); }
    }
}

// Two FFTs at once.  In: v0 / v1 hold element t + r T of the two inputs (the layout the first pass wants).  Out: the
// last pass's results stay in registers: radix 8: v[r] = X[t + r T]; radix 4: v[4 b + r] = X[t + b T + 2 r T].
// re / im: [2][WORDS] each.  All T threads of the group call it (the group is the CTA: __syncthreads).
// LAST = false stops after the last-but-one pass has been stored (and synchronised): the caller finishes by itself.
template <int N, bool LAST = true>
__device__ __forceinline__ void fft2_run(float2 (&v0)[8], float2 (&v1)[8], int t, float *re, float *im, const float2 *__restrict__ tw)
{
    using G = FftGeo<N>;
    int Ns = 1;
#define AT_FFT_PASS(P)                                                                                          \
    {                                                                                                           \
        constexpr int R = G::radix(P);                                                                          \
        if (P > 0) {                                                                                            \
            fft_load<N, R>(v0, t, re, im); fft_load<N, R>(v1, t, re + G::WORDS, im + G::WORDS);                 \
            __syncthreads();                                                                                    \
        }                                                                                                       \
        fft_twiddle_dft<N, R>(v0, t, Ns, tw); fft_twiddle_dft<N, R>(v1, t, Ns, tw);                             \
        if (P < G::NPASS - 1) {                                                                                 \
            fft_store<N, R>(v0, t, Ns, re, im); fft_store<N, R>(v1, t, Ns, re + G::WORDS, im + G::WORDS);       \
            __syncthreads();                                                                                    \
        }                                                                                                       \
        Ns *= R;                                                                                                \
    }
    AT_FFT_PASS(0) AT_FFT_PASS(1) AT_FFT_PASS(2)
    if constexpr (LAST) AT_FFT_PASS(3)
#undef AT_FFT_PASS
}
// One output of the last pass only: X[i] for i = j + rout Ns (Ns = N / R, butterfly j < Ns, rout = 0 or R - 1):
// sum_r in[j + r Ns] (w rho)^r with w = exp(-2 pi i j / N) and rho = exp(-2 pi i rout / R).
template <int N>
__device__ __forceinline__ float2 fft_last_one(int j, bool top, const float *re, const float *im, const float2 *__restrict__ tw)
{
    constexpr int R = FftGeo<N>::radix(3), Ns = N / R;
    float2 b = __ldg(&tw[2 * j]);
    if (top) b = R == 8 ? cmul(b, make_float2(0.70710678118654752440f, 0.70710678118654752440f)) : make_float2(-b.y, b.x);   // * exp(+2 pi i / R)
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int r = R - 1; r >= 0; r--) {       // Horner
        const int a = fft_pad(j + r * Ns);
        acc = cadd(cmul(acc, b), make_float2(re[a], im[a]));
    }
    return acc;
}
// index of the element a thread holds in register slot q after fft2_run
template <int N>
__device__ __forceinline__ int fft_out_index(int t, int q)
{
    constexpr int T = N / 8;
    return FftGeo<N>::radix(3) == 8 ? t + q * T : t + (q >> 2) * T + (q & 3) * 2 * T;
}

template <int NBITS>
__global__ void __launch_bounds__((1 << NBITS) / 8, 8192 >> NBITS) gcc_forward_kernel(const uint8_t *adc, const int32_t *heads, const int16_t *window,
                                                                       int n_mics, const float2 *__restrict__ tw, __half2 *spec /*[F][M][N+2], or tiled*/,
                                                                                      int fg_log2, __half2 *nyq)
{
    constexpr int N = 1 << NBITS, N2 = 2 * N, T = N / 8;
    using G = FftGeo<N>;
    extern __shared__ __align__(16) float fft_smem[];
    float *re = fft_smem, *im = fft_smem + 2 * G::WORDS;
    __shared__ int red[2][32];
    __shared__ int mean_s[2];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const size_t f = blockIdx.y;          // the CTAs of one frame are neighbours: its spectra stay in L2 for the pair kernel
    const int ca = 2 * blockIdx.x, cb = ca + 1 < n_mics ? ca + 1 : -1;
    const int head = heads ? (heads[f] & (N - 1)) : 0;
    const uint8_t *pa = adc + (f * n_mics + ca) * N, *pb = cb >= 0 ? adc + (f * n_mics + cb) * N : nullptr;
    // this thread's 8 chronological samples n = t + r T of both channels; channel sums -> floor mean (rolling_buffer.c:48-64)
    int ra[8], rb[8], sa = 0, sb = 0;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int j = (head + t + r * T) & (N - 1);
        ra[r] = pa[j]; rb[r] = pb ? pb[j] : 0;
        sa += ra[r]; sb += rb[r];
    }
    for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
    if (lane == 0) { red[0][warp] = sa; red[1][warp] = sb; }
    __syncthreads();
    if (t < 2) { int s = 0; for (int w = 0; w < T / 32; w++) s += red[t][w]; mean_s[t] = (int)(short)(s >> NBITS); }
    __syncthreads();
    // prepared samples (integer path of at_fused_common.cuh) as floats: z = a + i b, and z w^n for the odd bins
    float2 v0[8], v1[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int n = t + r * T, w = window[n];
        v0[r] = make_float2((float)prep_sample(ra[r], mean_s[0], w), pb ? (float)prep_sample(rb[r], mean_s[1], w) : 0.f);
        v1[r] = cmul(v0[r], __ldg(&tw[n]));
    }
    fft2_run<N>(v0, v1, t, re, im, tw);
    // E[k] = Z[2k] and O[k] = Z[2k+1] to shared memory (natural order), then the separation needs Z[m] and Z[2N - m]
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int a = fft_pad(fft_out_index<N>(t, q));
        re[a] = v0[q].x; im[a] = v0[q].y; re[G::WORDS + a] = v1[q].x; im[G::WORDS + a] = v1[q].y;
    }
    __syncthreads();
    // Z = A + i B with A, B the spectra of the two real channels: A[m] = (Z[m] + conj Z[-m]) / 2, B[m] = (Z[m] - conj Z[-m]) / 2i;
    // whitened (PHAT): U = X / |X|
    // fg_log2 < 0: [frame][mic][N + 2] (pair kernel).  Else, for the tensor-core inverse (at_gccphat_dft.cu): one contiguous tile
    // per (group of 2^fg_log2 frames, chunk of 32 bins): [group][chunk][frame in group][mic][36]; bin N goes to nyq[frame][mic]
    __half2 *ua = spec + (f * n_mics + ca) * (size_t)(N + 2), *ub = cb >= 0 ? spec + (f * n_mics + cb) * (size_t)(N + 2) : nullptr;
    // (rows of 36 half2, 32 used: 144-byte rows keep the converter threads' 16-byte reads of different mics off each other's banks)
    const size_t tile = ((size_t)n_mics << (fg_log2 < 0 ? 0 : fg_log2)) * 36;                       // half2 per chunk tile
    const size_t tbase = fg_log2 < 0 ? 0 : (f >> fg_log2) * (size_t)(N / 32) * tile + (f & ((1u << fg_log2) - 1)) * (size_t)n_mics * 36;
    for (int m = t; m <= N; m += T) {
        const int odd = m & 1, k = m >> 1, kc = odd ? N - 1 - k : (N - k) & (N - 1);     // Z[2N - m]: same parity
        const int a = odd * G::WORDS + fft_pad(k), ac = odd * G::WORDS + fft_pad(kc);
        const float2 z = make_float2(re[a], im[a]), zc = make_float2(re[ac], im[ac]);
        float2 A = make_float2(0.5f * (z.x + zc.x), 0.5f * (z.y - zc.y)), B = make_float2(0.5f * (z.y + zc.y), -0.5f * (z.x - zc.x));
        const float ma = A.x * A.x + A.y * A.y, mb = B.x * B.x + B.y * B.y;
        const float ia = ma > 1e-30f ? rsqrtf(ma) : 0.f, ib = mb > 1e-30f ? rsqrtf(mb) : 0.f;
        const __half2 ha = __floats2half2_rn(A.x * ia, A.y * ia), hb = __floats2half2_rn(B.x * ib, B.y * ib);   // unit magnitude: half precision costs 2^-11 of phase
        if (fg_log2 < 0) {
            ua[m] = ha;
            if (ub) ub[m] = hb;
        } else if (m < N) {
            __half2 *d = spec + tbase + (size_t)(m >> 5) * tile + (m & 31);
            d[(size_t)ca * 36] = ha;
            if (cb >= 0) d[(size_t)cb * 36] = hb;
        } else {
            nyq[f * n_mics + ca] = ha;
            if (cb >= 0) nyq[f * n_mics + cb] = hb;
        }
    }
    (void)N2;
}

// pair index -> (first, second) microphone, pairs in the order (0,1) (0,2) ... (M-2,M-1)
__device__ __forceinline__ void pair_mics(int pr, int n_mics, int &ma, int &mb)
{
    ma = 0;
    int rem = pr;
    while (rem >= n_mics - 1 - ma) { rem -= n_mics - 1 - ma; ma++; }
    mb = ma + 1 + rem;
}

template <int NBITS>
__global__ void __launch_bounds__((1 << NBITS) / 8, 8192 >> NBITS) gcc_pair_kernel(const __half2 *__restrict__ spec, int n_mics, int L, const float2 *__restrict__ tw,
                                                                                   int32_t *lags, float *peak)
{
    constexpr int N = 1 << NBITS, N2 = 2 * N, T = N / 8;
    using G = FftGeo<N>;
    extern __shared__ __align__(16) float fft_smem[];
    float *re = fft_smem, *im = fft_smem + 2 * G::WORDS;
    __shared__ float ys[2][256];
    const int t = threadIdx.x, lane = t & 31;
    const size_t f = blockIdx.y;
    const int P = n_mics * (n_mics - 1) / 2, p0 = 2 * blockIdx.x, p1 = p0 + 1 < P ? p0 + 1 : -1;
    int a0, b0, a1 = 0, b1 = 0;
    pair_mics(p0, n_mics, a0, b0);
    if (p1 >= 0) pair_mics(p1, n_mics, a1, b1);
    const __half2 *S = spec + f * n_mics * (size_t)(N + 2);
    const __half2 *Xa = S + a0 * (size_t)(N + 2), *Xb = S + b0 * (size_t)(N + 2), *Xc = S + a1 * (size_t)(N + 2), *Xd = S + b1 * (size_t)(N + 2);
    // corr[s] = sum_i x[i] y[i+s]  <->  G = conj(U_x) U_y (unit magnitude), Hermitian in m; the two pairs ride on the real
    // and the imaginary part of one transform: Z = G_p + i G_q.  The inverse transform is conj(FFT(conj Z)).
    auto zc = [&](float2 xa, float2 xb, float2 xc, float2 xd, bool mir) {      // conj(Z[m]) from the four whitened bins
        float2 gp = cmul(cconj(xa), xb), gq = p1 >= 0 ? cmul(cconj(xc), xd) : make_float2(0.f, 0.f);
        if (mir) { gp.y = -gp.y; gq.y = -gq.y; }
        return make_float2(gp.x - gq.y, -(gp.y + gq.x));                        // conj(gp + i gq)
    };
    auto zin = [&](int m) {            // m < 2N
        const bool mir = m > N;
        const int mm = mir ? N2 - m : m;
        return zc(__half22float2(__ldg(&Xa[mm])), __half22float2(__ldg(&Xb[mm])), __half22float2(__ldg(&Xc[mm])), __half22float2(__ldg(&Xd[mm])), mir);
    };
    auto ld2 = [](const __half2 *p, float2 &lo, float2 &hi) {                   // bins m, m + 1 (m even): one 8-byte load
        const uint2 u = __ldg(reinterpret_cast<const uint2 *>(p));
        lo = __half22float2(*reinterpret_cast<const __half2 *>(&u.x)); hi = __half22float2(*reinterpret_cast<const __half2 *>(&u.y));
    };
    float2 v0[8], v1[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int k = t + r * T;
        if (r < 4) {      // bins 2k, 2k + 1 <= N: not mirrored
            float2 xa0, xa1, xb0, xb1, xc0, xc1, xd0, xd1;
            ld2(Xa + 2 * k, xa0, xa1); ld2(Xb + 2 * k, xb0, xb1); ld2(Xc + 2 * k, xc0, xc1); ld2(Xd + 2 * k, xd0, xd1);
            v0[r] = zc(xa0, xb0, xc0, xd0, false); v1[r] = zc(xa1, xb1, xc1, xd1, false);
        } else {
            v0[r] = zin(2 * k); v1[r] = zin(2 * k + 1);
        }
    }
    // y[s] = conj(E[s mod N]) + exp(+i pi s / N) conj(O[s mod N]); wanted: s = -L..L, i.e. i = s mod N in [0, L] or [N-L, N)
    auto put = [&](int i, float2 e, float2 o) {
        const int s = i <= L ? i : i - N;
        const float2 w = __ldg(&tw[(s + N2) & (N2 - 1)]);                    // exp(-i pi s / N); its conjugate is wanted
        const float2 y = cadd(cconj(e), cmul(cconj(w), cconj(o)));
        ys[0][s + L] = y.x; ys[1][s + L] = y.y;
    };
    constexpr int RL = G::radix(3), NsL = N / RL;                            // last pass: butterflies j < NsL, outputs j + r NsL
    // only 2L + 1 of the N outputs are wanted (L < T, checked by the launcher): each is one 8- (4-) term sum of the last pass
    fft2_run<N, false>(v0, v1, t, re, im, tw);
    const int jlo = t, jhi = NsL - T + t;                                     // outputs jlo (r = 0) and jhi + (RL - 1) NsL
    if (jlo <= L) put(jlo, fft_last_one<N>(jlo, false, re, im, tw), fft_last_one<N>(jlo, false, re + G::WORDS, im + G::WORDS, tw));
    if (jhi + (RL - 1) * NsL >= N - L)
        put(jhi + (RL - 1) * NsL, fft_last_one<N>(jhi, true, re, im, tw), fft_last_one<N>(jhi, true, re + G::WORDS, im + G::WORDS, tw));
    __syncthreads();
    if (t < 64) {       // first-max arg-max over s = -L..L (ascending), as correlations.c:20-23 does; warp 0: pair p0, warp 1: p1
        const int pr = t < 32 ? p0 : p1;
        const float *y = ys[t < 32 ? 0 : 1];
        float bv = -INFINITY; int bs = 0x7fffffff;
        if (pr >= 0)
            for (int li = lane; li < 2 * L + 1; li += 32) {
                const float v = y[li];
                if (v > bv) { bv = v; bs = li - L; }
            }
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int os = __shfl_xor_sync(0xffffffffu, bs, o);
            if (ov > bv || (ov == bv && os < bs)) { bv = ov; bs = os; }
        }
        if (lane == 0 && pr >= 0) { lags[f * P + pr] = bs; if (peak) peak[f * P + pr] = bv / (float)N2; }
    }
}

} // namespace atk

using namespace atk;

// exp(-2 pi i n / (2 N)), n < 2 N, evaluated in double on the host
void at_gccphat_twiddles(int n_bits, float2 *h_tw)
{
    const int N2 = 2 << n_bits;
    for (int n = 0; n < N2; n++) {
        const double a = -2.0 * M_PI * (double)n / (double)N2;
        h_tw[n] = make_float2((float)cos(a), (float)sin(a));
    }
}

// fg_log2 < 0: forward + pair kernels (inverse FFTs).  fg_log2 >= 0: forward kernel only, spectra tiled for the tensor-core
// inverse (at_launch_gccphat_dft), d_nyq = bin N of every (frame, mic).
cudaError_t at_launch_gccphat(int n_mics, int n_bits, int L, const uint8_t *d_adc, const int32_t *d_heads,
                              const int16_t *d_window, size_t n_frames, const float2 *d_tw, void *d_spec, int fg_log2, void *d_nyq,
                              int32_t *d_lags, float *d_peak, cudaStream_t st)
{
    if (!n_frames) return cudaSuccess;
    const int P = n_mics * (n_mics - 1) / 2;
    if (n_frames > 65535) return cudaErrorInvalidValue;      // the caller chunks (at_gccphat_device)
    const dim3 g1((unsigned)((n_mics + 1) / 2), (unsigned)n_frames), g2((unsigned)((P + 1) / 2), (unsigned)n_frames);
    cudaError_t e;
#define AT_GCC(NB)                                                                                              \
    {                                                                                                           \
        const int smem = (int)(sizeof(float) * 4 * FftGeo<(1 << NB)>::WORDS);                                   \
        if (L >= (1 << NB) / 8) return cudaErrorInvalidValue;      /* lag window vs the pruned last pass */                       \
        if ((e = cudaFuncSetAttribute(gcc_forward_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e; \
        if ((e = cudaFuncSetAttribute(gcc_pair_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;    \
        gcc_forward_kernel<NB><<<g1, (1 << NB) / 8, smem, st>>>(d_adc, d_heads, d_window, n_mics, d_tw, (__half2 *)d_spec, fg_log2, (__half2 *)d_nyq); \
        if (fg_log2 < 0) gcc_pair_kernel<NB><<<g2, (1 << NB) / 8, smem, st>>>((const __half2 *)d_spec, n_mics, L, d_tw, d_lags, d_peak); \
    }
    if (n_bits == 10) AT_GCC(10)
    else if (n_bits == 12) AT_GCC(12)
    else return cudaErrorInvalidValue;
#undef AT_GCC
    at_count_launch(fg_log2 < 0 ? 2 : 1);
    return cudaGetLastError();
}
