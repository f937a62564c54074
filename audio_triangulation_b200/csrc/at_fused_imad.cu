// at_fused_imad.cu -- fused localization kernel, integer-pipe form (AT_KERNEL_IMAD).
//
// One persistent CTA per resident slot walks frames f = blockIdx.x, += gridDim.x.  Per frame:
//   stage   uint8 frame, 1-D bulk async copy (TMA engine) into a double-buffered smem slot,
//           next frame prefetched while this one is computed;
//   prep    DC removal, <<8, Q15 window -> int16 rows in smem, zero padded by the lag range;
//   xcorr   direct integer cross-correlation (ref: components/correlations.c:9-18): a warp owns
//           one (pair, block of 8 lags); each lane takes 8-sample chunks strided by 32 lanes
//           (conflict-free 16-byte LDS), keeps 8 int64 accumulators and issues 64 IMAD.WIDE per
//           3 LDS.128; lanes are combined by a transposing shuffle reduction;
//   epilogue arg-max / Gaussian / likelihood arg-max (at_fused_common.cuh).
// Exact: int16 x int16 -> int32 products accumulated in int64, as the reference does.
#include <limits.h>

#include "at_fused_common.cuh"

namespace atk {

template <int NMICS, int NBITS, int L, int THREADS, bool IN_I16>
struct ImadSmem {
    using G = Geo<NBITS, L>;
    static constexpr int N = G::N;
    static constexpr int RAW_BYTES = NMICS * N * (IN_I16 ? 2 : 1);
    alignas(16) uint8_t raw[2][RAW_BYTES];
    alignas(16) int16_t sig[NMICS][G::ROW];
    alignas(16) int16_t win[N];
    EpiSmem<NMICS, NBITS, L> epi;
    float gauss[2 * L + 1];
    int dcs[NMICS];
    alignas(8) uint64_t bar[2];
};

// 8 accumulators spread over 32 lanes -> lanes 0..7 end with the total of accumulator
// ((lane&1)<<2 | (lane&2) | (lane&4)>>2).
__device__ __forceinline__ long long reduce8(long long (&a)[8], int lane)
{
    long long b[4], c[2], d;
    const bool h0 = lane & 1, h1 = lane & 2, h2 = lane & 4;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const long long keep = h0 ? a[j + 4] : a[j], give = h0 ? a[j] : a[j + 4];
        b[j] = keep + __shfl_xor_sync(0xffffffffu, give, 1);
    }
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const long long keep = h1 ? b[j + 2] : b[j], give = h1 ? b[j] : b[j + 2];
        c[j] = keep + __shfl_xor_sync(0xffffffffu, give, 2);
    }
    {
        const long long keep = h2 ? c[1] : c[0], give = h2 ? c[0] : c[1];
        d = keep + __shfl_xor_sync(0xffffffffu, give, 4);
    }
    d += __shfl_xor_sync(0xffffffffu, d, 8);
    d += __shfl_xor_sync(0xffffffffu, d, 16);
    return d;
}

template <int NMICS, int NBITS, int L, int THREADS, bool IN_I16>
__global__ void __launch_bounds__(THREADS, (NBITS <= 10 && NMICS <= 3) ? 2 : 1)
at_fused_imad_kernel(const AtFusedParams p)
{
    using S = ImadSmem<NMICS, NBITS, L, THREADS, IN_I16>;
    using G = Geo<NBITS, L>;
    constexpr int N = G::N, P = NMICS * (NMICS - 1) / 2, NWARPS = THREADS / 32;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    S &s = *reinterpret_cast<S *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // one-time CTA setup: zero the rows (pads stay zero for the CTA's lifetime), tables, barriers
    for (int i = tid; i < NMICS * G::ROW / 2; i += THREADS) reinterpret_cast<uint32_t *>(&s.sig[0][0])[i] = 0u;
    for (int i = tid; i < N; i += THREADS) s.win[i] = p.window[i];
    for (int i = tid; i < 2 * L + 1; i += THREADS) s.gauss[i] = p.gauss[i];
    if (tid == 0) {
        mbar_init(&s.bar[0], 1);
        mbar_init(&s.bar[1], 1);
        fence_barrier_init();
    }
    __syncthreads();

    const uint8_t *src = IN_I16 ? reinterpret_cast<const uint8_t *>(p.sig16) : p.adc;
    unsigned long long f = blockIdx.x;
    if (tid == 0 && f < p.n_frames) {
        mbar_expect_tx(&s.bar[0], S::RAW_BYTES);
        bulk_g2s(s.raw[0], src + f * S::RAW_BYTES, S::RAW_BYTES, &s.bar[0]);
    }
    uint32_t it = 0;
    for (; f < p.n_frames; f += gridDim.x, it++) {
        const int slot = it & 1;
        const unsigned long long fn = f + gridDim.x;
        if (tid == 0 && fn < p.n_frames) {   // prefetch the next frame into the other slot
            fence_proxy_async();             // its previous generic-proxy reads happened before the last barrier
            mbar_expect_tx(&s.bar[slot ^ 1], S::RAW_BYTES);
            bulk_g2s(s.raw[slot ^ 1], src + fn * S::RAW_BYTES, S::RAW_BYTES, &s.bar[slot ^ 1]);
        }
        mbar_wait(&s.bar[slot], (it >> 1) & 1);

        if (IN_I16) {
            const int16_t *r16 = reinterpret_cast<const int16_t *>(s.raw[slot]);
            for (int idx = tid; idx < NMICS * N; idx += THREADS)
                s.sig[idx / N][G::PADL + idx % N] = r16[idx];
            __syncthreads();
        } else {
            const int head = p.heads ? (p.heads[f] & (N - 1)) : 0;
            prep_frame_u8<NMICS, NBITS, L, THREADS>(s.raw[slot], head, s.dcs, s.win, &s.sig[0][0], p, f);
        }

        // ---- direct cross-correlation: unit = (pair, block of 8 lags) per warp ----
        for (int u = warp; u < P * G::NBLK; u += NWARPS) {
            const int pr = u / G::NBLK, blk = u % G::NBLK;
            int ma = 0, rem = pr;                       // pair index -> (ma, mb), lexicographic
            while (rem >= NMICS - 1 - ma) { rem -= NMICS - 1 - ma; ma++; }
            const int mb = ma + 1 + rem;
            const int16_t *X = &s.sig[ma][G::PADL];      // X[i] = a[i]
            const int16_t *Y = &s.sig[mb][8 * blk];      // Y[i + l] = b[i + s], s = 8*blk + l - PADL
            long long acc[8];
#pragma unroll
            for (int l = 0; l < 8; l++) acc[l] = 0;
#pragma unroll 2
            for (int i0 = 8 * lane; i0 < N; i0 += 256) {
                const uint4 xv = *reinterpret_cast<const uint4 *>(X + i0);
                const uint4 ya = *reinterpret_cast<const uint4 *>(Y + i0);
                const uint4 yb = *reinterpret_cast<const uint4 *>(Y + i0 + 8);
                const int x[8] = {sext_lo16(xv.x), sext_hi16(xv.x), sext_lo16(xv.y), sext_hi16(xv.y),
                                  sext_lo16(xv.z), sext_hi16(xv.z), sext_lo16(xv.w), sext_hi16(xv.w)};
                const int y[16] = {sext_lo16(ya.x), sext_hi16(ya.x), sext_lo16(ya.y), sext_hi16(ya.y),
                                   sext_lo16(ya.z), sext_hi16(ya.z), sext_lo16(ya.w), sext_hi16(ya.w),
                                   sext_lo16(yb.x), sext_hi16(yb.x), sext_lo16(yb.y), sext_hi16(yb.y),
                                   sext_lo16(yb.z), sext_hi16(yb.z), sext_lo16(yb.w), sext_hi16(yb.w)};
#pragma unroll
                for (int k = 0; k < 8; k++)
#pragma unroll
                    for (int l = 0; l < 8; l++)   // PTX on purpose: keeps one IMAD.WIDE per MAC (nvcc otherwise
                                                  // lowers 16-bit-ranged products to IMAD + SHF + IADD3 + IADD3.X)
                        asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(acc[l]) : "r"(x[k]), "r"(y[k + l]));
            }
            const long long tot = reduce8(acc, lane);
            if (lane < 8) {
                const int l = ((lane & 1) << 2) | (lane & 2) | ((lane & 4) >> 2);
                s.epi.curve[pr][8 * blk + l] = tot;
            }
        }
        __syncthreads();
        epilogue<NMICS, NBITS, L, THREADS>(s.epi, s.gauss, p, f);
        __syncthreads();   // curve / sig / raw[slot] free for the next iteration
    }
}

template <int NMICS, int NBITS, int L, int THREADS, bool IN_I16>
static cudaError_t launch(const AtFusedParams &p, int sm_count, cudaStream_t st)
{
    using S = ImadSmem<NMICS, NBITS, L, THREADS, IN_I16>;
    auto kern = at_fused_imad_kernel<NMICS, NBITS, L, THREADS, IN_I16>;
    const int smem = (int)sizeof(S);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    unsigned long long grid = (unsigned long long)sm_count * per_sm;
    if (grid > p.n_frames) grid = p.n_frames;
    if (grid == 0) return cudaSuccess;
    kern<<<(unsigned)grid, THREADS, smem, st>>>(p);
    at_count_launch();
    return cudaGetLastError();
}

} // namespace atk

bool at_fused_imad_supports(const AtShape &sh)
{   // keep in step with the AT_CASE list below
    const int m = sh.n_mics, nb = sh.n_bits, l = sh.max_shift;
    return (m == 3 && nb == 10 && (l == 46 || l == 44)) || (m == 2 && nb == 10 && l == 46) || (m == 8 && (nb == 12 || nb == 10) && l == 46) ||
           (m == 3 && nb == 12 && l == 46) || (m == 4 && nb == 10 && l == 46);
}

cudaError_t at_launch_fused_imad(const AtShape &sh, const AtFusedParams &p, int sm_count, cudaStream_t st)
{
    const bool i16 = p.sig16 != nullptr;
#define AT_CASE(M, NB, LL, T)                                                            \
    if (sh.n_mics == M && sh.n_bits == NB && sh.max_shift == LL)                         \
        return i16 ? atk::launch<M, NB, LL, T, true>(p, sm_count, st)                    \
                   : atk::launch<M, NB, LL, T, false>(p, sm_count, st);
    AT_CASE(3, 10, 46, 384)   // reference shape: 36 units over 12 warps
    AT_CASE(2, 10, 46, 384)   // one pair (drop-in correlations_init)
    AT_CASE(8, 12, 46, 512)   // BASELINE config 4: 8 mics x 4096 samples, 28 pairs
    AT_CASE(8, 10, 46, 512)
    AT_CASE(3, 12, 46, 384)
    AT_CASE(3, 10, 44, 384)   // 48 kHz rule: 48000*32/34300 = 44
    AT_CASE(4, 10, 46, 384)
#undef AT_CASE
    return cudaErrorInvalidValue;
}
