// at_fused_imma_cta.cu -- tensor-core localization kernel for general arrays: M microphones (up to 8,
// 28 pairs) and 1024- or 4096-sample frames (BASELINE config 4).  Same exact mathematics as
// at_fused_imma.cu (byte-split Toeplitz x Hankel IMMA tiles, int32 accumulate, int64 recombine), different
// mapping: ONE CTA PER FRAME.  The CTA prepares all channels once into shared byte planes; its warps then
// share them and each takes "pair groups" -- one x microphone against up to three y microphones -- so a
// group needs 2 Toeplitz planes + up to 6 Hankel planes per k-step for up to 12 IMMA, the same density as
// the 3-microphone kernel.  The raw curves land in the block epilogue of at_fused_common.cuh (arg-max,
// Gaussian re-weighting, likelihood map over all P pairs).
//
// No reference counterpart exists for these shapes (the reference is 3 mics x 1024): parity is against the
// generalised oracle (oracle/at_oracle.c), "unpinned" in the sense of DESIGN.md section 2.
#include <limits.h>

#include "at_imma_common.cuh"

namespace atk {

template <int NBITS, int L>
struct CtaGeo {
    static constexpr int N = 1 << NBITS;
    static constexpr int PAD = round_up(L, 8);                  // = Geo::PADL: curve index j = s + PAD
    static constexpr int KSTEPS = ceil_div(N + 8, 32);
    static constexpr int PLANE = round_up(32 * KSTEPS + 8 * 15 + 8, 16);
    static_assert(PAD % 16 == 0, "plane stores are 16-byte aligned");
    static_assert(Geo<NBITS, L>::PADL == PAD && Geo<NBITS, L>::NLAGS_PAD >= 96, "epilogue curve layout");
};

template <int NMICS, int NBITS, int L, int WARPS>
struct CtaSmem {
    using G = CtaGeo<NBITS, L>;
    alignas(16) uint32_t win2[G::N];                       // pre-masked 2*W, chunk-interleaved (see at_fused_imma.cu)
    alignas(16) uint8_t plane[NMICS][2][G::PLANE];         // [channel][hi, lo], zero padded
    EpiSmem<NMICS, NBITS, L> epi;
    float gauss[2 * L + 1];
    int mean[NMICS];
};

__device__ __forceinline__ uint32_t cta_lds32(uint32_t smem_addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_addr));
    return v;
}

// One pair group: x microphone (Toeplitz side) against NY y microphones (Hankel side), all k-steps.
template <int NY, int KSTEPS, int PLANE, int NJ, int PAD>
__device__ __forceinline__ void pair_group(uint32_t planes_s, uint32_t four, int x, int y0, int lane, int pair0,
                                           long long (*curve)[NJ])
{
    const int g = lane >> 2, t = lane & 3;
    const int boff = 8 * t - g + PAD;                      // B: plane index of (k = 8t, column g)
    const uint32_t xb = planes_s + (2 * x) * PLANE + (boff & ~3);
    const int bsh = (boff & 3) * 8;
    const uint32_t ya = planes_s + (2 * y0) * PLANE + 8 * t + 8 * g;
    const uint32_t ya4 = ya + four;                        // "+4" from a kernel parameter: keeps the loads unfused
    int acc[NY][3][4];
#pragma unroll
    for (int a = 0; a < NY; a++)
#pragma unroll
        for (int b = 0; b < 3; b++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[a][b][c] = 0;
#pragma unroll 3
    for (int ks = 0; ks < KSTEPS; ks++) {
        const int k0 = 32 * ks;
        uint32_t Xh[2], Xl[2];
        {
            const uint32_t w0 = cta_lds32(xb + k0), w1 = cta_lds32(xb + k0 + 4), w2 = cta_lds32(xb + k0 + 8);
            Xh[0] = __funnelshift_r(w0, w1, bsh); Xh[1] = __funnelshift_r(w1, w2, bsh);
            const uint32_t v0 = cta_lds32(xb + PLANE + k0), v1 = cta_lds32(xb + PLANE + k0 + 4), v2 = cta_lds32(xb + PLANE + k0 + 8);
            Xl[0] = __funnelshift_r(v0, v1, bsh); Xl[1] = __funnelshift_r(v1, v2, bsh);
        }
#pragma unroll
        for (int a = 0; a < NY; a++) {
            uint32_t Yh[4], Yl[4];
            const uint32_t oh = (2 * a) * PLANE + k0, ol = oh + PLANE;
            Yh[0] = cta_lds32(ya + oh); Yh[2] = cta_lds32(ya4 + oh); Yh[1] = cta_lds32(ya + oh + 64); Yh[3] = cta_lds32(ya4 + oh + 64);
            Yl[0] = cta_lds32(ya + ol); Yl[2] = cta_lds32(ya4 + ol); Yl[1] = cta_lds32(ya + ol + 64); Yl[3] = cta_lds32(ya4 + ol + 64);
            mma_s8_s8(acc[a][0], Yh, Xh); mma_s8_u8(acc[a][1], Yh, Xl);
            mma_u8_u8(acc[a][2], Yl, Xl); mma_u8_s8(acc[a][1], Yl, Xh);
        }
    }
#pragma unroll
    for (int a = 0; a < NY; a++)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int j = 8 * (g + 8 * (i >> 1)) + 2 * t + (i & 1);
            if (j < NJ) curve[pair0 + a][j] = 65536LL * acc[a][0][i] + 256LL * acc[a][1][i] + (long long)acc[a][2][i];
        }
}

template <int NMICS, int NBITS, int L, int WARPS, int CTAS_PER_SM>
__global__ void __launch_bounds__(WARPS * 32, CTAS_PER_SM) at_fused_imma_cta_kernel(const AtFusedParams p)
{
    using G = CtaGeo<NBITS, L>;
    using S = CtaSmem<NMICS, NBITS, L, WARPS>;
    constexpr int N = G::N, PAD = G::PAD, PLANE = G::PLANE, THREADS = WARPS * 32;
    constexpr int NJ = Geo<NBITS, L>::NLAGS_PAD;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    S &s = *reinterpret_cast<S *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < (int)(sizeof(s.plane) / 16); i += THREADS)
        reinterpret_cast<uint4 *>(&s.plane[0][0][0])[i] = make_uint4(0, 0, 0, 0);
    imma_win_fill(s.win2, p.window, N, tid, THREADS);
    for (int i = tid; i < 2 * L + 1; i += THREADS) s.gauss[i] = p.gauss[i];
    __syncthreads();
    const uint32_t planes_s = smem_u32(&s.plane[0][0][0]);

    for (unsigned long long f = blockIdx.x; f < p.n_frames; f += gridDim.x) {
        const uint8_t *src = p.adc + f * (unsigned long long)(NMICS * N);
        const int head = p.heads ? (p.heads[f] & (N - 1)) : 0;
        // ---- channel sums -> floor mean (rolling_buffer.c:48-64): one warp per channel
        for (int ch = warp; ch < NMICS; ch += WARPS) {
            unsigned sum = 0;
            for (int k = lane; k < N / 16; k += 32) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + ch * N) + k);
                sum = __dp4a(v.x, 0x01010101u, sum); sum = __dp4a(v.y, 0x01010101u, sum);
                sum = __dp4a(v.z, 0x01010101u, sum); sum = __dp4a(v.w, 0x01010101u, sum);
            }
            sum = __reduce_add_sync(0xffffffffu, sum);
            if (lane == 0) s.mean[ch] = (int)(sum >> NBITS);
        }
        __syncthreads();
        // ---- DC removal, <<8, window -> hi / lo byte planes; one 16-sample chunk per thread-iteration
        for (int item = tid; item < NMICS * (N / 16); item += THREADS) {
            const int ch = item / (N / 16), j0 = (item % (N / 16)) * 16;
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + ch * N + j0));
            const uint32_t rw[4] = {v.x, v.y, v.z, v.w};
            const int mean = s.mean[ch];
            if ((head & 15) == 0) {
                const int i0 = (j0 - head) & (N - 1);
                uint32_t hi[4], lo[4];
                imma_prep16(rw, mean, s.win2, i0, hi, lo);
                *reinterpret_cast<uint4 *>(&s.plane[ch][0][PAD + i0]) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4 *>(&s.plane[ch][1][PAD + i0]) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            } else {
                for (int e = 0; e < 16; e++) {
                    const int i = (j0 + e - head) & (N - 1);
                    const int pr = imma_prep1(rw[e >> 2] >> (8 * (e & 3)), mean, s.win2, i);
                    s.plane[ch][0][PAD + i] = (uint8_t)(pr >> 16);
                    s.plane[ch][1][PAD + i] = (uint8_t)(pr >> 8);
                }
            }
        }
        if (p.power) {
            for (int ch = warp; ch < NMICS; ch += WARPS) {
                long long acc = 0;
                for (int k = lane; k < N; k += 32) { const int dv = (int)src[ch * N + k] - s.mean[ch]; acc += (long long)dv * dv; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) p.power[f * NMICS + ch] = acc;
            }
        }
        __syncthreads();
        if (p.windowed)
            for (int idx = tid; idx < NMICS * N; idx += THREADS) {
                const int ch = idx / N, i = idx % N;
                p.windowed[f * (unsigned long long)(NMICS * N) + idx] =
                    (int16_t)(((int)(signed char)s.plane[ch][0][PAD + i] << 8) | s.plane[ch][1][PAD + i]);
            }

        // ---- pair groups: x against up to three y's; groups are dealt to warps round-robin
        int gi = 0;
        for (int x = 0; x < NMICS - 1; x++) {
            const int pair_x0 = x * NMICS - x * (x + 1) / 2;            // pair index of (x, x+1)
            for (int y0 = x + 1; y0 < NMICS; y0 += 3, gi++) {
                if (gi % WARPS != warp) continue;
                const int ny = NMICS - y0 < 3 ? NMICS - y0 : 3;
                const int pair0 = pair_x0 + (y0 - x - 1);
                const uint32_t four = (uint32_t)p.opaque_four;
                if (ny == 3) pair_group<3, G::KSTEPS, PLANE, NJ, PAD>(planes_s, four, x, y0, lane, pair0, s.epi.curve);
                else if (ny == 2) pair_group<2, G::KSTEPS, PLANE, NJ, PAD>(planes_s, four, x, y0, lane, pair0, s.epi.curve);
                else pair_group<1, G::KSTEPS, PLANE, NJ, PAD>(planes_s, four, x, y0, lane, pair0, s.epi.curve);
            }
        }
        __syncthreads();
        epilogue<NMICS, NBITS, L, THREADS>(s.epi, s.gauss, p, f);
        __syncthreads();   // curves and planes are rewritten by the next frame
    }
}

template <int NMICS, int NBITS, int L, int WARPS, int CTAS_PER_SM>
static cudaError_t launch_cta(const AtFusedParams &p, int sm_count, cudaStream_t st)
{
    using S = CtaSmem<NMICS, NBITS, L, WARPS>;
    auto kern = at_fused_imma_cta_kernel<NMICS, NBITS, L, WARPS, CTAS_PER_SM>;
    const int smem = (int)sizeof(S);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    unsigned long long grid = (unsigned long long)sm_count * per_sm;
    if (grid > p.n_frames) grid = p.n_frames;
    if (grid == 0) return cudaSuccess;
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(p);
    at_count_launch();
    return cudaGetLastError();
}

} // namespace atk

bool at_fused_imma_cta_supports(const AtShape &sh)
{
    return (sh.n_mics == 8 || sh.n_mics == 4) && (sh.n_bits == 10 || sh.n_bits == 12) && sh.max_shift == 46;
}

cudaError_t at_launch_fused_imma_cta(const AtShape &sh, const AtFusedParams &p, int sm_count, cudaStream_t st)
{
    if (p.sig16) return cudaErrorInvalidValue;
    if (sh.max_shift != 46) return cudaErrorInvalidValue;
    if (sh.n_mics == 8 && sh.n_bits == 12) return atk::launch_cta<8, 12, 46, 4, 2>(p, sm_count, st);
    if (sh.n_mics == 8 && sh.n_bits == 10) return atk::launch_cta<8, 10, 46, 4, 4>(p, sm_count, st);
    if (sh.n_mics == 4 && sh.n_bits == 10) return atk::launch_cta<4, 10, 46, 4, 4>(p, sm_count, st);
    if (sh.n_mics == 4 && sh.n_bits == 12) return atk::launch_cta<4, 12, 46, 4, 2>(p, sm_count, st);
    return cudaErrorInvalidValue;
}
