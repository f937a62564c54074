// at_fused_umma_m.cu -- tcgen05 (UMMA) localization kernel for 8-microphone arrays (28 pairs), 1024- or 4096-sample
// frames (BASELINE config 4).  Same polyphase Hankel formulation and digit arithmetic as at_fused_umma.cu:
//     D[m][phi] = sum_q Y[m + 16 q] * x[phi + 16 q],   corr[s] = sum_phi D[s + PAD + phi][phi],
// A = one digit plane of the y microphone read as an overlapping MN-major Hankel operand, B = the digit planes of ALL x
// microphones below y side by side.  With 8 microphones B is up to 14 planes wide (N = 224): here the tensor core works
// near its array rate (an M128 x N x K32 int8 MMA costs max(~63, N / 2) cycles, tools/probes/umma_probe.cu), which
// the 3-microphone problem (N = 64) cannot reach -- see DESIGN.md 4.5.
//
// No reference counterpart exists for these shapes: parity is against the generalised oracle (oracle/at_oracle.c),
// "unpinned" in the sense of DESIGN.md section 2, and against the mma.sync kernel (at_fused_imma_cta.cu).
//
// Work per frame: for y = 1..7 and each digit d of y, the product  (y.d plane) x [x0.h x0.l ... x(y-1).h x(y-1).l]  over
// all K-steps fills 32 y TMEM columns.  The 14 (y, d) groups are packed into 7 passes of exactly 256 columns
// ({7h,1h} {7l,1l} {6h,2h} {6l,2l} {5h,3h} {5l,3l} {4h,4l}), TMEM holds two passes, so the tensor core fills one slot
// while the CUDA cores drain the other.  CTA = 25 warps, one CTA per SM, persistent, warp-specialised:
//   warp 0       one elected lane issues the MMAs (frames and passes in order) and commits them to mbarriers;
//   warps 1-8    prep, one channel each: loads, DC removal, <<8, window, balanced digit planes to shared memory
//                (double-buffered frames);
//   warps 9-24   epilogue (512 threads, four warps per TMEM lane quadrant): the 16 tiles of a pass are 8 couples
//                [y.d * x.h | y.d * x.l]; tcgen05.ld 32x32b, couple folded to 256 * (y.d x.h) + (y.d x.l) (fits
//                int32), transposing scatter through shared memory, 16- (or 8-) term diagonal sums in int64 added with the
//                weight of y's digit (2^8 or 1) into the curves; after the last pass the block epilogue of
//                at_fused_common.cuh (arg-max, Gaussian re-weighting, outputs).  The scatter is what bounds the
//                kernel: every accumulator entry crosses shared memory once (write + read at 128 B/clk).
#include <limits.h>
#include <stdlib.h>

#include "at_imma_common.cuh"
#include "at_umma_common.cuh"

namespace atk {

template <int NBITS, int L>
struct UmmaMGeo {
    static constexpr int NM = 8, P = 28;
    static constexpr int N = 1 << NBITS;
    static constexpr int PAD = 48;                       // lag index j = s + PAD; also the left zero pad of a plane
    static constexpr int PLANE = N + 128;                // 48 zeros, N samples, 80 zeros
    static constexpr int KSTEPS = N / 512;               // one MMA (K = 32 rows of 16 bytes) covers 512 samples
    // 1024-sample frames: every plane is stored twice, the second copy advanced by 8 bytes, and both copies are
    // accumulated into the same tiles (D2[m][phi] = D[m][phi] + D[m+8][phi+8], phi < 8, as in at_fused_umma.cu): the MMA
    // time doubles, where it is small, and the diagonal sums -- the bound -- halve.  4096-sample frames keep one copy.
    static constexpr int COPIES = NBITS <= 10 ? 2 : 1;
    static constexpr int PH = 16 / COPIES;               // phases per tile the epilogue has to add up
    static constexpr int FRAME = COPIES * 2 * NM * PLANE;    // planes [copy][channel][h, l]
    static constexpr int NJ = 96;
    static constexpr int ZP = 112;                       // words per (tile, phase) column of the transposing scratch:
                                                         // index = lag index + PH - 1, only lag indices < 96 are kept
    static constexpr int TCOLS = 256;                    // TMEM columns per pass
    static constexpr int PASSES = 7;
    static constexpr int EPI_WARPS = 16, EPI_THREADS = 32 * EPI_WARPS;
    static constexpr int THREADS = 32 * (1 + NM) + EPI_THREADS;   // 800
    static_assert(Geo<NBITS, L>::PADL == PAD && Geo<NBITS, L>::NLAGS_PAD == NJ, "curve layout of the block epilogue");
    static_assert(PAD + L + 15 < 128, "lag window must fit the 128-row tile");
    static_assert(127 + 16 * (N / 16 - 1) + 15 < PLANE, "A operand reads stay inside a plane buffer");
};

template <int NBITS, int L>
struct UmmaMSmem {
    using G = UmmaMGeo<NBITS, L>;
    alignas(128) uint8_t planes[2][G::FRAME];
    alignas(16) int z[8][G::PH][G::ZP];                  // [couple of the pass][phase][row - phase + PH - 1]
    alignas(16) EpiSmem<G::NM, NBITS, L> epi;            // epi.curve accumulates the weighted diagonal sums of a frame
    alignas(16) uint32_t winp[G::N / 2];               // packed window table (umma_prep16p)
    uint16_t couple_tab[G::PASSES][8];                   // (pair << 8) | weight shift (8: y.h, 0: y.l) of each couple
    float gauss[2 * L + 1];
    alignas(8) uint64_t full[2], empty[2], ready[2], sfree[2];
    uint32_t tmem_base;
};

// pass g holds two (y, digit) groups: group 0 at column 0, group 1 at column 32 * y0
__device__ __forceinline__ constexpr int pass_y(int g, int k) { return k == 0 ? 7 - (g >> 1) : (g == 6 ? 4 : 1 + (g >> 1)); }
__device__ __forceinline__ constexpr int pass_d(int g, int k) { return g == 6 ? k : (g & 1); }

template <int NBITS, int L>
__global__ void __launch_bounds__(UmmaMGeo<NBITS, L>::THREADS, 1) at_fused_umma_m_kernel(const AtFusedParams p)
{
    using G = UmmaMGeo<NBITS, L>;
    using S = UmmaMSmem<NBITS, L>;
    constexpr int N = G::N, PAD = G::PAD, PLANE = G::PLANE, NM = G::NM, NJ = G::NJ;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    S &s = *reinterpret_cast<S *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- one-time CTA set-up
    for (int i = tid; i < (int)(sizeof(s.planes) / 16); i += G::THREADS)
        reinterpret_cast<uint4 *>(&s.planes[0][0])[i] = make_uint4(0, 0, 0, 0);
    umma_win_fill_packed(s.winp, p.window, N, tid, G::THREADS);
    for (int i = tid; i < 2 * L + 1; i += G::THREADS) s.gauss[i] = p.gauss[i];
    if (tid < G::PASSES * 8) {      // couple = 32 columns [y.d * x.h | y.d * x.l]: (y, d) of its group, x channel in order
        const int g = tid >> 3, c = tid & 7, y0 = pass_y(g, 0);
        const int grp = c >= y0 ? 1 : 0, x = grp ? c - y0 : c, y = pass_y(g, grp), d = pass_d(g, grp);
        const int pr = x * NM - x * (x + 1) / 2 + (y - x - 1);            // pair (x, y), x < y
        s.couple_tab[g][c] = (uint16_t)((pr << 8) | (d == 0 ? 8 : 0));
    }
    if (tid == 0) {
        for (int k = 0; k < 2; k++) {
            mbar_init(&s.full[k], 1); mbar_init(&s.empty[k], G::EPI_WARPS);
            mbar_init(&s.ready[k], NM); mbar_init(&s.sfree[k], 1);
        }
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s.tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    const unsigned long long nf = p.n_frames, gstride = gridDim.x;
    // frames of this CTA: f = blockIdx.x + gridDim.x * i; frame i uses plane buffer i & 1; its pass g is pass number
    // 7 i + g of the CTA and uses TMEM slot (7 i + g) & 1.  Every mbarrier is waited on in phase order.

    if (warp == 0) {
        // =================================================================== MMA issue
        constexpr uint32_t LBO = (128u >> 4) << 16;
        constexpr uint32_t HI_A = (16u >> 4) | 0x4000u, HI_B = ((uint32_t)PLANE >> 4) | 0x4000u;
        for (unsigned long long i = 0;; i++) {
            if (blockIdx.x + gstride * i >= nf) break;
            const unsigned b = (unsigned)(i & 1);
            mbar_wait(&s.ready[b], (unsigned)(i >> 1) & 1);
            const uint32_t b16 = (smem_u32(&s.planes[b][0]) >> 4) + LBO;          // start-address field of plane 0 (x0.h)
#pragma unroll 1
            for (int g = 0; g < G::PASSES; g++) {
                const unsigned long long gp = 7 * i + g;
                const unsigned slot = (unsigned)(gp & 1), u = (unsigned)(gp >> 1);
                if (u >= 1) mbar_wait(&s.empty[slot], (u - 1) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t cb = tmem + slot * G::TCOLS;
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        const int y = pass_y(g, k), d = pass_d(g, k);
                        const uint32_t col = k == 0 ? 0u : 32u * (uint32_t)pass_y(g, 0);
                        const uint32_t idesc = umma_idesc(32 * y);
#pragma unroll
                        for (int copy = 0; copy < G::COPIES; copy++) {
                            const uint32_t c0 = b16 + (uint32_t)(copy * 2 * NM * (PLANE >> 4));           // plane x0.h of this copy
                            const uint32_t a0 = c0 + (uint32_t)((2 * y + d) * (PLANE >> 4)), x0 = c0 + (PAD >> 4);
#pragma unroll
                            for (int kk = 0; kk < G::KSTEPS; kk++)
                                umma_i8_lohi(cb + col, a0 + 32 * kk, HI_A, x0 + 32 * kk, HI_B, idesc, (kk | copy) ? 1u : 0u);
                        }
                    }
                    umma_commit(&s.full[slot]);
                    if (g == G::PASSES - 1) umma_commit(&s.sfree[b]);
                }
                __syncwarp();
            }
        }
    } else if (warp <= NM) {
        // =================================================================== prep warps, one channel each
        const int ch = warp - 1;
        constexpr int Q = N / 512;
        for (unsigned long long i = 0;; i++) {
            const unsigned long long f = blockIdx.x + gstride * i;
            if (f >= nf) break;
            const unsigned b = (unsigned)(i & 1), v = (unsigned)(i >> 1);
            uint8_t *const ph = &s.planes[b][(2 * ch) * PLANE], *const pl = ph + PLANE;
            const uint8_t *src = p.adc + (f * NM + ch) * (unsigned long long)N;
            const int head = p.heads ? (p.heads[f] & (N - 1)) : 0;
            const int uhead = __shfl_sync(0xffffffffu, head, 0);      // tells the compiler what it cannot see: warp-uniform
            uint4 raw[Q];
            unsigned sum = 0;
#pragma unroll
            for (int q = 0; q < Q; q++) {
                // the lane's chronological samples [512 q + 16 l, +16), un-rotated from the ring: one aligned 16-byte load, or two
                // and a byte shift when the head is not 16-aligned (warp-uniform branch)
                const int r = (q * 512 + lane * 16 + uhead) & (N - 1), r0 = r & ~15, sh = r & 15;
                uint4 x = ldg_stream(src + r0);
                if (sh) x = realign16(x, ldg_stream(src + ((r0 + 16) & (N - 1))), sh);
                raw[q] = x;
                sum = __dp4a(x.x, 0x01010101u, sum); sum = __dp4a(x.y, 0x01010101u, sum);
                sum = __dp4a(x.z, 0x01010101u, sum); sum = __dp4a(x.w, 0x01010101u, sum);
            }
            sum = __reduce_add_sync(0xffffffffu, sum);
            const int mean = (int)(sum >> NBITS);                       // rolling_buffer.c:48-64
            {   // this channel of the next frame -> L1/L2
                const unsigned long long fn = f + gstride;
                if (fn < nf && lane * 128 < N) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.adc + (fn * NM + ch) * (unsigned long long)N + lane * 128));
            }
            if (v >= 1) mbar_wait(&s.sfree[b], (v - 1) & 1);            // the tensor core is done with frame i - 2
#pragma unroll
            for (int q = 0; q < Q; q++) {
                const int i0 = q * 512 + lane * 16;
                const uint32_t rw[4] = {raw[q].x, raw[q].y, raw[q].z, raw[q].w};
                uint32_t hi[4], lo[4];
                umma_prep16p(rw, mean, s.winp, N, i0, hi, lo);
                *reinterpret_cast<uint4 *>(ph + PAD + i0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4 *>(pl + PAD + i0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                if (G::COPIES == 2) {   // second copy, advanced by 8 bytes: sample i sits at PAD - 8 + i
                    uint8_t *const ph2 = ph + 2 * NM * PLANE, *const pl2 = ph2 + PLANE;
                    *reinterpret_cast<uint2 *>(ph2 + PAD - 8 + i0) = make_uint2(hi[0], hi[1]);
                    *reinterpret_cast<uint2 *>(ph2 + PAD + i0) = make_uint2(hi[2], hi[3]);
                    *reinterpret_cast<uint2 *>(pl2 + PAD - 8 + i0) = make_uint2(lo[0], lo[1]);
                    *reinterpret_cast<uint2 *>(pl2 + PAD + i0) = make_uint2(lo[2], lo[3]);
                }
            }
            if (p.power) {   // rolling_buffer.c:68-70
                long long acc = 0;
                for (int k = lane; k < N; k += 32) { const int dv = (int)src[k] - mean; acc += (long long)dv * dv; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) p.power[f * NM + ch] = acc;
            }
            __syncwarp();
            if (p.windowed)
                for (int ii = lane; ii < N; ii += 32)
                    p.windowed[(f * NM + ch) * (unsigned long long)N + ii] =
                        (int16_t)((int)(signed char)ph[PAD + ii] * 256 + (int)(signed char)pl[PAD + ii]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.ready[b]);
        }
    } else {
        // =================================================================== epilogue group (256 threads)
        const int et = tid - 32 * (NM + 1), wq = warp & 3, part = (warp - (NM + 1)) >> 2, m = wq * 32 + lane;
        long long *const curvef = &s.epi.curve[0][0];
        for (unsigned long long i = 0;; i++) {
            const unsigned long long f = blockIdx.x + gstride * i;
            if (f >= nf) break;
            // Passes 2i and 2i+1 hold the two digits of the same (pair, lag) items in the same order: a thread keeps the y.h sum in a
            // register and stores the finished entry after the y.l pass -- no read-modify-write, no zeroing.  Only the last pass
            // {4h, 4l} adds both digits of its four pairs (x, 4) in one go (atomics on entries zeroed here).
            if (et < 4 * NJ) {
                const int x = et / NJ;
                curvef[(x * NM - x * (x + 1) / 2 + (4 - x - 1)) * NJ + (et - x * NJ)] = 0;
            }
            long long keep[2] = {0, 0};
            named_bar(1, G::EPI_THREADS);
#pragma unroll 1
            for (int g = 0; g < G::PASSES; g++) {
                const unsigned long long gp = 7 * i + g;
                const unsigned slot = (unsigned)(gp & 1), u = (unsigned)(gp >> 1);
                mbar_wait(&s.full[slot], u & 1);
                tc_fence_after();
                const uint32_t ta = tmem + ((uint32_t)(wq * 32) << 16) + slot * G::TCOLS;
                // this warp moves couples 2 part and 2 part + 1 of its lane quadrant into the scratch
                {
                    uint32_t t[4][G::PH];     // with two copies only columns 0..7 of a tile are distinct
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        if constexpr (G::PH == 16) tmem_ld16(ta + 16 * (4 * part + e), t[e]);
                        else tmem_ld8(ta + 16 * (4 * part + e), t[e]);
                    }
                    tmem_ld_wait();
                    tc_fence_before();     // the pass is out of TMEM
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s.empty[slot]);
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        int *const zp = &s.z[2 * part + e][0][m + G::PH - 1];
#pragma unroll
                        for (int ph = 0; ph < G::PH; ph++)  // entry (m, phi): lag index m - phi; |256 a + b| < 2^31
                            if (m - ph < NJ) {
                                AT_CHECK(m + G::PH - 1 - ph >= 0 && m + G::PH - 1 - ph < G::ZP && 2 * part + e < 8);
                                zp[ph * G::ZP - ph] = 256 * (int)t[2 * e][ph] + (int)t[2 * e + 1][ph];
                            }
                    }
                }
                named_bar(1, G::EPI_THREADS);
                // diagonal sums of the eight couples: 768 (couple, lag) items over 512 threads.  The last pass {4h, 4l}
                // holds both digits of the same pairs, so two threads can meet on one curve entry: atomic add there.
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int item = et + G::EPI_THREADS * q, k = item / NJ, j = item - k * NJ;
                    if (item >= 8 * NJ) break;
                    const int *zr = &s.z[k][0][j + G::PH - 1];
                    long long sum = 0;
#pragma unroll
                    for (int ph = 0; ph < G::PH; ph++) sum += zr[ph * G::ZP];
                    const unsigned tt = s.couple_tab[g][k];
                    AT_CHECK((int)(tt >> 8) < NM * (NM - 1) / 2 && j >= 0 && j < NJ && j + G::PH - 1 + (G::PH - 1) * G::ZP < G::PH * G::ZP + G::ZP);
                    long long *const dst = &curvef[(tt >> 8) * NJ + j];
                    if (g == G::PASSES - 1) atomicAdd(reinterpret_cast<unsigned long long *>(dst), (unsigned long long)(sum << (tt & 31)));
                    else if ((g & 1) == 0) keep[q] = sum << 8;          // y.h
                    else *dst = keep[q] + sum;                          // y.l completes the entry
                }
                named_bar(1, G::EPI_THREADS);
            }
            epilogue<NM, NBITS, L, G::EPI_THREADS, 1>(s.epi, s.gauss, p, f, et);
            named_bar(1, G::EPI_THREADS);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

template <int NBITS, int L>
static cudaError_t launch_umma_m(const AtFusedParams &p, int sm_count, cudaStream_t st)
{
    auto kern = at_fused_umma_m_kernel<NBITS, L>;
    const int smem = (int)sizeof(UmmaMSmem<NBITS, L>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    unsigned long long grid = (unsigned long long)sm_count;
    if (grid > p.n_frames) grid = p.n_frames;
    if (grid == 0) return cudaSuccess;
    kern<<<(unsigned)grid, UmmaMGeo<NBITS, L>::THREADS, smem, st>>>(p);
    at_count_launch();
    return cudaGetLastError();
}

} // namespace atk

bool at_fused_umma_m_supports(const AtShape &sh)
{
    return sh.n_mics == 8 && (sh.n_bits == 10 || sh.n_bits == 12) && sh.max_shift == 46;
}

cudaError_t at_launch_fused_umma_m(const AtShape &sh, const AtFusedParams &p, int sm_count, cudaStream_t st)
{
    if (p.sig16 || !at_fused_umma_m_supports(sh)) return cudaErrorInvalidValue;
    return sh.n_bits == 12 ? atk::launch_umma_m<12, 46>(p, sm_count, st) : atk::launch_umma_m<10, 46>(p, sm_count, st);
}
