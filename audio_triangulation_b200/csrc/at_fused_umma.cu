// at_fused_umma.cu -- fused localization kernel of the reference shape (3 microphones x 1024 samples) on the
// 5th-generation tensor cores (tcgen05 / UMMA, accumulators in TMEM): AT_KERNEL_UMMA, the default for this shape.
//
// Polyphase form of the lagged cross-correlation (ref: components/correlations.c:9-18,
//     corr[s] = sum_i x[i] * y[i+s]).  Split time as i = 16 q + phi.  With Y the zero-padded y frame
// (sample i at Y[PAD + i]) the int8 matrix product
//     D[m][phi] = sum_q Y[m + 16 q] * x[phi + 16 q]            m = 0..127, phi = 0..15, q = 0..63
// holds, on its diagonals, every term of the correlation:  corr[s] = sum_phi D[s + PAD + phi][phi].
// Both operands are the PLAIN byte planes in shared memory, read as MN-major UMMA operands without swizzle:
//     A[m][q] = Y[m + 16 q]  is a Hankel matrix -- MN chunks 16 bytes apart (SBO = 16 B), K rows 16 bytes apart, so
//                            the chunks overlap in memory and nothing is materialised;
//     B[n][q] = x[n + 16 q]  is the polyphase matrix of x; several planes sit side by side (SBO = plane stride).
// One tcgen05.mma kind::i8 (K = 32) covers 512 samples: 2 K-steps per frame instead of 33 mma.sync steps.
//
// int16 samples are split into balanced signed digits w = 256 h + l, h and l both in [-128, 127];
//     corr = 65536 (h.h) + 256 (h.l + l.h) + (l.l).
// Tiles of 16 TMEM columns per pair: hh, mid = hl + lh (two MMAs accumulate into the same columns) and, in the exact
// variant, ll; int32 (|sum| < 2^26).  With H(p) the Hankel view of plane p, per K-step
//     H(c.h) x [a.h b.h a.l b.l] -> [hh_ac hh_bc mid_ac mid_bc]        H(b.h) x [a.h a.l] -> [hh_ab mid_ab]
//     H(c.l) x [a.h b.h]        +-> [mid_ac mid_bc]                     H(b.l) x [a.h]    +-> [mid_ab]
// = 8 MMAs per frame (419 cycles per frame and SM, tools/probes/epi_probe.cu); the exact variant adds
//     H(c.l) x [a.l b.l] -> [ll_ac ll_bc]   H(b.l) x [a.l] -> [ll_ab]    (10 MMAs, 507 cycles).
//
// The diagonal sums are the CUDA cores' job and never touch shared memory: an epilogue warp reads its 32 TMEM lanes
// (tcgen05.ld 32x32b: lane = row m, 16 registers = the phases), packs U = 256 hh + mid (|U| < 2^31 for this window,
// checked at context creation) and runs a five-stage register butterfly (diag_butterfly, at_umma_common.cuh): after
// stage k a lane holds the lags congruent to it modulo 2^(k+1); 16 shuffles per tile.  Only the 15 partial sums that
// cross a 32-row quarter go through shared memory.
//
// Two instantiations.
//   CERT  (lags / cell / xy / gate): only C9 = 256 U is computed.  The missing l.l product obeys
//         |ll[s]| <= sqrt(Sl_x Sl_y) =: B (Cauchy-Schwarz; the sums of squared low digits come from the frame
//         preparation), so if the maximum of U exceeds every other lag by more than 2 B / 256 its lag is THE arg-max of
//         the exact curve (no tie possible) and the position follows from the peak-tuple table.  Frames that cannot
//         be settled this way are appended to a list in global memory.
//   EXACT (whole curves, and the frames on that list): all twelve products, int64 curves, the exact epilogue.
// at_launch_fused_umma launches CERT followed by EXACT on the list (its length is read on the device), so every
// result is bit-identical to the reference either way.
//
// CTA = 28 warps, one CTA per SM, persistent, warp-specialised.  The warp scheduler prefers the highest warp id of a
// sub-partition, so the roles are numbered by how little they may be delayed:
//   warp 27      one elected lane issues the MMAs of a frame and commits them to two mbarriers
//                (accumulators ready / planes free);
//   warps 20-26  frame preparation, one frame each, two staging and two plane buffers per warp: the raw frame arrives
//                by a 1-D bulk copy (TMA) one frame ahead; DC removal, <<8, window (held in registers: lane l prepares
//                samples [16 l, +16) and their mirror image [1008 - 16 l, +16), the window is symmetric, so 16 registers
//                serve both; any ring head: unaligned heads are realigned in registers), digit planes to shared memory;
//   warps 0-19   five epilogue sets of four warps (warp w owns TMEM lane quarter w % 4); TMEM holds four accumulator
//                slots (three in the exact variant) -- a number that does not divide five, so the tensor core works a
//                frame ahead of every set.  CERT: the quarters exchange two sums per pair through a triple-buffered
//                array and ONE warp (the role rotates) decides frame n while the set already works on frame n + 1.
//                EXACT: peak-tuple look-up, for frames of the redo list a one-warp bounded search, else the group
//                epilogue of at_fused_common.cuh.
// What bounds it: the shared-memory data pipe (operand fetch 41 % + LSU 56 %), DESIGN.md 4.1.
#include <limits.h>
#include <stdlib.h>

#include "at_imma_common.cuh"
#include "at_umma_common.cuh"

// make VARIANT=prof: every warp accumulates the cycles it spends in each section of its role (lane 0, clock64) and adds
// them to p.prof[role * 8 + section] when it leaves the kernel; tools/umma_prof.py prints them per frame.
#ifdef AT_PROF
#define PROF_DECL unsigned long long prof_t = clock64(), prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PROF_MARK(k) do { const unsigned long long t_ = clock64(); prof_acc[k] += t_ - prof_t; prof_t = t_; } while (0)
#define PROF_FLUSH(role) do { if (lane == 0 && p.prof) for (int k_ = 0; k_ < 8; k_++) atomicAdd(&p.prof[(role) * 8 + k_], prof_acc[k_]); } while (0)
#else
#define PROF_DECL do { } while (0)
#define PROF_MARK(k) do { } while (0)
#define PROF_FLUSH(role) do { } while (0)
#endif

namespace atk {

template <int L, bool CERT>
struct UmmaGeo {
    static constexpr int N = 1024, NBITS = 10;
    static constexpr int PAD = 48;                      // lag index j = s + PAD; also the left zero pad of a plane
    static constexpr int PLANE = 1152;                  // bytes per plane buffer: 48 zeros, 1024 samples, 80 zeros
    static constexpr int NPLANES = 6;                   // a.h b.h a.l b.l c.h c.l
    static constexpr int FRAME = NPLANES * PLANE;       // 6 912 bytes of planes per frame
    static constexpr int NJ = 96;                       // lag slots kept (j = 0..95), j in [PAD-L, PAD+L] are real
    static constexpr int NL = 2 * L + 1;
    static constexpr int TCOLS = CERT ? 96 : 160;       // TMEM columns per accumulator slot (96 / 144 used)
    // Accumulator slots in TMEM.  Their number must not divide SETS: frame i then re-uses the slot of frame i - SLOTS, which
    // another epilogue set has drained, and the tensor core can work one frame ahead of every set.  It must not exceed SETS
    // either: the next frame of a set (i + SETS) then cannot complete before every warp of the set has released frame i, so
    // a set's `full` barrier is never two phases ahead of a waiting warp (a parity wait could not tell).
#ifndef AT_UMMA_SETS
#define AT_UMMA_SETS 5
#endif
#ifndef AT_UMMA_PREP
#define AT_UMMA_PREP 7
#endif
    static constexpr int PREP_WARPS = AT_UMMA_PREP, PBUF = 2, SETS = AT_UMMA_SETS, META = 32;
    static constexpr int SLOTS = CERT ? (SETS % 4 != 0 ? 4 : 3) : 3;
    static constexpr int THREADS = 32 * (1 + PREP_WARPS + 4 * SETS);
    static_assert(PAD >= L && PAD + L + 15 < 128 && PAD + L < NJ, "lag window must fit the 128-row tile");
    static_assert(127 + 16 * 63 + 16 <= PLANE, "A operand reads stay inside a plane buffer");
    static_assert(SLOTS * TCOLS <= 512 && SETS % SLOTS != 0 && SLOTS <= SETS, "TMEM columns / slot rotation");
    static constexpr int PREP0 = 4 * SETS, MMAW = PREP0 + PREP_WARPS;   // first prep warp, MMA warp (epilogue warp w owns TMEM lane quarter w % 4)
    static_assert(META >= PREP_WARPS * PBUF + SLOTS + 2 * SETS, "meta ring must outlive every frame in flight (decisions lag one frame)");
    static_assert(!CERT || 3 * SETS <= 15, "one named barrier per set and exchange buffer");
    // first TMEM column (within a slot) of the hh (0) / mid (1) / ll (2) tile of pair 0 = (a,b), 1 = (a,c), 2 = (b,c)
    __host__ __device__ static constexpr int col(int pr, int cls)
    {
        return CERT ? (pr == 0 ? 64 + 16 * cls : 32 * cls + (pr == 2 ? 16 : 0))
                    : (pr == 0 ? 96 + 16 * cls : 32 * cls + (pr == 2 ? 16 : 0));
    }
};
// plane index of (channel, digit): the x-side operand lists [a.h b.h a.l b.l] and [a.h a.l] must be equally spaced
__host__ __device__ constexpr int umma_plane(int ch, int digit) { return ch == 2 ? 4 + digit : ch + 2 * digit; }

template <int L, bool CERT>
struct UmmaSmem {
    using G = UmmaGeo<L, CERT>;
    alignas(128) uint8_t planes[G::PREP_WARPS * G::PBUF][G::FRAME];
    alignas(128) uint8_t rawb[G::PREP_WARPS * G::PBUF][3 * G::N];   // ring-ordered ADC bytes, staged by bulk copies (TMA)
    // CERT: diagonal sums of the three pairs by lag index, [set][frame mod 3][n0 | n1][pair][lag index]; the n1 entries no
    // quarter ever writes stay zero from the kernel's start
    alignas(16) int ubuf[CERT ? G::SETS : 1][3][2][3][128];
    alignas(16) int4 ptab[4 * G::SETS];                  // CERT: per epilogue warp, the peak-tuple table entry of its last certified frame (cp.async)
    // exact variant
    alignas(16) EpiSmem<3, 10, L> epi[CERT ? 1 : G::SETS];       // raw curves by lag index + scratch of the group epilogue
    alignas(16) long long part64[CERT ? 1 : G::SETS][3][4];      // per-warp arg-max keys
    int boxok[G::SETS];                                          // exact pass over the redo list: did the bounded search settle the frame?
    alignas(16) int spill[CERT ? 1 : G::SETS][2][6][3][32];      // partial sums that cross a lane quarter: [array][quarter below][lane]
    alignas(16) uint32_t meta[G::META][4];              // per frame: sum of squared low digits of each channel
    float gauss[2 * L + 1];
    alignas(8) uint64_t full[G::SETS], empty[G::SLOTS], ready[G::PREP_WARPS * G::PBUF], sfree[G::PREP_WARPS * G::PBUF],
        rawfull[G::PREP_WARPS * G::PBUF];
    uint32_t tmem_base;
};

// 16 ring-ordered ADC bytes of chronological samples [i0, i0 + 16) of one channel (staged in shared memory), any head
__device__ __forceinline__ uint4 load_chrono16(const uint8_t *chan, int i0, int head)
{
    constexpr int N = 1024;
    const int r = (i0 + head) & (N - 1), r0 = r & ~15, sh = r & 15;
    uint4 x = *reinterpret_cast<const uint4 *>(chan + r0);
    if (sh) x = realign16(x, *reinterpret_cast<const uint4 *>(chan + ((r0 + 16) & (N - 1))), sh);   // warp-uniform
    return x;
}

template <int L, bool CERT>
__global__ void __launch_bounds__(UmmaGeo<L, CERT>::THREADS, 1) at_fused_umma_kernel(const AtFusedParams p)
{
    using G = UmmaGeo<L, CERT>;
    using S = UmmaSmem<L, CERT>;
    constexpr int N = G::N, PAD = G::PAD, PLANE = G::PLANE, P = G::PREP_WARPS;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    S &s = *reinterpret_cast<S *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- one-time CTA set-up: zero the planes (pads stay zero), Gaussian factors, barriers, TMEM
    for (int i = tid; i < (int)(sizeof(s.planes) / 16); i += G::THREADS)
        reinterpret_cast<uint4 *>(&s.planes[0][0])[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < (int)(sizeof(s.ubuf) / 4); i += G::THREADS) (&s.ubuf[0][0][0][0][0])[i] = 0;
    for (int i = tid; i < 2 * L + 1; i += G::THREADS) s.gauss[i] = p.gauss[i];
    if (tid == 0) {
        for (int k = 0; k < G::SETS; k++) mbar_init(&s.full[k], 1);
        for (int k = 0; k < G::SLOTS; k++) mbar_init(&s.empty[k], 4);
        for (int w = 0; w < P * G::PBUF; w++) { mbar_init(&s.ready[w], 1); mbar_init(&s.sfree[w], 1); mbar_init(&s.rawfull[w], 1); }
        fence_barrier_init();
    }
    if (warp == G::MMAW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s.tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    // Work list: frames 0 .. n_frames-1, or the frames named by frame_list (its length is read here, on the device).
    const unsigned long long nf = p.frame_list ? (unsigned long long)*p.list_count : p.n_frames, gstride = gridDim.x;
    auto frame_of = [&](unsigned long long k) -> unsigned long long { return p.frame_list ? (unsigned long long)p.frame_list[k] : k; };
    // Frames of this CTA: k = blockIdx.x + gridDim.x * i, i = 0 .. mine-1; frame i is prepared by prep warp i % P into
    // its buffers (i / P) % 2, accumulated in TMEM slot i % SLOTS and finished by epilogue set i % SETS.  Every
    // mbarrier is waited on in phase order by exactly one party (a parity wait only distinguishes the current from the
    // preceding phase): ready / sfree / rawfull per plane buffer, full per epilogue set, empty per TMEM slot.
    const unsigned mine = nf > blockIdx.x ? (unsigned)((nf - blockIdx.x + gstride - 1) / gstride) : 0u;

    if (warp == G::MMAW) {
        // =================================================================== MMA issue (one elected lane, frames in order)
        constexpr uint32_t I64 = umma_idesc(64), I32 = umma_idesc(32), I16 = umma_idesc(16);
        constexpr uint32_t LBO = (128u >> 4) << 16;                          // K groups of 8 rows are 128 bytes apart
        constexpr uint32_t HI_A = (16u >> 4) | 0x4000u;                      // A: MN chunks 16 bytes apart (Hankel), version 1
        constexpr uint32_t HI_B1 = ((uint32_t)PLANE >> 4) | 0x4000u;         // B: 16 phases per plane, consecutive planes
        constexpr uint32_t HI_B2 = ((uint32_t)(2 * PLANE) >> 4) | 0x4000u;   // B: every other plane ([a.h a.l])
        static_assert(umma_plane(0, 0) == 0 && umma_plane(1, 0) == 1 && umma_plane(0, 1) == 2 && umma_plane(1, 1) == 3, "x-side operand order");
        constexpr int CH = umma_plane(2, 0), CL = umma_plane(2, 1), BH = umma_plane(1, 0), BL = umma_plane(1, 1), AL = umma_plane(0, 1);
        PROF_DECL;
        // i % P, i / P, i % SLOTS, i / SLOTS, i % SETS kept incrementally
        unsigned pw = 0, v = 0, slot = 0, u = 0, eset = 0;
        for (unsigned i = 0; i < mine; i++) {
            const unsigned pb = pw * G::PBUF + (v & 1), pv = v >> 1;
            AT_CHECK(pw == i % P && v == i / P && slot == i % G::SLOTS && u == i / G::SLOTS && eset == i % G::SETS);   // incremental counters
            PROF_MARK(0);
            mbar_wait(&s.ready[pb], pv & 1);                        // planes of frame i are in shared memory
            PROF_MARK(1);
            if (u >= 1) mbar_wait(&s.empty[slot], (u - 1) & 1);     // the epilogue has drained the slot
            PROF_MARK(2);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t b16 = (smem_u32(&s.planes[pb][0]) >> 4) + LBO;
                const uint32_t cb = tmem + slot * G::TCOLS;
                auto hk = [&](int pl, int kk) { return b16 + (uint32_t)((pl * PLANE + 512 * kk) >> 4); };          // Hankel view of a plane
                auto xs = [&](int pl, int kk) { return b16 + (uint32_t)((pl * PLANE + PAD + 512 * kk) >> 4); };    // polyphase view, first plane pl
                if (!(p.debug_skip & 2)) {
                    if constexpr (CERT) {
#pragma unroll
                        for (int kk = 0; kk < 2; kk++) {
                            umma_i8_lohi(cb + G::col(1, 0), hk(CH, kk), HI_A, xs(0, kk), HI_B1, I64, kk);
                            umma_i8_lohi(cb + G::col(1, 1), hk(CL, kk), HI_A, xs(0, kk), HI_B1, I32, 1);
                            umma_i8_lohi(cb + G::col(0, 0), hk(BH, kk), HI_A, xs(0, kk), HI_B2, I32, kk);
                            umma_i8_lohi(cb + G::col(0, 1), hk(BL, kk), HI_A, xs(0, kk), HI_B2, I16, 1);
                        }
                    } else {
                        umma_i8_lohi(cb + G::col(1, 0), hk(CH, 0), HI_A, xs(0, 0), HI_B1, I64, 0);
                        umma_i8_lohi(cb + G::col(1, 1), hk(CL, 0), HI_A, xs(0, 0), HI_B1, I32, 1);
                        umma_i8_lohi(cb + G::col(1, 2), hk(CL, 0), HI_A, xs(AL, 0), HI_B1, I32, 0);
                        umma_i8_lohi(cb + G::col(0, 0), hk(BH, 0), HI_A, xs(0, 0), HI_B2, I32, 0);
                        umma_i8_lohi(cb + G::col(0, 1), hk(BL, 0), HI_A, xs(0, 0), HI_B2, I16, 1);
                        umma_i8_lohi(cb + G::col(0, 2), hk(BL, 0), HI_A, xs(AL, 0), HI_B2, I16, 0);
                        umma_i8_lohi(cb + G::col(1, 0), hk(CH, 1), HI_A, xs(0, 1), HI_B1, I64, 1);
                        umma_i8_lohi(cb + G::col(1, 1), hk(CL, 1), HI_A, xs(0, 1), HI_B1, I64, 1);
                        umma_i8_lohi(cb + G::col(0, 0), hk(BH, 1), HI_A, xs(0, 1), HI_B2, I32, 1);
                        umma_i8_lohi(cb + G::col(0, 1), hk(BL, 1), HI_A, xs(0, 1), HI_B2, I32, 1);
                    }
                }
                umma_commit(&s.full[eset]);
                umma_commit(&s.sfree[pb]);
            }
            __syncwarp();
            if (++pw == P) { pw = 0; v++; }
            if (++slot == G::SLOTS) { slot = 0; u++; }
            if (++eset == G::SETS) eset = 0;
            PROF_MARK(3);
        }
        PROF_FLUSH(0);
    } else if (warp >= G::PREP0) {
        // =================================================================== prep warps
        const int w = warp - G::PREP0;
        // Lane l owns the chronological samples [16 l, +16) and their mirror image of every channel and frame: its slice of the doubled
        // window, pre-masked for IDP.2A (even samples in the low half, odd samples in the high half), lives in registers.
        // Its second chunk is the MIRROR of the first, [N - 16 - 16 l, +16): the window is symmetric, so the same 16 registers
        // serve both (umma_prep16m).
        uint32_t wr[16];
#pragma unroll
        for (int e = 0; e < 16; e++) wr[e] = (uint32_t)(2 * (int)p.window[lane * 16 + e]) << ((e & 1) * 16);
        auto chunk_at = [&](int q) { return q == 0 ? lane * 16 : N - 16 - lane * 16; };
        PROF_DECL;
        // Raw frames arrive by 1-D bulk copies (TMA) into this warp's two staging buffers, one frame ahead: lane 0 starts
        // the copy of frame i + P as soon as the buffer's previous frame has been consumed.
        auto stage = [&](unsigned n2) {                // lane 0 only; n2 = index among this warp's frames
            const unsigned i2 = (unsigned)w + (unsigned)P * n2;
            if (i2 >= mine) return;
            const unsigned rb = (unsigned)w * G::PBUF + (n2 & 1);
            mbar_expect_tx(&s.rawfull[rb], 3 * N);
            bulk_g2s(&s.rawb[rb][0], p.adc + frame_of(blockIdx.x + gstride * i2) * (unsigned long long)(3 * N), 3 * N, &s.rawfull[rb]);
        };
        if (lane == 0) { stage(0); stage(1); }
        unsigned mi = (unsigned)w % G::META;            // i % META, kept incrementally
        for (unsigned n = 0, i = (unsigned)w; i < mine; n++, i += P, mi = mi + P >= G::META ? mi + P - G::META : mi + P) {
            const unsigned long long f = frame_of(blockIdx.x + gstride * i);
            PROF_MARK(0);
            const unsigned pb = (unsigned)w * G::PBUF + (n & 1), pv = n >> 1;
            AT_CHECK(pb < (unsigned)(P * G::PBUF) && mi < (unsigned)G::META);
            uint8_t *const buf = &s.planes[pb][0];
            const uint8_t *const src = &s.rawb[pb][0];
            const int head = p.heads ? (p.heads[f] & (N - 1)) : 0;
            mbar_wait_slack(&s.rawfull[pb], pv & 1);         // the frame's bytes are in shared memory

            // The lane's chronological chunks [512 q + 16 l, +16) of every channel, un-rotated from the ring (one aligned
            // 16-byte load, or two and a byte shift when the head is not 16-aligned); they stay in registers for the second
            // pass.  Channel sums -> floor mean (rolling_buffer.c:48-64).
            const int uhead = __shfl_sync(0xffffffffu, head, 0);      // tells the compiler what it cannot see: warp-uniform
            uint4 raw[6];
            int mean[3];
            {
                unsigned sum[3];
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    sum[ch] = 0;
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const uint4 x = load_chrono16(src + ch * N, chunk_at(q), uhead);
                        raw[ch * 2 + q] = x;
                        sum[ch] = __dp4a(x.x, 0x01010101u, sum[ch]); sum[ch] = __dp4a(x.y, 0x01010101u, sum[ch]);
                        sum[ch] = __dp4a(x.z, 0x01010101u, sum[ch]); sum[ch] = __dp4a(x.w, 0x01010101u, sum[ch]);
                    }
                }
#pragma unroll
                for (int ch = 0; ch < 3; ch++) mean[ch] = (int)(__reduce_add_sync(0xffffffffu, sum[ch]) >> 10);
            }
            PROF_MARK(1);
            // the tensor core must be done with the frame that used this plane buffer before
            if (pv >= 1) mbar_wait_slack(&s.sfree[pb], (pv - 1) & 1);
            PROF_MARK(2);
            unsigned sl[3] = {0, 0, 0};
            auto prep_chunk = [&](int ch, int q, const uint4 x) {
                const int i0 = chunk_at(q);
                const uint32_t rw[4] = {x.x, x.y, x.z, x.w};
                uint32_t hi[4], lo[4];
                if (q == 0) umma_prep16r(rw, mean[ch], wr, hi, lo);
                else umma_prep16m(rw, mean[ch], wr, hi, lo);
                AT_CHECK(umma_plane(ch, 1) * PLANE + PAD + i0 + 16 <= G::FRAME);
                *reinterpret_cast<uint4 *>(buf + umma_plane(ch, 0) * PLANE + PAD + i0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4 *>(buf + umma_plane(ch, 1) * PLANE + PAD + i0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                if constexpr (CERT) {
#pragma unroll
                    for (int w4 = 0; w4 < 4; w4++) sl[ch] = (unsigned)__dp4a((int)lo[w4], (int)lo[w4], (int)sl[ch]);
                }
            };
            if (!(p.debug_skip & 1)) {
#pragma unroll
                for (int ch = 0; ch < 3; ch++)
#pragma unroll
                    for (int q = 0; q < 2; q++) prep_chunk(ch, q, raw[ch * 2 + q]);
            }
            if constexpr (CERT) {
#pragma unroll
                for (int ch = 0; ch < 3; ch++) sl[ch] = __reduce_add_sync(0xffffffffu, sl[ch]);
                if (lane < 3) s.meta[mi][lane] = lane == 0 ? sl[0] : (lane == 1 ? sl[1] : sl[2]);
            }
            PROF_MARK(3);
            if (p.power) {   // rolling_buffer.c:68-70
                for (int ch = 0; ch < 3; ch++) {
                    long long acc = 0;
                    for (int j = lane; j < N; j += 32) { const int dv = (int)src[ch * N + j] - mean[ch]; acc += (long long)dv * dv; }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    if (lane == 0) p.power[f * 3 + ch] = acc;
                }
            }
            __syncwarp();
            if (p.windowed)
                for (int idx = lane; idx < 3 * N; idx += 32) {
                    const int ch = idx / N, ii = idx % N;
                    p.windowed[f * (unsigned long long)(3 * N) + idx] =
                        (int16_t)((int)(signed char)buf[umma_plane(ch, 0) * PLANE + PAD + ii] * 256 + (int)(signed char)buf[umma_plane(ch, 1) * PLANE + PAD + ii]);
                }
            fence_proxy_async();          // this lane's plane bytes -> visible to the tensor core's reads
            __syncwarp();                 // ... and every lane is done reading the staged bytes
            if (lane == 0) { mbar_arrive(&s.ready[pb]); stage(n + 2); }
            PROF_MARK(4);
        }
        PROF_FLUSH(1);
    } else {
        // =================================================================== epilogue sets
        const int set = warp >> 2, wq = warp & 3, m = wq * 32 + lane;   // m = TMEM lane = tile row = lag index of n0
        const bool valid = m >= PAD - L && m <= PAD + L;
        const bool wants_pos = p.cell || p.xy;
        const int bar_a = 1 + 2 * set, bar_b = 2 + 2 * set;      // exact variant
        unsigned par = 0;
        // CERT: a certified frame whose peak-tuple table entry is on its way into shared memory (cp.async, lane 0); it is
        // consumed at this warp's next decision, so the table's latency stays off every chain
        bool pend = false;
        unsigned long long pend_f = 0;
        auto flush_pending = [&]() {
            if (!pend) return;
            pend = false;
            if (lane == 0) {
                asm volatile("cp.async.wait_all;" ::: "memory");
                const int4 e = s.ptab[warp];
                if (e.x >= 0) {
                    if (p.cell) p.cell[pend_f] = e.x;
                    if (p.xy) reinterpret_cast<float2 *>(p.xy)[pend_f] = make_float2(__int_as_float(e.y), __int_as_float(e.z));
                    if (p.stats) { atomicAdd(&p.stats[3], 1ull); atomicAdd(&p.stats[4], 1ull); }
                } else {
                    p.redo_list[atomicAdd(p.redo_count, 1u)] = (uint32_t)pend_f;     // lags that are no tuple of the LUT: exact search
                }
            }
        };
        // CERT: the decision of frame n_d of this set (exchange buffer kb_d = n_d % 3), by the warp whose turn it is
        auto decide = [&](unsigned n_d, unsigned kb_d) {
            if constexpr (CERT) {
                named_bar(1 + 3 * set + (int)kb_d, 128);       // completes at once: the other quarters arrived a frame ago
                flush_pending();
                const unsigned i_d = (unsigned)set + (unsigned)G::SETS * n_d;
                const unsigned long long f = frame_of(blockIdx.x + gstride * i_d);
                const uint32_t *const sl = s.meta[i_d % G::META];
                int (*const ub)[3][128] = s.ubuf[set][kb_d];
                bool sure = true;
                int b3[3];
#pragma unroll
                for (int pr = 0; pr < 3; pr++) {
                    // lane holds lag indices j = lane, lane + 32, lane + 64 (ascending): U[j] = n0[j] + n1[j]
                    int v[3];
#pragma unroll
                    for (int t = 0; t < 3; t++) {
                        const int j = lane + 32 * t;
                        v[t] = (j >= PAD - L && j <= PAD + L) ? ub[0][pr][j] + ub[1][pr][j] : INT_MIN;
                    }
                    // first-max arg-max (correlations.c:20-23 on C9 = 256 U) and the runner-up
                    const int top_l = max(v[0], max(v[1], v[2]));
                    const int j_l = v[0] == top_l ? lane : (v[1] == top_l ? lane + 32 : lane + 64);
                    const int top = __reduce_max_sync(0xffffffffu, top_l);
                    const int j1 = __reduce_min_sync(0xffffffffu, top_l == top ? j_l : 0x7fffffff);
                    const int sec_l = max(j1 == lane ? INT_MIN : v[0], max(j1 == lane + 32 ? INT_MIN : v[1], j1 == lane + 64 ? INT_MIN : v[2]));
                    const int second = __reduce_max_sync(0xffffffffu, sec_l);
                    // |ll| <= B = sqrt(Sl_x Sl_y) (rounded up, < 2^25): the arg-max is certain when 256 (top - second) > 2 B
                    // and the exact peak 256 top - B >= 2048; both tested a little conservatively in 32-bit arithmetic
                    const int xc = pr == 2 ? 1 : 0, yc = pr == 0 ? 1 : 2;
                    const unsigned bnd = (unsigned)sqrt_prod_up(sl[xc], sl[yc]) + 1u;
                    sure = sure && (unsigned)(top - second) > (bnd >> 7) + 1u && top > (int)((bnd + 2048u) >> 8) + 1;
                    b3[pr] = j1 - PAD;
                }
                if (sure) {
                    // certified lags are final; the position comes from the peak-tuple table
                    if (lane < 3 && p.lags) p.lags[f * 3 + lane] = lane == 0 ? b3[0] : (lane == 1 ? b3[1] : b3[2]);
                    if (lane == 0 && p.gate) p.gate[f] = (b3[0] * b3[0] + b3[1] * b3[1] + b3[2] * b3[2]) > 4 ? 1 : 0;   // sample_compute.h:124-134
                    if (wants_pos) {
                        if (lane == 0) {
                            const int idx = ((b3[0] + L) * (2 * L + 1) + (b3[1] + L)) * (2 * L + 1) + (b3[2] + L);
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smem_u32(&s.ptab[warp])), "l"(p.peak_tab + idx) : "memory");
                        }
                        pend_f = f; pend = true;
                    } else if (lane == 0 && p.stats) atomicAdd(&p.stats[4], 1ull);
                } else if (lane == 0) {
                    p.redo_list[atomicAdd(p.redo_count, 1u)] = (uint32_t)f;      // the exact variant finishes this frame
                }
            }
        };
        PROF_DECL;
        unsigned slot = (unsigned)set % G::SLOTS, mi = (unsigned)set % G::META;   // i % SLOTS, i % META, kept incrementally
        unsigned kb = 0;      // n % 3
        for (unsigned n = 0, i = (unsigned)set; i < mine; n++, i += G::SETS, par ^= 1, kb = kb == 2 ? 0 : kb + 1,
                      slot = (slot + G::SETS) % G::SLOTS, mi = mi + G::SETS >= G::META ? mi + G::SETS - G::META : mi + G::SETS) {
            const unsigned long long k = blockIdx.x + gstride * i;
            PROF_MARK(0);
            mbar_wait_slack(&s.full[set], n & 1);            // full[] is per set (waited in order by that set), empty[] per slot
            PROF_MARK(1);
            tc_fence_after();
            auto release_slot = [&]() {   // this warp's accumulators are out of TMEM: hand the slot back to the tensor core
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.empty[slot]);
            };
            if (p.debug_skip & 4) { release_slot(); continue; }      // timing experiments: drain the slot without looking at it
            AT_CHECK(slot * G::TCOLS + (CERT ? 96 : 144) <= 512 && slot < G::SLOTS);
            const uint32_t ta = tmem + ((uint32_t)(wq * 32) << 16) + slot * G::TCOLS;

            if constexpr (CERT) {
                // ---- U = 256 hh + mid of the three pairs, diagonal sums in registers
                int u0[3], u1[3];
                {
                    // all six tiles leave TMEM before the first butterfly, so that the slot goes back to the tensor core early
                    Arr<16> a1, a2, a0;
                    {   // one pair at a time: 32 registers in flight next to the packed arrays (no spills); the two extra TMEM
                        // round trips are covered by the other sets
                        uint32_t h[16], md[16];
                        tmem_ld16(ta + G::col(1, 0), h); tmem_ld16(ta + G::col(1, 1), md);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; j++) a1.v[j] = (int)h[j] * 256 + (int)md[j];
                        asm volatile("" :: "r"(a1.v[0]), "r"(a1.v[5]), "r"(a1.v[10]), "r"(a1.v[15]));   // pack before the next loads are issued
                        tmem_ld16(ta + G::col(2, 0), h); tmem_ld16(ta + G::col(2, 1), md);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; j++) a2.v[j] = (int)h[j] * 256 + (int)md[j];
                        asm volatile("" :: "r"(a2.v[0]), "r"(a2.v[5]), "r"(a2.v[10]), "r"(a2.v[15]));
                        tmem_ld16(ta + G::col(0, 0), h); tmem_ld16(ta + G::col(0, 1), md);
                        tmem_ld_wait();
                        release_slot();
#pragma unroll
                        for (int j = 0; j < 16; j++) a0.v[j] = (int)h[j] * 256 + (int)md[j];
                    }
                    if (p.debug_skip & 32) {     // timing experiments: no diagonal sums
#pragma unroll
                        for (int pr = 0; pr < 3; pr++) { u0[pr] = a1.v[pr] ^ a2.v[pr + 3] ^ a0.v[pr + 6]; u1[pr] = a1.v[pr + 9] + a2.v[pr + 12] + a0.v[15 - pr]; }
                    } else {
                        diag_butterfly(a0, lane, u0[0], u1[0]);          // the single one first: fewest registers live at the peak
                        diag_butterfly2(a1, a2, lane, u0[1], u1[1], u0[2], u1[2]);
                    }
                }
                PROF_MARK(2);
                // Every quarter leaves its sums in shared memory: n0 = lag index m (complete up to the rows of the next
                // quarter) and, lanes >= 17, n1 = the part of lag index m - 32 that this quarter's rows hold.  One warp per
                // frame -- the role rotates over the four quarters -- adds the two, finds the arg-max and decides; the other
                // three arrive on the barrier and go on to their next frame.
                int (*const ub)[3][128] = s.ubuf[set][kb];
#pragma unroll
                for (int pr = 0; pr < 3; pr++) {
                    AT_CHECK(m >= 0 && m < 128 && kb < 3 && set < G::SETS);
                    ub[0][pr][m] = u0[pr];
                    if (wq >= 1 && lane >= 17) { AT_CHECK(m - 32 >= 17 && m - 32 < 96); ub[1][pr][m - 32] = u1[pr]; }
                }
                if (p.debug_skip & 16) continue;   // timing experiments: nobody decides
                // (bar.arrive orders the stores above: PTX producer / consumer pattern.)  The deciding warp of this frame does
                // not stop here either: it synchronises and decides one frame later, after its own work on the next frame, when
                // the other three have long arrived -- no warp ever waits for a decision.
                if (wq != (int)(n & 3)) named_bar_arrive(1 + 3 * set + (int)kb, 128);
                PROF_MARK(3);
                if (n >= 1 && wq == (int)((n - 1) & 3)) decide(n - 1, kb == 0 ? 2u : kb - 1);
                PROF_MARK(6);
            } else {
                // ---- exact variant: U = 256 hh + mid and ll of the three pairs; corr = 256 U + ll
                const unsigned long long f = frame_of(k);
                int (*const spill)[3][32] = s.spill[set][par];
                EpiSmem<3, 10, L> &epi = s.epi[set];
                static_assert(Geo<10, L>::PADL == PAD && Geo<10, L>::NLAGS_PAD == G::NJ, "curve layout of the group epilogue");
                int u0[3], u1[3], l0[3], l1[3];
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    const int pr = q == 0 ? 1 : (q == 1 ? 2 : 0);
                    uint32_t h[16], md[16], ll[16];
                    tmem_ld16(ta + G::col(pr, 0), h); tmem_ld16(ta + G::col(pr, 1), md); tmem_ld16(ta + G::col(pr, 2), ll);
                    tmem_ld_wait();
                    if (q == 2) release_slot();
                    Arr<16> a, b;
#pragma unroll
                    for (int j = 0; j < 16; j++) { a.v[j] = (int)h[j] * 256 + (int)md[j]; b.v[j] = (int)ll[j]; }
                    diag_butterfly2(a, b, lane, u0[pr], u1[pr], l0[pr], l1[pr]);
                }
                PROF_MARK(2);
                if (wq >= 1 && lane >= 17) {
#pragma unroll
                    for (int pr = 0; pr < 3; pr++) { spill[pr][wq - 1][lane] = u1[pr]; spill[3 + pr][wq - 1][lane] = l1[pr]; }
                }
                named_bar(bar_a, 128);
                PROF_MARK(3);
#pragma unroll
                for (int pr = 0; pr < 3; pr++) {
                    if (wq < 3 && lane >= 17) { u0[pr] += spill[pr][wq][lane]; l0[pr] += spill[3 + pr][wq][lane]; }
                    const long long val = 256LL * (long long)u0[pr] + (long long)l0[pr];
                    if (m < G::NJ) epi.curve[pr][m] = val;
                    long long key = valid ? val * 128 + (127 - m) : LLONG_MIN;   // largest value, then lowest lag
                    key = warp_max_i64(key);
                    if (lane == 0) s.part64[set][pr][wq] = key;
                }
                named_bar(bar_b, 128);
                PROF_MARK(4);
                {   // every thread of the set: the three first-max lags and peaks (correlations.c:20-23)
                    int best3[3];
                    long long peak[3];
#pragma unroll
                    for (int pr = 0; pr < 3; pr++) {
                        long long key = s.part64[set][pr][0];
#pragma unroll
                        for (int q = 1; q < 4; q++) if (s.part64[set][pr][q] > key) key = s.part64[set][pr][q];
                        best3[pr] = 127 - (int)(key & 127) - PAD;
                        peak[pr] = key >> 7;
                    }
                    if (m < 3 && p.lags) p.lags[f * 3 + m] = m == 0 ? best3[0] : (m == 1 ? best3[1] : best3[2]);
                    const bool extras = p.gate || p.raw || p.corr || p.cell || p.highest || p.xy || p.classes;
                    bool settled = !extras;
                    // position products only: consistent peaks are settled by one table load (thread 0 of the set stores)
                    if (extras && !(p.raw || p.corr || p.classes))
                        settled = peak_tuple_lookup<L>(p, f, m, best3[0], best3[1], best3[2], peak);
                    // everything else -- whole curves, flat or inconsistent curves -- by the whole set: Gaussian re-weighting,
                    // result stores and the likelihood maximum over all LUT tuples with 128 threads
                    // a frame the certified pass could not settle usually has peaked curves whose three lags miss the LUT by one:
                    // one warp's exact bounded search around the peak tuple (at_imma_common.cuh) decides most of them; frames of
                    // an exact-only launch (whole curves, or input that certifies nothing) go straight to the full scan
                    if (!settled && p.frame_list && !(p.raw || p.corr || p.classes)) {
                        if (wq == 0) {
                            const bool ok = epilogue_warp<L, PAD, G::NJ, G::NJ, false>(&epi.curve[0][0], best3[0], best3[1], best3[2], s.gauss, p, f, lane);
                            if (lane == 0) s.boxok[set] = ok ? 1 : 0;
                        }
                        named_bar(bar_b, 128);
                        settled = s.boxok[set] != 0;
                    }
                    if (!settled) {
                        if (m == 0 && p.stats && (p.cell || p.highest || p.xy || p.classes)) atomicAdd(&p.stats[2], 1ull);   // route: full scan
                        if (m < 3) epi.best[m] = m == 0 ? best3[0] : (m == 1 ? best3[1] : best3[2]);      // the group epilogue skips its own arg-max
                        epilogue<3, 10, L, 128, 0, true>(epi, s.gauss, p, f, m, bar_a);
                    }
                }
                named_bar(bar_a, 128);      // curve / part64 are rewritten by the next frame of this set
                PROF_MARK(6);
            }
        }
        if constexpr (CERT) {
            // the set's last frame has not been decided yet
            const unsigned cnt = mine > (unsigned)set ? (mine - (unsigned)set + G::SETS - 1) / G::SETS : 0u;
            if (cnt >= 1 && !(p.debug_skip & (4 | 16)) && wq == (int)((cnt - 1) & 3)) decide(cnt - 1, (cnt - 1) % 3);
            flush_pending();
        }
        PROF_FLUSH(2 + (wq == 0 ? 0 : 1));
    }
    tc_fence_before();
    __syncthreads();
    if (warp == G::MMAW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

} // namespace atk

bool at_fused_umma_supports(const AtShape &sh)
{
    return sh.n_mics == 3 && sh.n_bits == 10 && (sh.max_shift == 46 || sh.max_shift == 44);
}

// |256 hh + mid| must fit an int32 for every possible frame: |h| <= hmax_i = ((W_i + 128) >> 8) + 1, |l| <= 128;
// and the window must be symmetric (the prep warps keep half of it in registers)
bool at_fused_umma_window_ok(const int16_t *window, int n)
{
    for (int i = 0; i < n / 2; i++)
        if (window[i] != window[n - 1 - i]) return false;
    long long s2 = 0, s1 = 0;
    for (int i = 0; i < n; i++) { const long long h = (((long long)window[i] + 128) >> 8) + 1; s2 += h * h; s1 += h; }
    return 256 * s2 + 2 * 128 * s1 < (1ll << 31);
}

template <int L, bool CERT>
static cudaError_t launch_umma(const AtFusedParams &p, unsigned long long work, int sm_count, cudaStream_t st)
{
    auto kern = atk::at_fused_umma_kernel<L, CERT>;
    const int smem = (int)sizeof(atk::UmmaSmem<L, CERT>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    unsigned long long grid = (unsigned long long)sm_count;
    if (grid > work) grid = work;
    if (grid == 0) return cudaSuccess;
    kern<<<(unsigned)grid, atk::UmmaGeo<L, CERT>::THREADS, smem, st>>>(p);
    at_count_launch();
    return cudaGetLastError();
}

// redo: device scratch of 4 + 4 * n_frames bytes (a counter, zero on entry, followed by the list) or NULL.
cudaError_t at_launch_fused_umma(const AtShape &sh, const AtFusedParams &p_in, uint32_t *redo, int sm_count, cudaStream_t st)
{
    if (p_in.sig16 || !at_fused_umma_supports(sh)) return cudaErrorInvalidValue;
    AtFusedParams p = p_in;
    const bool wants_curves = p.raw || p.corr || p.classes || p.highest;
    const bool cert = redo && !wants_curves && (!(p.cell || p.xy) || p.peak_tab) && !(p.debug_skip & 8) && p.n_frames < (1ull << 32);
    cudaError_t e;
    if (!cert) return sh.max_shift == 46 ? launch_umma<46, false>(p, p.n_frames, sm_count, st) : launch_umma<44, false>(p, p.n_frames, sm_count, st);
    p.redo_count = redo; p.redo_list = redo + 1;
    e = sh.max_shift == 46 ? launch_umma<46, true>(p, p.n_frames, sm_count, st) : launch_umma<44, true>(p, p.n_frames, sm_count, st);
    if (e != cudaSuccess) return e;
    // the frames the certified pass could not settle: exact variant over the list (length read on the device)
    p.frame_list = redo + 1; p.list_count = redo; p.redo_count = nullptr; p.redo_list = nullptr;
    return sh.max_shift == 46 ? launch_umma<46, false>(p, p.n_frames, sm_count, st) : launch_umma<44, false>(p, p.n_frames, sm_count, st);
}

// ------------------------------------------------------------------ roofline denominators, measured live (at_microbench)
namespace atk {
// which = 0: dense tcgen05.mma kind::i8 M128 x N256 x K32 back to back (the tensor core's int8 rate);
// which = 1: the eight MMAs of a frame of the certified variant (Hankel A, N = 64 / 32 / 32 / 16 per K-step) from rotating
//            plane buffers: what this formulation can draw from the tensor core when nothing else runs.
__global__ void __launch_bounds__(128) umma_ubench_kernel(int which, int reps, long long *cycles)
{
    constexpr int PLANE = 1152, FRAMEB = 6 * PLANE, NBUF = 7, PAD = 48;
    extern __shared__ __align__(128) uint8_t dyn[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < NBUF * FRAMEB / 4; i += 128) reinterpret_cast<uint32_t *>(dyn)[i] = (uint32_t)i * 2654435761u;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    const long long t0 = clock64();
    if (warp == 0) {
        if (elect_one()) {
            constexpr uint32_t LBO = (128u >> 4) << 16, HI_A = (16u >> 4) | 0x4000u;
            constexpr uint32_t HI_B1 = ((uint32_t)PLANE >> 4) | 0x4000u, HI_B2 = ((uint32_t)(2 * PLANE) >> 4) | 0x4000u;
            constexpr uint32_t I256 = umma_idesc(256), I64 = umma_idesc(64), I32 = umma_idesc(32), I16 = umma_idesc(16);
            for (int r = 0; r < reps; r++) {
                const uint32_t b16 = (smem_u32(dyn + (r % NBUF) * FRAMEB) >> 4) + LBO;
                const uint32_t cb = tmem + (uint32_t)(r & 3) * 128;
                auto hk = [&](int pl, int kk) { return b16 + (uint32_t)((pl * PLANE + 512 * kk) >> 4); };
                auto xs = [&](int pl, int kk) { return b16 + (uint32_t)((pl * PLANE + PAD + 512 * kk) >> 4); };
                if (which == 0) {
                    // B: 16 "planes" of 16 phases, 128 bytes apart (any readable bytes will do for a rate measurement)
                    umma_i8_lohi(tmem + (uint32_t)(r & 1) * 256, hk(0, r & 1), HI_A, xs(1, 0), (128u >> 4) | 0x4000u, I256, 1);
                } else {
#pragma unroll
                    for (int kk = 0; kk < 2; kk++) {
                        umma_i8_lohi(cb + 0, hk(4, kk), HI_A, xs(0, kk), HI_B1, I64, kk);
                        umma_i8_lohi(cb + 32, hk(5, kk), HI_A, xs(0, kk), HI_B1, I32, 1);
                        umma_i8_lohi(cb + 64, hk(1, kk), HI_A, xs(0, kk), HI_B2, I32, kk);
                        umma_i8_lohi(cb + 80, hk(3, kk), HI_A, xs(0, kk), HI_B2, I16, 1);
                    }
                }
            }
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
    }
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}
} // namespace atk

// which = 0: *gops = G int8-MAC/s of dense tcgen05 MMAs; which = 1: *gops = G frames/s of the certified variant's MMA sequence
cudaError_t at_run_microbench_umma(int which, int sm_count, double *gops, double *mhz, cudaStream_t st)
{
    long long *d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(long long));
    if (e != cudaSuccess) return e;
    const int smem = 7 * 6 * 1152, reps = which == 0 ? 20000 : 4000;
    e = cudaFuncSetAttribute(atk::umma_ubench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { cudaFree(d); return e; }
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0); cudaEventCreate(&ev1);
    atk::umma_ubench_kernel<<<sm_count, 128, smem, st>>>(which, 64, d);          // warm-up
    cudaEventRecord(ev0, st);
    atk::umma_ubench_kernel<<<sm_count, 128, smem, st>>>(which, reps, d);
    cudaEventRecord(ev1, st);
    at_count_launch(2);
    e = cudaEventSynchronize(ev1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    long long cyc = 0;
    cudaMemcpy(&cyc, d, sizeof cyc, cudaMemcpyDeviceToHost);
    cudaFree(d);
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    if (e != cudaSuccess) return e;
    const double per_rep = which == 0 ? 128.0 * 256.0 * 32.0 : 1.0;
    *gops = per_rep * reps * sm_count / (ms * 1e-3) / 1e9;
    if (mhz) *mhz = (double)cyc / (ms * 1e-3) / 1e6;
    return cudaGetLastError();
}
