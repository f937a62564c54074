// at_fused_common.cuh -- device building blocks shared by the fused localization kernels:
// bulk-copy (TMA 1-D) frame staging, frame preparation (DC removal, <<8, window), and the
// epilogue (first-max arg-max, Gaussian re-weighting, likelihood-map arg-max, result stores).
//
// Reference semantics restated here (paths relative to the reference's src/):
//   components/rolling_buffer.c:43-71   un-rotate ring, subtract (int16)(sum >> n_bits), power
//   components/buffer.c:13-18           x <<= 8 with int16 wrap
//   components/buffer.c:4-11            x = (int16)((x * W[i]) >> 15)
//   components/correlations.c:20-23     arg-max, strict '>', ascending lag
//   components/correlations.c:26-33     c = (int64)((float)c * (float)exp(-(s-best)^2/36.f))
//   components/vga/vga_heatmap.h:96-108 L(cell) = sum_pairs c[lut], max, (ours) first cell
//   sample_compute.h:124-134            gate = sum best^2 > 4
#pragma once
#include <stdio.h>

#include "at_internal.h"

namespace atk {

// Debug build (make CHECKED=1 -> -DAT_CHECKED): index / ownership assertions inside the kernels; a failed one prints
// its location and traps, so the launch returns an error instead of silently corrupting shared memory or TMEM.
#ifdef AT_CHECKED
#define AT_CHECK(cond) do { if (!(cond)) { printf("AT_CHECK failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); __trap(); } } while (0)
#else
#define AT_CHECK(cond) do { } while (0)
#endif

constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }
constexpr int round_up(int a, int b) { return ceil_div(a, b) * b; }

// Lag geometry of the shared-memory signal rows.  A row holds PADL zeros, the N prepared
// samples, then zeros, so that x[i] * y[i+s] needs no bounds test for |s| <= L.
template <int NBITS, int L>
struct Geo {
    static constexpr int N = 1 << NBITS;
    static constexpr int PADL = round_up(L, 8);                 // 48 for L = 46
    static constexpr int NBLK = ceil_div(PADL + L + 1, 8);      // lag blocks of 8: 12
    static constexpr int NLAGS_PAD = NBLK * 8;                  // 96 computed lags, s = j - PADL
    static constexpr int ROW = PADL + N + round_up(NLAGS_PAD - PADL + 8, 8); // int16 elements: 1128
    static constexpr int NL = 2 * L + 1;
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    // try_wait suspends the thread until the phase completes or the time hint (ns) expires; without a hint it returns
    // after a short system-defined time and the retry loop eats issue slots of the warps that have work
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
}
// Wait of a warp that has slack: one probe, then fixed naps between probes.  A try_wait with a suspend hint compiles to
// NANOSLEEP.SYNCS, which wakes on every mbarrier event of the CTA -- with a dozen waiting warps and twenty events per frame
// that is ~200 wake-ups (each two shared-memory probes) per frame; a plain nap costs a little wake-up latency instead.
#ifndef AT_WAIT_NS
#define AT_WAIT_NS 100
#endif
__device__ __forceinline__ void mbar_wait_slack(uint64_t *bar, uint32_t parity)
{
#if AT_WAIT_NS > 0
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "NAP_%=:\n\t"
        "nanosleep.u32 %2;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra NAP_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)AT_WAIT_NS) : "memory");
#else
    mbar_wait(bar, parity);
#endif
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ int sext_lo16(uint32_t w) { return (int)(short)(w & 0xFFFFu); }
__device__ __forceinline__ int sext_hi16(uint32_t w) { return ((int)w) >> 16; }

// ---------------------------------------------------------------- frame preparation
// One prepared sample from one ADC byte (or ring int16): DC removal, <<8 wrap, Q15 window.
__device__ __forceinline__ int prep_sample(int raw, int mean16, int w)
{
    const int v = (int)(short)(raw - mean16);         // rolling_buffer.c:66 (int16 -= int16)
    const int u = (int)(short)((unsigned)v << 8);     // buffer.c:16
    return (int)(short)((u * w) >> 15);               // buffer.c:8-9
}

// Prepare all channels of one frame held in shared memory as ring-ordered bytes.
//   raw  : [NMICS][N] uint8 (ring order)           sig : [NMICS][ROW] int16 (pads pre-zeroed)
//   dcs  : [NMICS] int scratch                      win : [N] int16 window
// Must be called by all THREADS threads; contains __syncthreads().
template <int NMICS, int NBITS, int L, int THREADS>
__device__ __forceinline__ void prep_frame_u8(const uint8_t *raw, int head, int *dcs, const int16_t *win,
                                              int16_t *sig, const AtFusedParams &p, unsigned long long f)
{
    using G = Geo<NBITS, L>;
    constexpr int N = G::N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARPS = THREADS / 32;
    // channel sums: one warp per channel, 4 bytes per dp4a (the sum is rotation-invariant)
    for (int ch = warp; ch < NMICS; ch += NWARPS) {
        const uint32_t *w32 = reinterpret_cast<const uint32_t *>(raw + ch * N);
        int s = 0;
#pragma unroll 4
        for (int k = lane; k < N / 4; k += 32) s = (int)__dp4a(w32[k], 0x01010101u, (unsigned)s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) dcs[ch] = (int)(short)(s >> NBITS);          // rolling_buffer.c:64
    }
    __syncthreads();
    // 8 ring positions per thread-iteration
    for (int idx = tid; idx < NMICS * (N / 8); idx += THREADS) {
        const int ch = idx / (N / 8), j = (idx % (N / 8)) * 8;
        const uint2 b = *reinterpret_cast<const uint2 *>(raw + ch * N + j);
        const int mean = dcs[ch];
        const int i0 = (j - head) & (N - 1);                          // chronological index of ring slot j
        int16_t *dst = sig + ch * G::ROW + G::PADL;
        int o[8];
        if ((head & 7) == 0) {
            const uint4 wv = *reinterpret_cast<const uint4 *>(win + i0);
            const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int byte = (int)(((k < 4 ? b.x : b.y) >> (8 * (k & 3))) & 0xFFu);
                const int w = (k & 1) ? sext_hi16(ww[k >> 1]) : sext_lo16(ww[k >> 1]);
                o[k] = prep_sample(byte, mean, w);
            }
            uint4 pk;
            pk.x = (uint32_t)(o[0] & 0xFFFF) | ((uint32_t)o[1] << 16);
            pk.y = (uint32_t)(o[2] & 0xFFFF) | ((uint32_t)o[3] << 16);
            pk.z = (uint32_t)(o[4] & 0xFFFF) | ((uint32_t)o[5] << 16);
            pk.w = (uint32_t)(o[6] & 0xFFFF) | ((uint32_t)o[7] << 16);
            *reinterpret_cast<uint4 *>(dst + i0) = pk;
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int byte = (int)(((k < 4 ? b.x : b.y) >> (8 * (k & 3))) & 0xFFu);
                const int i = (i0 + k) & (N - 1);
                dst[i] = (int16_t)prep_sample(byte, mean, win[i]);
            }
        }
    }
    __syncthreads();
    // optional debug/parity products
    if (p.windowed) {
        for (int idx = tid; idx < NMICS * N; idx += THREADS)
            p.windowed[f * (unsigned long long)(NMICS * N) + idx] = sig[(idx / N) * G::ROW + G::PADL + (idx % N)];
    }
    if (p.power) {   // rolling_buffer.c:68-70 on the DC-removed (pre-shift) samples
        for (int ch = warp; ch < NMICS; ch += NWARPS) {
            long long acc = 0;
            for (int k = lane; k < N; k += 32) {
                const int v = (int)(short)((int)raw[ch * N + k] - dcs[ch]);
                acc += (long long)v * v;
            }
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o2);
            if (lane == 0) p.power[f * NMICS + ch] = acc;
        }
    }
}

// ---------------------------------------------------------------- epilogue
struct Best { long long v; int i; };
__device__ __forceinline__ Best best_of(Best a, Best b)
{   // larger value wins; equal values: lower index wins (== first in ascending scan order)
    return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ Best warp_best(Best x)
{   // three REDUX instead of five rounds of 64-bit shuffles: maximum of the high words, of the low words among those, lowest index
    const int hi = (int)(x.v >> 32);
    const unsigned lo = (unsigned)x.v;
    const int mh = __reduce_max_sync(0xffffffffu, hi);
    const bool in1 = hi == mh;
    const unsigned ml = __reduce_max_sync(0xffffffffu, in1 ? lo : 0u);
    const bool in2 = in1 && lo == ml;
    const int mi = __reduce_min_sync(0xffffffffu, in2 ? x.i : 0x7fffffff);
    Best r;
    r.v = (long long)(((unsigned long long)(unsigned)mh << 32) | ml);
    r.i = mi;
    return r;
}

// Shared-memory working set of the epilogue.
template <int NMICS, int NBITS, int L>
struct EpiSmem {
    using G = Geo<NBITS, L>;
    static constexpr int P = NMICS * (NMICS - 1) / 2;
    long long curve[P][G::NLAGS_PAD];   // raw sums, index j <-> lag s = j - PADL; later post-Gaussian
    int best[P];
    long long red_v[32];
    int red_i[32];
};

// curve[][] holds the raw correlation sums.  All THREADS threads of the group call; it synchronises the group with
// barrier bar_id (0 and THREADS = blockDim.x: the whole CTA, i.e. __syncthreads; a warp-specialised kernel passes its
// own barrier id and the index of the thread within the group).
// HAVE_BEST: the caller has already found the first-max lags and stored them in e.best[] (and in p.lags); step (1) is skipped.
template <int NMICS, int NBITS, int L, int THREADS, int BAR = 0, bool HAVE_BEST = false>
__device__ __forceinline__ void epilogue(EpiSmem<NMICS, NBITS, L> &e, const float *gauss_s,
                                         const AtFusedParams &p, unsigned long long f, int tid = threadIdx.x, int bar_id = BAR)
{
    using G = Geo<NBITS, L>;
    constexpr int P = NMICS * (NMICS - 1) / 2, NL = G::NL, OFF = G::PADL - L; // curve index of lag -L
    constexpr int NWARPS = THREADS / 32;
    const int lane = tid & 31, warp = tid >> 5;
    auto group_sync = [bar_id] { asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "n"(THREADS) : "memory"); };   // bar_id: group-uniform
#ifdef AT_PROF
    // role 4 of the cycle account: [arg-max, gate + raw, Gaussian, curve stores, tuple scan, reduction + outputs, classes], thread 0 of the group
    unsigned long long ep_t = clock64();
#define EPI_MARK(k) do { if (tid == 0 && p.prof) { const unsigned long long t_ = clock64(); atomicAdd(&p.prof[32 + (k)], t_ - ep_t); ep_t = t_; } } while (0)
#else
#define EPI_MARK(k) do { } while (0)
#endif

    // (1) arg-max per pair (correlations.c:20-23): one warp per pair
    if constexpr (!HAVE_BEST)
    for (int pr = warp; pr < P; pr += NWARPS) {
        Best b = {LLONG_MIN, 0x7fffffff};
        for (int li = lane; li < NL; li += 32) {
            const long long v = e.curve[pr][OFF + li];
            if (v > b.v) { b.v = v; b.i = li; }
        }
        b = warp_best(b);
        if (lane == 0) {
            e.best[pr] = b.i - L;
            if (p.lags) p.lags[f * P + pr] = b.i - L;
        }
    }
    group_sync();
    EPI_MARK(0);
    if (p.gate && tid == 0) {           // sample_compute.h:124-134
        int tot = 0;
        for (int pr = 0; pr < P; pr++) tot += e.best[pr] * e.best[pr];
        p.gate[f] = tot > 4 ? 1 : 0;
    }
    if (p.raw) {
        for (int idx = tid; idx < P * NL; idx += THREADS)
            p.raw[f * (unsigned long long)(P * NL) + idx] = e.curve[idx / NL][OFF + idx % NL];
    }
    const bool need_gauss = p.corr || p.cell || p.highest || p.xy || p.classes;
    if (!need_gauss) return;
    group_sync();   // raw reads done before the in-place re-weighting
    EPI_MARK(1);

    // (2) Gaussian re-weighting in place (correlations.c:26-33)
    for (int idx = tid; idx < P * NL; idx += THREADS) {
        const int pr = idx / NL, li = idx % NL;
        int d = (li - L) - e.best[pr];
        d = d < 0 ? -d : d;
        const float c = __ll2float_rn(e.curve[pr][OFF + li]);
        e.curve[pr][OFF + li] = __float2ll_rz(__fmul_rn(c, gauss_s[d]));
    }
    group_sync();
    EPI_MARK(2);
    if (p.corr) {
        if (p.corr_struct) {   // struct correlations_t [F][P]: 93 x int64, int best_shift, pad, uint64 last_update
            long long *base = reinterpret_cast<long long *>(p.corr) + f * (unsigned long long)(P * (NL + 2));
            for (int idx = tid; idx < P * (NL + 2); idx += THREADS) {
                const int pr = idx / (NL + 2), k = idx % (NL + 2);
                long long v;
                if (k < NL) v = e.curve[pr][OFF + k];
                else if (k == NL) v = (long long)(unsigned)e.best[pr];   // best_shift in the low word, pad = 0
                else v = (long long)p.now_us;                            // correlations.c:35
                base[idx] = v;
            }
        } else {
            long long *base = reinterpret_cast<long long *>(p.corr) + f * (unsigned long long)(P * NL);
            for (int idx = tid; idx < P * NL; idx += THREADS) base[idx] = e.curve[idx / NL][OFF + idx % NL];
        }
    }
    EPI_MARK(3);
    if (!(p.cell || p.highest || p.xy || p.classes)) return;

    // (3) likelihood maximum over the distinct LUT tuples (vga_heatmap.h:96-108); candidates are
    //     sorted by their first row-major cell, so "lowest candidate index" == "first cell".
    Best b = {LLONG_MIN, 0x7fffffff};
    if (P > 6 && p.cand_row32) {   // many pairs: a candidate's tuple is one 32-byte row, two 16-byte loads
        for (int c = tid; c < p.n_cand; c += THREADS) {
            const uint4 r0 = __ldg(&p.cand_row32[2 * c]), r1 = __ldg(&p.cand_row32[2 * c + 1]);
            const uint32_t w[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
            long long like = 0;
#pragma unroll
            for (int pr = 0; pr < P; pr++) like += e.curve[pr][OFF + ((w[pr >> 2] >> (8 * (pr & 3))) & 0xFFu)];
            if (like > b.v) { b.v = like; b.i = c; }
        }
    } else {   // U candidates per turn: their tuple bytes are all requested before the first dependent curve gather
        constexpr int U = P <= 6 ? 4 : 2;
        int c = tid;
        for (; c + (U - 1) * THREADS < p.n_cand; c += U * THREADS) {
            int ix[U][P];
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
                for (int pr = 0; pr < P; pr++) ix[u][pr] = __ldg(&p.cand_idx[pr * p.n_cand + c + u * THREADS]);
#pragma unroll
            for (int u = 0; u < U; u++) {
                long long like = 0;
#pragma unroll
                for (int pr = 0; pr < P; pr++) like += e.curve[pr][OFF + ix[u][pr]];
                if (like > b.v) { b.v = like; b.i = c + u * THREADS; }
            }
        }
        for (; c < p.n_cand; c += THREADS) {
            long long like = 0;
#pragma unroll
            for (int pr = 0; pr < P; pr++) like += e.curve[pr][OFF + __ldg(&p.cand_idx[pr * p.n_cand + c])];
            if (like > b.v) { b.v = like; b.i = c; }
        }
    }
    EPI_MARK(4);
    b = warp_best(b);
    if (lane == 0) { e.red_v[warp] = b.v; e.red_i[warp] = b.i; }
    group_sync();
    if (warp == 0) {
        Best r = {LLONG_MIN, 0x7fffffff};
        if (lane < NWARPS) { r.v = e.red_v[lane]; r.i = e.red_i[lane]; }
        r = warp_best(r);
        if (lane == 0) {
            const int4 cxy = __ldg(&p.cand_cxy[r.i]);                // {first cell of the tuple, x, y}: vga_heatmap.h:52-53, tabulated
            e.red_v[0] = r.v;
            if (p.cell) p.cell[f] = cxy.x;
            if (p.highest) p.highest[f] = r.v;
            if (p.xy) reinterpret_cast<float2 *>(p.xy)[f] = make_float2(__int_as_float(cxy.y), __int_as_float(cxy.z));
        }
    }
    EPI_MARK(5);
    if (p.classes) {   // vga_heatmap.h:111-126, colour codes of lib/vga/vga16_graphics.h:31-34
        group_sync();
        const long long top = e.red_v[0];
        const long long tw = (top * 63) >> 6, tg = (top * 31) >> 5, tr = (top * 15) >> 4, tb = (top * 7) >> 3;
        for (int c = tid; c < p.n_cells; c += THREADS) {
            long long like = 0;
#pragma unroll
            for (int pr = 0; pr < P; pr++) like += e.curve[pr][OFF + p.lut[pr * p.n_cells + c]];
            p.classes[f * (unsigned long long)p.n_cells + c] =
                like >= tw ? 15 : like >= tg ? 3 : like >= tr ? 8 : like >= tb ? 5 : 0;
        }
    }
}

} // namespace atk
