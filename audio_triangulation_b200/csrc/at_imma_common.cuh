// at_imma_common.cuh -- device helpers shared by the tensor-core localization kernels
// (at_fused_imma.cu, at_fused_umma.cu): IMMA wrappers, byte-wise prep arithmetic, REDUX reductions and
// the warp-scope epilogue (arg-max bookkeeping, Gaussian re-weighting, bounded likelihood search).
#pragma once
#include <limits.h>

#include "at_fused_common.cuh"

namespace atk {

// D += A * B, m16n8k32, int8 operands with per-operand signedness, int32 accumulate
#define AT_MMA(TA, TB)                                                                                          \
    __device__ __forceinline__ void mma_##TA##_##TB(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])  \
    {                                                                                                           \
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32." #TA "." #TB ".s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, " \
                     "{%8,%9}, {%0,%1,%2,%3};"                                                                  \
                     : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])                                           \
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));                       \
    }
AT_MMA(s8, s8)
AT_MMA(s8, u8)
AT_MMA(u8, s8)
AT_MMA(u8, u8)
#undef AT_MMA

__device__ __forceinline__ uint4 ldg_stream(const uint8_t *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// a = two unsigned 16-bit halves, b = four signed bytes: a.lo*b0 + a.hi*b1 (lo) / a.lo*b2 + a.hi*b3 (hi)
__device__ __forceinline__ int dp2a_lo_u16s8(uint32_t a, uint32_t b)
{
    int r;
    asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(0));
    return r;
}
__device__ __forceinline__ int dp2a_hi_u16s8(uint32_t a, uint32_t b)
{
    int r;
    asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(0));
    return r;
}

// four independent byte additions x + k (mod 256): k7 = low 7 bits of k per byte, kM = top bits of k
__device__ __forceinline__ uint32_t sub_bytes(uint32_t x, uint32_t k7, uint32_t kM)
{
    const uint32_t t = (x & 0x7F7F7F7Fu) + k7;
    return t ^ (x & 0x80808080u) ^ kM;
}

// ---------------------------------------------------------------- frame preparation shared by the tensor kernels
// The doubled window 2*W[i] ((a * 2W) >> 8 == (a * W) >> 7, so the prepared sample's low/high bytes are bytes 1/2 of
// the product) is kept in shared memory one 32-bit word per sample, pre-masked for IDP.2A (even samples in the low
// half, odd samples in the high half) and chunk-interleaved so that lanes working on consecutive 16-sample chunks read
// consecutive 16-byte groups (conflict-free LDS.128).
__device__ __forceinline__ int imma_win_index(int i)
{
    const int c = i >> 4, w = (i >> 2) & 3, e = i & 3;
    return ((((c >> 5) * 4 + w) * 32) + (c & 31)) * 4 + e;
}
__device__ __forceinline__ void imma_win_fill(uint32_t *win2, const int16_t *window, int n, int tid, int nthreads)
{
    for (int i = tid; i < n; i += nthreads) win2[imma_win_index(i)] = (uint32_t)(2 * (int)window[i]) << ((i & 1) * 16);
}
// 16 consecutive samples (chronological index i0, 16-aligned): raw ADC bytes rw[4] -> hi / lo byte-plane words.
// ref: rolling_buffer.c:66 (x - mean), buffer.c:16 (<<8 wraps to 256 * sext8), buffer.c:8-9 (Q15 window).
__device__ __forceinline__ void imma_prep16(const uint32_t (&rw)[4], int mean, const uint32_t *win2, int i0,
                                            uint32_t (&hi)[4], uint32_t (&lo)[4])
{
    // per-byte x - mean (mod 256) = x + k with k = 256 - mean: add the low 7 bits, xor the top bits
    const uint32_t k4 = (uint32_t)((256 - mean) & 0xFF) * 0x01010101u;
    const uint32_t k7 = k4 & 0x7F7F7F7Fu, kM = k4 & 0x80808080u;
#pragma unroll
    for (int w4 = 0; w4 < 4; w4++) {
        const uint4 ww = *reinterpret_cast<const uint4 *>(&win2[imma_win_index(i0 + 4 * w4)]);
        const uint32_t d = sub_bytes(rw[w4], k7, kM);
        // IDP.2A does byte extraction, sign extension and the multiply in one instruction:
        // (u16 pair) . (s8 pair) with one u16 zero selects a single signed byte of d
        const int p0 = dp2a_lo_u16s8(ww.x, d), p1 = dp2a_lo_u16s8(ww.y, d);
        const int p2 = dp2a_hi_u16s8(ww.z, d), p3 = dp2a_hi_u16s8(ww.w, d);
        const uint32_t t01 = __byte_perm((uint32_t)p0, (uint32_t)p1, 0x6251);
        const uint32_t t23 = __byte_perm((uint32_t)p2, (uint32_t)p3, 0x6251);
        lo[w4] = __byte_perm(t01, t23, 0x5410);
        hi[w4] = __byte_perm(t01, t23, 0x7632);
    }
}
// one sample (unaligned ring heads): returns the 24-bit product whose bytes 2 / 1 are the high / low plane bytes
__device__ __forceinline__ int imma_prep1(uint32_t raw_byte, int mean, const uint32_t *win2, int i)
{
    const int a = (int)(signed char)((raw_byte - (uint32_t)mean) & 0xFFu);
    return a * (int)(win2[imma_win_index(i)] >> ((i & 1) * 16));
}

// sqrt(a * b) rounded up (a, b < 2^27): float product and approximate root, both within 2^-20, times 1 + 2^-16
__device__ __forceinline__ float sqrt_prod_up(unsigned a, unsigned b)
{
    return sqrtf((float)a * (float)b) * 1.0000153f + 1.0f;
}

// 64-bit maximum across the warp with two REDUX instructions (high word signed, low word unsigned).
__device__ __forceinline__ long long warp_max_i64(long long key)
{
    const int hi = (int)(key >> 32);
    const unsigned lo = (unsigned)key;
    const int mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    return ((long long)mhi << 32) | (long long)mlo;
}

// (largest value, then lowest cell) across the warp with three REDUX-class reductions
__device__ __forceinline__ Best warp_best_cell(Best x)
{
    const long long top = warp_max_i64(x.v);
    const int cell = __reduce_min_sync(0xffffffffu, x.v == top ? x.i : 0x7fffffff);
    Best r = {top, cell};
    return r;
}

// Peak-tuple look-up (3-pair arrays): when the three first-max lags (b0, b1, b2) form a tuple of the LUT, that tuple's
// likelihood is the sum of the three curve maxima and nothing can tie it once every raw peak is >= 2048 (an entry off
// the peak is <= trunc(fl(peak) * g[1]), g[1] = exp(-1/36) < 0.973, far below trunc(fl(peak))); the tuples are
// distinct, so the first row-major cell of that tuple is the reference's answer (vga_heatmap.h:96-108).  One 16-byte
// load settles cell, xy and highest_L without ever materialising the curves.  peak[] = raw curve maxima.
// Returns false when the frame needs the search of epilogue_warp.
template <int L>
__device__ __forceinline__ bool peak_tuple_lookup(const AtFusedParams &p, unsigned long long f, int lane,
                                                  int b0s, int b1s, int b2s, const long long (&peak)[3])
{
    constexpr int NL = 2 * L + 1;
    if (!p.peak_tab || peak[0] < 2048 || peak[1] < 2048 || peak[2] < 2048) return false;
    const int4 e = __ldg(&p.peak_tab[((b0s + L) * NL + (b1s + L)) * NL + (b2s + L)]);
    if (e.x < 0) return false;
    if (lane == 0) {
        if (p.cell) p.cell[f] = e.x;
        if (p.xy) reinterpret_cast<float2 *>(p.xy)[f] = make_float2(__int_as_float(e.y), __int_as_float(e.z));
        if (p.highest)   // g[0] = 1: the post-Gaussian peak is the float-rounded raw peak (correlations.c:30-31)
            p.highest[f] = __float2ll_rz(__ll2float_rn(peak[0])) + __float2ll_rz(__ll2float_rn(peak[1])) +
                           __float2ll_rz(__ll2float_rn(peak[2]));
        if (p.gate) p.gate[f] = (b0s * b0s + b1s * b1s + b2s * b2s) > 4 ? 1 : 0;   // sample_compute.h:124-134
        if (p.stats) atomicAdd(&p.stats[3], 1ull);
    }
    return true;
}

// Warp-scope epilogue for the optional products; curve[][] holds raw sums indexed by j = s + PAD,
// b0..b2 are the three best shifts (warp-uniform).
// FULL_SCAN = false: when neither the tuple look-up nor a bounded box settles the likelihood maximum the function stores
// nothing for cell / highest / xy and returns false -- the caller then scans all tuples with more threads (gate is written).
template <int L, int PAD, int NJ, int CSTRIDE, bool FULL_SCAN = true>
__device__ __forceinline__ bool epilogue_warp(long long *curve_base, int b0s, int b1s, int b2s, const float *gauss_s,
                                              const AtFusedParams &p, unsigned long long f, int lane)
{
    constexpr int P = 3, NL = 2 * L + 1, OFF = PAD - L;
    // curve(pr, x): raw sums of pair pr, x = lag index j = s + PAD; rows are CSTRIDE int64 apart
#define CURVE(pr, x) curve_base[(pr) * CSTRIDE + (x)]
    const int best[3] = {b0s, b1s, b2s};
    if (p.gate && lane == 0) {                                   // sample_compute.h:124-134
        const int tot = b0s * b0s + b1s * b1s + b2s * b2s;
        p.gate[f] = tot > 4 ? 1 : 0;
    }
    if (p.raw)
        for (int idx = lane; idx < P * NL; idx += 32) p.raw[f * (unsigned long long)(P * NL) + idx] = CURVE(idx / NL, OFF + idx % NL);
    if (!(p.corr || p.cell || p.highest || p.xy || p.classes)) return true;
    // Gaussian re-weighting (correlations.c:26-33): in place when whole curves are wanted, otherwise
    // evaluated on demand for the few entries the bounded likelihood search touches.
    const bool weighted = p.corr || p.classes;
    if (weighted) {
        __syncwarp();
        for (int idx = lane; idx < P * NL; idx += 32) {
            const int pr = idx / NL, li = idx % NL;
            int d = (li - L) - best[pr];
            d = d < 0 ? -d : d;
            CURVE(pr, OFF + li) = __float2ll_rz(__fmul_rn(__ll2float_rn(CURVE(pr, OFF + li)), gauss_s[d]));
        }
        __syncwarp();
    }
    auto post = [&](int pr, int li) -> long long {
        const long long v = CURVE(pr, OFF + li);
        if (weighted) return v;
        int d = (li - L) - (pr == 0 ? b0s : (pr == 1 ? b1s : b2s));
        d = d < 0 ? -d : d;
        return __float2ll_rz(__fmul_rn(__ll2float_rn(v), gauss_s[d]));
    };
    if (p.corr) {
        if (p.corr_struct) {
            long long *base = reinterpret_cast<long long *>(p.corr) + f * (unsigned long long)(P * (NL + 2));
            for (int idx = lane; idx < P * (NL + 2); idx += 32) {
                const int pr = idx / (NL + 2), k = idx % (NL + 2);
                base[idx] = k < NL ? CURVE(pr, OFF + k) : (k == NL ? (long long)(unsigned)best[pr] : (long long)p.now_us);
            }
        } else {
            long long *base = reinterpret_cast<long long *>(p.corr) + f * (unsigned long long)(P * NL);
            for (int idx = lane; idx < P * NL; idx += 32) base[idx] = CURVE(idx / NL, OFF + idx % NL);
        }
    }
    if (!(p.cell || p.highest || p.xy || p.classes)) return true;
    // vga_heatmap.h:96-108: maximum of L = sum_pairs CURVE(pair, lut) over the distinct LUT tuples,
    // first row-major cell on ties.  Exact bounded search: every entry of curve p is <= Pmax_p =
    // max(CURVE(p, best_p), 0), and an entry at distance >= r from the peak is <= trunc(peak * g[r])
    // (the re-weighting is monotone), so tuples outside a box around (best_0, best_1) cannot reach
    // a likelihood already found inside it once  bound(r+1) + sum(other Pmax) < that likelihood.
    Best b = {LLONG_MIN, 0x7fffffff};   // .i holds the CELL index here (lower cell wins ties)
    {
        const int b0 = b0s + L, b1 = b1s + L;
        long long pmax[P];
        long long others0 = 0, others1 = 0;
#pragma unroll
        for (int pr = 0; pr < P; pr++) {
            const long long v = post(pr, best[pr] + L);
            pmax[pr] = v > 0 ? v : 0;
            if (pr != 0) others0 += pmax[pr];
            if (pr != 1) others1 += pmax[pr];
        }
        const float pk0 = __ll2float_rn(pmax[0]), pk1 = __ll2float_rn(pmax[1]);   // exact: both came from floats
        auto scan_box = [&](int r0, int r1) {
            Best bb = {LLONG_MIN, 0x7fffffff};
            const int w1 = 2 * r1 + 1, cells = (2 * r0 + 1) * w1;
            for (int q = lane; q < cells; q += 32) {
                const int i0 = b0 - r0 + q / w1, i1 = b1 - r1 + q % w1;
                if (i0 < 0 || i0 >= NL || i1 < 0 || i1 >= NL) continue;
                const int lo = p.cs_grid[i0 * NL + i1], hi = p.cs_grid[i0 * NL + i1 + 1];
                if (lo == hi) continue;
                const long long base01 = post(0, i0) + post(1, i1);
                for (int c = lo; c < hi; c++) {
                    const long long like = base01 + post(2, p.cs_idx[2 * p.n_cand + c]);
                    const int cell = p.cs_cell[c];
                    if (like > bb.v || (like == bb.v && cell < bb.i)) { bb.v = like; bb.i = cell; }
                }
            }
            return warp_best_cell(bb);
        };
        auto bound = [&](float pk, int r) -> long long { return r <= 2 * L ? __float2ll_rz(__fmul_rn(pk, gauss_s[r])) : 0; };
        constexpr int R_FIRST = 2, R_MAX = 12;
        int how = 0;
        bool ok = false;
        // Consistent peaks: if the LUT holds the tuple (best_0, best_1, best_2) itself, its likelihood is the sum of
        // the three curve maxima.  No other tuple can tie it when every peak is >= 1024: an entry off the peak is
        // <= trunc(fl(peak) * g[1]) with g[1] = exp(-1/36) < 0.973, at least 27 below trunc(fl(peak)).  The tuples
        // are distinct, so the hit is unique and cs_cell holds its first row-major cell.
        if (pmax[0] >= 1024 && pmax[1] >= 1024 && pmax[2] >= 1024) {
            const int lo = p.cs_grid[b0 * NL + b1], hi = p.cs_grid[b0 * NL + b1 + 1];
            const int want2 = b2s + L;
            for (int c0 = lo; c0 < hi && !ok; c0 += 32) {
                const int c = c0 + lane;
                const unsigned hit = __ballot_sync(0xffffffffu, c < hi && p.cs_idx[2 * p.n_cand + c] == want2);
                if (hit) {
                    b.v = pmax[0] + pmax[1] + pmax[2];
                    b.i = p.cs_cell[c0 + __ffs(hit) - 1];
                    ok = true; how = 3;
                }
            }
        }
        if (!ok) {
            b = scan_box(R_FIRST, R_FIRST);
            ok = b.v > bound(pk0, R_FIRST + 1) + others0 && b.v > bound(pk1, R_FIRST + 1) + others1;
        }
        if (!ok && !FULL_SCAN) return false;     // the caller scans with more threads than this warp has
        if (!ok && b.v != LLONG_MIN) {          // widen: smallest radii whose outside bound is below what we hold
            int r0 = -1, r1 = -1;
            for (int r = R_FIRST; r <= R_MAX && (r0 < 0 || r1 < 0); r++) {
                if (r0 < 0 && b.v > bound(pk0, r + 1) + others0) r0 = r;
                if (r1 < 0 && b.v > bound(pk1, r + 1) + others1) r1 = r;
            }
            if (r0 >= 0 && r1 >= 0) { b = scan_box(r0, r1); ok = true; how = 1; }
        }
        if (!ok) {                               // flat or inconsistent curves: scan every tuple
            how = 2;
            b.v = LLONG_MIN; b.i = 0x7fffffff;
            for (int c = lane; c < p.n_cand; c += 32) {
                long long like = 0;
#pragma unroll
                for (int pr = 0; pr < P; pr++) like += post(pr, p.cs_idx[pr * p.n_cand + c]);
                const int cell = p.cs_cell[c];
                if (like > b.v || (like == b.v && cell < b.i)) { b.v = like; b.i = cell; }
            }
            b = warp_best_cell(b);
        }
        if (p.stats && lane == 0) atomicAdd(&p.stats[how], 1ull);
    }
    if (lane == 0) {
        const int cellidx = b.i;
        if (p.cell) p.cell[f] = cellidx;
        if (p.highest) p.highest[f] = b.v;
        if (p.xy) reinterpret_cast<float2 *>(p.xy)[f] = p.cell_xy[cellidx];   // vga_heatmap.h:52-53, tabulated per cell
    }
    if (p.classes) {                                             // vga_heatmap.h:111-126
        const long long top = b.v;
        const long long tw = (top * 63) >> 6, tg = (top * 31) >> 5, tr = (top * 15) >> 4, tb = (top * 7) >> 3;
        for (int c = lane; c < p.n_cells; c += 32) {
            long long like = 0;
#pragma unroll
            for (int pr = 0; pr < P; pr++) like += CURVE(pr, OFF + p.lut[pr * p.n_cells + c]);
            p.classes[f * (unsigned long long)p.n_cells + c] = like >= tw ? 15 : like >= tg ? 3 : like >= tr ? 8 : like >= tb ? 5 : 0;
        }
    }
    return true;
}
#undef CURVE


} // namespace atk
