// at_umma_common.cuh -- tcgen05 (UMMA / TMEM) wrappers and the balanced-digit frame preparation shared by the
// tcgen05 localization kernels (at_fused_umma.cu, at_fused_umma_m.cu).  Descriptor layouts follow
// cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor) of the CUTLASS tree shipped with this image.
#pragma once
#include "at_imma_common.cuh"

namespace atk {

// ---------------------------------------------------------------- tcgen05 wrappers
// shared-memory matrix descriptor, no swizzle, version 1; lo word = start address and LBO (16-byte units)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)(((saddr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16)) | ((uint64_t)((sbo_bytes >> 4) | 0x4000u) << 32);
}
// instruction descriptor: S32 accumulators, signed int8 A and B, both MN-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc(int n)
{
    return (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// same with the descriptors given as (lo, hi) words: lo = start address and LBO, the only part that changes per MMA
__device__ __forceinline__ void umma_i8_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
                 "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, {%7, %7, %7, %7}, p;\n\t}\n"
                 :: "r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads)
{
    asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ int dp2a_lo_acc(uint32_t a, uint32_t b, int c)
{
    int r;
    asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ int dp2a_hi_acc(uint32_t a, uint32_t b, int c)
{
    int r;
    asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// 16 consecutive samples -> balanced digit planes.  As imma_prep16, but the product carries +0x8000 (the IDP addend,
// free): q = a * 2W + 0x8000, so byte 2 of q is h = (w + 128) >> 8 and byte 1 of q is l with its top bit flipped.
// ref: rolling_buffer.c:66, buffer.c:16, buffer.c:8-9.
__device__ __forceinline__ void umma_prep16(const uint32_t (&rw)[4], int mean, const uint32_t *win2, int i0,
                                            uint32_t (&hi)[4], uint32_t (&lo)[4])
{
    const uint32_t k4 = (uint32_t)((256 - mean) & 0xFF) * 0x01010101u;
    const uint32_t k7 = k4 & 0x7F7F7F7Fu, kM = k4 & 0x80808080u;
#pragma unroll
    for (int w4 = 0; w4 < 4; w4++) {
        const uint4 ww = *reinterpret_cast<const uint4 *>(&win2[imma_win_index(i0 + 4 * w4)]);
        const uint32_t d = sub_bytes(rw[w4], k7, kM);
        const int p0 = dp2a_lo_acc(ww.x, d, 0x8000), p1 = dp2a_lo_acc(ww.y, d, 0x8000);
        const int p2 = dp2a_hi_acc(ww.z, d, 0x8000), p3 = dp2a_hi_acc(ww.w, d, 0x8000);
        const uint32_t t01 = __byte_perm((uint32_t)p0, (uint32_t)p1, 0x6251);
        const uint32_t t23 = __byte_perm((uint32_t)p2, (uint32_t)p3, 0x6251);
        lo[w4] = __byte_perm(t01, t23, 0x5410) ^ 0x80808080u;
        hi[w4] = __byte_perm(t01, t23, 0x7632);
    }
}
// The same from a PACKED window table (half the shared-memory reads): word j of a 16-sample chunk c holds 2W of samples
// 2j (low half) and 2j + 1 (high half); the two halves of every chunk are stored in two arrays so that the lanes' 16-byte
// reads are contiguous: winp[(h * n / 16 + c) * 4 + (j & 3)], h = j >> 2.  The halves are masked apart here (IDP.2A adds the
// two 16-bit products of a word, so the other half must be zero).
__device__ __forceinline__ void umma_win_fill_packed(uint32_t *winp, const int16_t *window, int n, int tid, int nthreads)
{
    for (int i = tid; i < n / 2; i += nthreads) {
        const int c = i >> 3, j = i & 7;
        winp[((j >> 2) * (n / 16) + c) * 4 + (j & 3)] = ((uint32_t)(2 * (int)window[2 * i]) & 0xFFFFu) | ((uint32_t)(2 * (int)window[2 * i + 1]) << 16);
    }
}
__device__ __forceinline__ void umma_prep16p(const uint32_t (&rw)[4], int mean, const uint32_t *winp, int n, int i0,
                                             uint32_t (&hi)[4], uint32_t (&lo)[4])
{
    const uint32_t k4 = (uint32_t)((256 - mean) & 0xFF) * 0x01010101u;
    const uint32_t k7 = k4 & 0x7F7F7F7Fu, kM = k4 & 0x80808080u;
    const uint4 wa = *reinterpret_cast<const uint4 *>(&winp[(i0 >> 4) * 4]);
    const uint4 wb = *reinterpret_cast<const uint4 *>(&winp[(n / 16 + (i0 >> 4)) * 4]);
    const uint32_t ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
    for (int w4 = 0; w4 < 4; w4++) {
        const uint32_t d = sub_bytes(rw[w4], k7, kM);
        const uint32_t e0 = ww[2 * w4], e1 = ww[2 * w4 + 1];                    // samples (4 w4, 4 w4 + 1), (4 w4 + 2, 4 w4 + 3)
        const int p0 = dp2a_lo_acc(e0 & 0xFFFFu, d, 0x8000), p1 = dp2a_lo_acc(e0 & 0xFFFF0000u, d, 0x8000);
        const int p2 = dp2a_hi_acc(e1 & 0xFFFFu, d, 0x8000), p3 = dp2a_hi_acc(e1 & 0xFFFF0000u, d, 0x8000);
        const uint32_t t01 = __byte_perm((uint32_t)p0, (uint32_t)p1, 0x6251);
        const uint32_t t23 = __byte_perm((uint32_t)p2, (uint32_t)p3, 0x6251);
        lo[w4] = __byte_perm(t01, t23, 0x5410) ^ 0x80808080u;
        hi[w4] = __byte_perm(t01, t23, 0x7632);
    }
}
// the same with the lane's 16 window words (doubled, pre-masked: even samples low half, odd samples high half) in registers
__device__ __forceinline__ void umma_prep16r(const uint32_t (&rw)[4], int mean, const uint32_t (&wr)[16],
                                             uint32_t (&hi)[4], uint32_t (&lo)[4])
{
    const uint32_t k4 = (uint32_t)((256 - mean) & 0xFF) * 0x01010101u;
    const uint32_t k7 = k4 & 0x7F7F7F7Fu, kM = k4 & 0x80808080u;
#pragma unroll
    for (int w4 = 0; w4 < 4; w4++) {
        const uint32_t d = sub_bytes(rw[w4], k7, kM);
        const int p0 = dp2a_lo_acc(wr[4 * w4 + 0], d, 0x8000), p1 = dp2a_lo_acc(wr[4 * w4 + 1], d, 0x8000);
        const int p2 = dp2a_hi_acc(wr[4 * w4 + 2], d, 0x8000), p3 = dp2a_hi_acc(wr[4 * w4 + 3], d, 0x8000);
        const uint32_t t01 = __byte_perm((uint32_t)p0, (uint32_t)p1, 0x6251);
        const uint32_t t23 = __byte_perm((uint32_t)p2, (uint32_t)p3, 0x6251);
        lo[w4] = __byte_perm(t01, t23, 0x5410) ^ 0x80808080u;
        hi[w4] = __byte_perm(t01, t23, 0x7632);
    }
}
// The mirrored chunk: the window is symmetric (W[i] == W[N-1-i]), so sample e of the chunk that mirrors the lane's first one
// uses the window word of sample 15 - e -- whose parity, hence pre-masked half, is the other one.  Swapping the bytes of
// each data half-word (one PRMT per word) puts the byte under the half that holds the window: 16 window registers serve
// both chunks.
__device__ __forceinline__ void umma_prep16m(const uint32_t (&rw)[4], int mean, const uint32_t (&wr)[16],
                                             uint32_t (&hi)[4], uint32_t (&lo)[4])
{
    const uint32_t k4 = (uint32_t)((256 - mean) & 0xFF) * 0x01010101u;
    const uint32_t k7 = k4 & 0x7F7F7F7Fu, kM = k4 & 0x80808080u;
#pragma unroll
    for (int w4 = 0; w4 < 4; w4++) {
        const uint32_t d0 = sub_bytes(rw[w4], k7, kM), d = __byte_perm(d0, d0, 0x2301);
        const int p0 = dp2a_lo_acc(wr[15 - 4 * w4], d, 0x8000), p1 = dp2a_lo_acc(wr[14 - 4 * w4], d, 0x8000);
        const int p2 = dp2a_hi_acc(wr[13 - 4 * w4], d, 0x8000), p3 = dp2a_hi_acc(wr[12 - 4 * w4], d, 0x8000);
        const uint32_t t01 = __byte_perm((uint32_t)p0, (uint32_t)p1, 0x6251);
        const uint32_t t23 = __byte_perm((uint32_t)p2, (uint32_t)p3, 0x6251);
        lo[w4] = __byte_perm(t01, t23, 0x5410) ^ 0x80808080u;
        hi[w4] = __byte_perm(t01, t23, 0x7632);
    }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}

// ---------------------------------------------------------------- diagonal sums of a polyphase tile, in registers
// A 32-row x 16-phase block held one row per lane (tcgen05.ld 32x32b: lane = row, register = phase): entry
// (row, phi) belongs to lag index row - phi.  Five exchange stages (lane ^ 1, 2, 4, 8, 16): in stage k a lane sends the
// odd entries of its array to its partner and keeps the even ones, so that afterwards it holds the lags congruent to
// itself modulo 2^(k+1); the upper lane of a pair sees its partner's entries one slot later (its row is 2^k higher).
// 16 shuffles per block.  Result: n0 = sum for lag index (lane), n1 = partial sum for lag index (lane - 32), non-zero
// for lanes >= 17 only -- the 15 lags whose rows straddle two lane quarters.  (tools/probes/epi_probe.cu checks it.)
template <int N>
struct Arr { int v[N]; };
template <int N, int K>
__device__ __forceinline__ Arr<N / 2 + 1> bfly_stage(const Arr<N> &in, int lane)
{
    constexpr int NE = N / 2, NO = N / 2 + 1;
    const bool upper = (lane >> K) & 1;
    int r[NE];
#pragma unroll
    for (int s = 0; s < NE; s++) r[s] = __shfl_xor_sync(0xffffffffu, in.v[2 * s + 1], 1 << K);
    Arr<NO> out;
#pragma unroll
    for (int s = 0; s < NO; s++) {
        const int base = 2 * s < N ? in.v[2 * s] : 0;
        const int lo = s < NE ? r[s] : 0, up = s >= 1 ? r[s - 1] : 0;
        out.v[s] = base + (upper ? up : lo);
    }
    return out;
}
__device__ __forceinline__ void diag_butterfly(const Arr<16> &a, int lane, int &n0, int &n1)
{
    const Arr<9> b = bfly_stage<16, 0>(a, lane);
    const Arr<5> c = bfly_stage<9, 1>(b, lane);
    const Arr<3> d = bfly_stage<5, 2>(c, lane);
    const Arr<2> e = bfly_stage<3, 3>(d, lane);
    const Arr<2> f = bfly_stage<2, 4>(e, lane);
    n0 = f.v[0]; n1 = f.v[1];
}
// two independent blocks, stages interleaved (twice the shuffles in flight per warp)
__device__ __forceinline__ void diag_butterfly2(const Arr<16> &a, const Arr<16> &a2, int lane, int &n0, int &n1, int &p0, int &p1)
{
    const Arr<9> b = bfly_stage<16, 0>(a, lane), b2 = bfly_stage<16, 0>(a2, lane);
    const Arr<5> c = bfly_stage<9, 1>(b, lane), c2 = bfly_stage<9, 1>(b2, lane);
    const Arr<3> d = bfly_stage<5, 2>(c, lane), d2 = bfly_stage<5, 2>(c2, lane);
    const Arr<2> e = bfly_stage<3, 3>(d, lane), e2 = bfly_stage<3, 3>(d2, lane);
    const Arr<2> f = bfly_stage<2, 4>(e, lane), f2 = bfly_stage<2, 4>(e2, lane);
    n0 = f.v[0]; n1 = f.v[1]; p0 = f2.v[0]; p1 = f2.v[1];
}

// 16 bytes starting sh bytes (1..15) into the 32-byte pair (a, b); sh is warp-uniform (one ring head per frame), so the
// word offset is a branch, not a chain of selects
__device__ __forceinline__ uint4 realign16(const uint4 a, const uint4 b, int sh)
{
    const int ws = sh >> 2, bs = (sh & 3) * 8;
    uint32_t t0, t1, t2, t3, t4;
    if (ws == 0) { t0 = a.x; t1 = a.y; t2 = a.z; t3 = a.w; t4 = b.x; }
    else if (ws == 1) { t0 = a.y; t1 = a.z; t2 = a.w; t3 = b.x; t4 = b.y; }
    else if (ws == 2) { t0 = a.z; t1 = a.w; t2 = b.x; t3 = b.y; t4 = b.z; }
    else { t0 = a.w; t1 = b.x; t2 = b.y; t3 = b.z; t4 = b.w; }
    return make_uint4(__funnelshift_r(t0, t1, bs), __funnelshift_r(t1, t2, bs), __funnelshift_r(t2, t3, bs), __funnelshift_r(t3, t4, bs));
}


} // namespace atk
