// at_fused_imma.cu -- fused localization kernel, int8 tensor-core form (AT_KERNEL_IMMA).
//
// The lagged cross-correlation of one microphone pair (ref: components/correlations.c:9-18),
//     corr[s] = sum_i x[i] * y[i+s],   s in [-L, L],
// is a Toeplitz x Hankel matrix product once the lag index j = s + PAD is split as j = 8 r + n:
//     D[r][n] = sum_p  A[r][p] * B[p][n],   A[r][p] = y~[p + 8 r - PAD]  (Hankel, row stride 8)
//                                           B[p][n] = x~[p - n]          (Toeplitz)
// with x~, y~ the zero-extended frames and p running over N + 8 samples.  That is exactly one
// mma.sync m16n8k32 tile (16 x 8 = 128 lags, 93 used) per 32 samples.  int16 operands are split
// into a signed high byte and an unsigned low byte, w = 256 h + l, giving four int8 products
//     corr = 65536 * (h.h) + 256 * (h.l + l.h) + (l.l)
// whose int32 accumulators cannot overflow over a frame (|h.h| <= 2^24, |h.l + l.h| < 2^26,
// l.l < 2^26 at N = 1024; 4x that at N = 4096), recombined in int64: bit-identical to the
// reference's int64 accumulation.
//
// Mapping: ONE WARP PER FRAME.  A warp loads its frame (coalesced 16-byte global loads), removes
// DC, applies <<8 and the window (imma_prep16), writes hi/lo byte planes of the three channels to
// its private shared-memory slice, runs 33 k-steps x 12 IMMA (3 pairs x {hh, hl, lh, ll}) with
// fragments loaded straight from the planes (A: four 32-bit loads per plane, the Hankel rows overlap
// in memory and nothing is materialised; B: three aligned 32-bit loads + two funnel shifts for the
// per-column byte offset), then finds the three first-max lags with REDUX reductions and, if asked,
// runs the warp-scope epilogue (Gaussian re-weighting, bounded likelihood search).  No block-level
// synchronisation after start-up.  Why the loop looks the way it does: DESIGN.md section 4.1.
//
// Two instantiations per shape.  PRUNE = false is the kernel just described; it serves requests for whole curves
// (raw / corr / classes / highest).  PRUNE = true serves lags / cell / xy / gate: it first runs nine of the twelve
// products (no l.l; imma_kloop_reuse), certifies the arg-max with a Cauchy-Schwarz bound on the missing product and
// settles the position by the peak-tuple look-up; frames it cannot certify get the l.l product from a second pass and
// continue exactly as above, so every result stays bit-identical to the reference.
#include <limits.h>
#include <stdlib.h>

#include "at_imma_common.cuh"

namespace atk {

template <int NBITS, int L>
struct ImmaGeo {
    static constexpr int N = 1 << NBITS;
    static constexpr int PAD = round_up(L, 16);                 // lag index j = s + PAD; 48 for L = 46 / 44
    static constexpr int KSTEPS = ceil_div(N + 8, 32);          // 33: p runs over [0, N + 8)
    static constexpr int PLANE = round_up(32 * KSTEPS + 8 * 15 + 8, 16);   // bytes per byte-plane: 1184
    static constexpr int NJ = 96;                               // lag slots kept in the epilogue scratch (rows 0..11)
    static constexpr int NL = 2 * L + 1;
    static_assert(PAD + L < NJ, "lag range does not fit rows 0..11 of the 16 x 8 tile");
};

template <int NBITS, int L, int WARPS>
struct ImmaSmem {
    using G = ImmaGeo<NBITS, L>;
    // doubled, pre-masked, chunk-interleaved window (layout: imma_win_index in at_imma_common.cuh)
    alignas(16) uint32_t win2[G::N];
    float gauss[2 * L + 1];
    alignas(16) uint8_t plane[WARPS][3][2][G::PLANE]; // [warp][channel][hi, lo]
    // The epilogue's int64 curves (3 x NJ x 8 bytes) reuse the DATA region [PAD, PAD + N) of this
    // warp's byte planes 0..2 once the MMA loop is done; the zero pads are never touched.
    static_assert(G::NJ * 8 <= G::N && G::PAD % 8 == 0, "curve scratch must fit one plane's data region");
};

// B fragment (32 x 8, column n = x shifted by n bytes): three aligned words + funnel shift
__device__ __forceinline__ void load_b(uint32_t (&b)[2], const uint8_t *plane_k0_bal, int bsh)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(plane_k0_bal);
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    b[0] = __funnelshift_r(w0, w1, bsh);
    b[1] = __funnelshift_r(w1, w2, bsh);
}

// The k-loop.  PRODUCTS selects the digit products it accumulates: 3 = all four (12 IMMA per step), 1 = all but l.l
// (9 per step), 2 = l.l alone (3 per step).  A fragments: four 32-bit loads per y-plane, each landing directly in its
// fragment register.  On this GPU every ALU instruction issued next to an IMMA costs issue time, so
// what counts is the instruction total: fusing the loads into 64-bit ones or re-using the Hankel overlap across k-steps
// both need register moves that cost more than the loads they save.  ya_4 = ya_s + 4 comes from a kernel parameter so
// that ptxas cannot prove two loads adjacent and fuse them.
template <int PLANE, int UNROLL, int PRODUCTS, int ROW2>     // ROW2: byte distance between the data of MMA rows g and g+8
__device__ __forceinline__ void imma_kloop(int (&acc)[3][3][4], uint32_t ya_s, uint32_t ya_4, const uint8_t *xb, int bsh, int nsteps)
{
    constexpr bool MAIN = (PRODUCTS & 1) != 0, LL = (PRODUCTS & 2) != 0;
#pragma unroll UNROLL
    for (int ks = 0; ks < nsteps; ks++) {
        const int k0 = 32 * ks;
        uint32_t Y[4][4], Xah[2], Xal[2], Xbh[2], Xbl[2];
#pragma unroll
        for (int q = 0; q < 4; q++) {          // Y[0] = b.hi, Y[1] = b.lo, Y[2] = c.hi, Y[3] = c.lo
            if (!MAIN && !(q & 1)) continue;   // l.l alone: the hi planes are not read
            const uint32_t off = (2 + q) * PLANE + k0;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(Y[q][0]) : "r"(ya_s + off));        // row g,   k 0..3
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(Y[q][2]) : "r"(ya_4 + off));        // row g,   k 4..7
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(Y[q][1]) : "r"(ya_s + off + ROW2));   // row g+8, k 0..3
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(Y[q][3]) : "r"(ya_4 + off + ROW2));   // row g+8, k 4..7
        }
        if (MAIN) load_b(Xah, xb + 0 * PLANE + k0, bsh);
        load_b(Xal, xb + 1 * PLANE + k0, bsh);
        if (MAIN) load_b(Xbh, xb + 2 * PLANE + k0, bsh);
        load_b(Xbl, xb + 3 * PLANE + k0, bsh);
        if (PRODUCTS == 3) {
            mma_s8_s8(acc[0][0], Y[0], Xah); mma_s8_u8(acc[0][1], Y[0], Xal); mma_u8_u8(acc[0][2], Y[1], Xal);
            mma_s8_s8(acc[1][0], Y[2], Xah); mma_s8_u8(acc[1][1], Y[2], Xal); mma_u8_u8(acc[1][2], Y[3], Xal);
            mma_s8_s8(acc[2][0], Y[2], Xbh); mma_s8_u8(acc[2][1], Y[2], Xbl); mma_u8_u8(acc[2][2], Y[3], Xbl);
            mma_u8_s8(acc[0][1], Y[1], Xah); mma_u8_s8(acc[1][1], Y[3], Xah); mma_u8_s8(acc[2][1], Y[3], Xbh);
        } else if (MAIN) {
            mma_s8_s8(acc[0][0], Y[0], Xah); mma_s8_u8(acc[0][1], Y[0], Xal);
            mma_s8_s8(acc[1][0], Y[2], Xah); mma_s8_u8(acc[1][1], Y[2], Xal);
            mma_s8_s8(acc[2][0], Y[2], Xbh); mma_s8_u8(acc[2][1], Y[2], Xbl);
            mma_u8_s8(acc[0][1], Y[1], Xah); mma_u8_s8(acc[1][1], Y[3], Xah); mma_u8_s8(acc[2][1], Y[3], Xbh);
        } else if (LL) {
            mma_u8_u8(acc[0][2], Y[1], Xal); mma_u8_u8(acc[1][2], Y[3], Xal); mma_u8_u8(acc[2][2], Y[3], Xbl);
        }
    }
}

// The nine-product loop with the Hankel overlap re-used and no register moves.  Every step loads ONE new 8-byte block
// per plane into the half of the fragment quad whose data just expired; on odd steps the tile therefore has its row
// halves exchanged and the product goes to a second accumulator set, folded back after the loop (c ^ 2 exchanges the
// rows of a fragment).  20 instead of 28 shared-memory wavefronts per step: with 9 IMMA per step the loop is bound by
// the shared-memory data pipe, not by the tensor pipe (profiles/r1_kloop_experiments.md).
template <int PLANE, int KSTEPS>
__device__ __forceinline__ void imma_kloop_reuse(int (&acc)[3][3][4], uint32_t ya_s, uint32_t ya_4, const uint8_t *xb, int bsh, int nsteps)
{
    static_assert(KSTEPS % 2 == 1, "schedule: one whole tile, then pairs of steps");
    int accS[3][2][4];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
            for (int c = 0; c < 4; c++) accS[a][b][c] = 0;
    uint32_t Y[4][4];
    auto load_half = [&](int slot, int kb) {   // slot 0: regs 0,2 (MMA row g); slot 1: regs 1,3 (MMA row g+8)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t off = (2 + q) * PLANE + kb;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(Y[q][slot]) : "r"(ya_s + off));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(Y[q][slot + 2]) : "r"(ya_4 + off));
        }
    };
    auto mma_step = [&](int (&h0)[4], int (&m0)[4], int (&h1)[4], int (&m1)[4], int (&h2)[4], int (&m2)[4], int k0) {
        uint32_t Xah[2], Xal[2], Xbh[2], Xbl[2];
        load_b(Xah, xb + 0 * PLANE + k0, bsh); load_b(Xal, xb + 1 * PLANE + k0, bsh);
        load_b(Xbh, xb + 2 * PLANE + k0, bsh); load_b(Xbl, xb + 3 * PLANE + k0, bsh);
        mma_s8_s8(h0, Y[0], Xah); mma_s8_u8(m0, Y[0], Xal);
        mma_s8_s8(h1, Y[2], Xah); mma_s8_u8(m1, Y[2], Xal);
        mma_s8_s8(h2, Y[2], Xbh); mma_s8_u8(m2, Y[2], Xbl);
        mma_u8_s8(m0, Y[1], Xah); mma_u8_s8(m1, Y[3], Xah); mma_u8_s8(m2, Y[3], Xbh);
    };
    if (nsteps) {
        load_half(0, 0); load_half(1, 32);
        mma_step(acc[0][0], acc[0][1], acc[1][0], acc[1][1], acc[2][0], acc[2][1], 0);
#pragma unroll 2
        for (int s = 1; s < KSTEPS; s += 2) {
            const int k0 = 32 * s;
            load_half(0, k0 + 32);        // block s+1 into the half that held block s-1: rows exchanged
            mma_step(accS[0][0], accS[0][1], accS[1][0], accS[1][1], accS[2][0], accS[2][1], k0);
            load_half(1, k0 + 64);        // block s+2 into the half that held block s: rows in order again
            mma_step(acc[0][0], acc[0][1], acc[1][0], acc[1][1], acc[2][0], acc[2][1], k0 + 32);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[a][b][c] += accS[a][b][c ^ 2];
}

template <int NBITS, int L, int WARPS, int CTAS_PER_SM, int UNROLL, bool PRUNE>
__global__ void __launch_bounds__(WARPS * 32, CTAS_PER_SM) at_fused_imma_kernel(const AtFusedParams p)
{
    using G = ImmaGeo<NBITS, L>;
    using S = ImmaSmem<NBITS, L, WARPS>;
    constexpr int N = G::N, PAD = G::PAD, PLANE = G::PLANE;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    S &s = *reinterpret_cast<S *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    // one-time CTA setup: zero every plane (the pads stay zero), doubled window, Gaussian factors
    for (int i = tid; i < (int)(sizeof(s.plane) / 16); i += WARPS * 32)
        reinterpret_cast<uint4 *>(&s.plane[0][0][0][0])[i] = make_uint4(0, 0, 0, 0);
    imma_win_fill(s.win2, p.window, N, tid, WARPS * 32);
    for (int i = tid; i < 2 * L + 1; i += WARPS * 32) s.gauss[i] = p.gauss[i];
    __syncthreads();

    uint8_t *const pl = &s.plane[warp][0][0][0];
    auto plane = [&](int ch, int hl) -> uint8_t * { return pl + (ch * 2 + hl) * PLANE; };
    // MMA rows (g, g+8) carry the lag rows (g, g+8) -- or, in the kernel with the certified shortcut, (rho, rho+4) with
    // rho = g + 4 (g / 4): 32 bytes = one k-step apart, so that the data of MMA row g+8 at step s is the data of MMA row g
    // at step s+1 (imma_kloop_reuse).  The plain loop is 4 % faster with the first mapping, hence both.
    const int row_lo = PRUNE ? g + 4 * (g >> 2) : g, row_hi = PRUNE ? row_lo + 4 : g + 8;
    const int aoff = 8 * t + 8 * row_lo;            // A: plane index of (MMA row g, k = 8t) at k0 = 0
    const int boff = 8 * t - g + PAD;               // B: plane index of (k = 8t, column g)
    const int bal = boff & ~3, bsh = (boff & 3) * 8;
    const bool extras = p.gate || p.raw || p.corr || p.cell || p.highest || p.xy || p.classes;

    // the certified 9-product shortcut (below) serves requests for lags / cell / xy / gate only and needs the peak-tuple table
    const bool prune_ok = PRUNE && p.peak_tab && !(p.raw || p.corr || p.classes || p.highest) && !p.debug_skip;
    const unsigned long long stride = (unsigned long long)gridDim.x * WARPS;
    for (unsigned long long f = (unsigned long long)blockIdx.x * WARPS + warp; f < p.n_frames; f += stride) {
        const uint8_t *src = p.adc + f * (unsigned long long)(3 * N);
        const int head = p.heads ? (p.heads[f] & (N - 1)) : 0;
        const bool prune = prune_ok && (head & 15) == 0;
        unsigned sl[3] = {0, 0, 0};          // sum of squared low digits of each channel (certified shortcut only)

        // ---- channel sums -> floor mean (rolling_buffer.c:48-64); the sum is rotation invariant
        constexpr int Q = N / 512;               // 16-byte loads per lane and channel
        constexpr bool KEEP = Q <= 2;            // N = 1024: the whole frame stays in 24 registers
        uint4 raw[KEEP ? 3 * Q : 1];
        int mean[3];
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            unsigned sum = 0;
#pragma unroll
            for (int q = 0; q < Q; q++) {
                const uint4 v = KEEP ? ldg_stream(src + ch * N + q * 512 + lane * 16)
                                     : __ldg(reinterpret_cast<const uint4 *>(src + ch * N + q * 512 + lane * 16));
                if (KEEP) raw[ch * Q + q] = v;
                sum = __dp4a(v.x, 0x01010101u, sum); sum = __dp4a(v.y, 0x01010101u, sum);
                sum = __dp4a(v.z, 0x01010101u, sum); sum = __dp4a(v.w, 0x01010101u, sum);
            }
            sum = __reduce_add_sync(0xffffffffu, sum);
            mean[ch] = (int)(sum >> NBITS);
        }

        if (!(p.debug_skip & 1))
        // ---- DC removal, <<8, window -> hi / lo byte planes (rolling_buffer.c:66, buffer.c:16, :8-9)
        //      (int16)((b - mean) << 8) = 256 * sext8(b - mean);  ((256 a) * W) >> 15 = (a * 2W) >> 8
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
#pragma unroll
            for (int q = 0; q < Q; q++) {
                const int j0 = q * 512 + lane * 16;                   // ring slot of this lane's 16 bytes
                const uint4 v = KEEP ? raw[ch * Q + q] : __ldg(reinterpret_cast<const uint4 *>(src + ch * N + j0));
                const uint32_t rw[4] = {v.x, v.y, v.z, v.w};
                if ((head & 15) == 0) {
                    const int i0 = (j0 - head) & (N - 1);             // chronological index, 16-aligned
                    uint32_t hi[4], lo[4];
                    imma_prep16(rw, mean[ch], s.win2, i0, hi, lo);
                    AT_CHECK(i0 >= 0 && PAD + i0 + 16 <= PLANE);
                    *reinterpret_cast<uint4 *>(plane(ch, 0) + PAD + i0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4 *>(plane(ch, 1) + PAD + i0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    if (PRUNE && prune) {
#pragma unroll
                        for (int w4 = 0; w4 < 4; w4++) sl[ch] = __dp4a(lo[w4], lo[w4], sl[ch]);
                    }
                } else {   // ring head not 16-aligned: scalar stores (rare; capture heads are arbitrary)
#pragma unroll
                    for (int e = 0; e < 16; e++) {
                        const int i = (j0 + e - head) & (N - 1);
                        const int pr = imma_prep1(rw[e >> 2] >> (8 * (e & 3)), mean[ch], s.win2, i);
                        plane(ch, 0)[PAD + i] = (uint8_t)(pr >> 16);
                        plane(ch, 1)[PAD + i] = (uint8_t)(pr >> 8);
                    }
                }
            }
            if (PRUNE && prune) sl[ch] = __reduce_add_sync(0xffffffffu, sl[ch]);
            if (p.power) {   // rolling_buffer.c:68-70: power of the DC-removed samples (9-bit differences)
                long long acc = 0;
                for (int k = lane; k < N; k += 32) { const int dv = (int)src[ch * N + k] - mean[ch]; acc += (long long)dv * dv; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) p.power[f * 3 + ch] = acc;
            }
        }
        __syncwarp();
        if (p.windowed)
            for (int idx = lane; idx < 3 * N; idx += 32) {
                const int ch = idx / N, i = idx % N;
                p.windowed[f * (unsigned long long)(3 * N) + idx] =
                    (int16_t)(((int)(signed char)plane(ch, 0)[PAD + i] << 8) | plane(ch, 1)[PAD + i]);
            }

        // next frame of this warp -> L1/L2 while the tensor pipe works (takes HBM latency off the chain)
        if (f + stride < p.n_frames) {
            const uint8_t *nx = p.adc + (f + stride) * (unsigned long long)(3 * N) + lane * 128;
            if (lane * 128 < 3 * N) asm volatile("prefetch.global.L1 [%0];" ::"l"(nx));
            if (3 * N > 4096 && lane * 128 + 4096 < 3 * N) asm volatile("prefetch.global.L1 [%0];" ::"l"(nx + 4096));
        }
        // ---- 33 k-steps x 12 IMMA: pairs (a,b), (a,c), (b,c); x = first, y = second mic
        int acc[3][3][4];
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[a][b][c] = 0;
        const uint8_t *ya = pl + aoff, *xb = pl + bal;   // + (ch*2+hl)*PLANE + k0
        const uint32_t ya_s = smem_u32(ya);
        const uint32_t ya_4 = ya_s + (uint32_t)p.opaque_four;
        const int nsteps = (p.debug_skip & 2) ? 0 : G::KSTEPS;   // profiling knob, kept out of the loop body
        if (PRUNE && prune) {
            // ---- certified shortcut.  corr = C9 + LL with C9 = 65536 HH + 256 (HL + LH) and, the low digits being
            //      unsigned, 0 <= LL[s] <= sqrt(Sl_x Sl_y) =: B by Cauchy-Schwarz (Sl = sum of squared low digits of a
            //      channel, accumulated during prep).  If C9 puts its maximum more than B above every other lag, that
            //      lag is THE arg-max of the exact curve (no tie possible); its peak is >= C9's, so the peak-tuple
            //      look-up applies when C9_max >= 2048.  9 instead of 12 IMMA per k-step.  Anything else -- small
            //      gaps, other outputs, no LUT tuple -- adds the l.l product below and takes the exact path.
            imma_kloop_reuse<PLANE, G::KSTEPS>(acc, ya_s, ya_4, xb, bsh, nsteps);
            int b3[3];
            bool sure = true;
#pragma unroll
            for (int pr = 0; pr < 3; pr++) {
                long long c9[4], key = LLONG_MIN, second = LLONG_MIN;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int j = 8 * ((i >> 1) ? row_hi : row_lo) + 2 * t + (i & 1);
                    c9[i] = 65536LL * acc[pr][0][i] + 256LL * acc[pr][1][i];
                    const long long k = c9[i] * 128 + (127 - j);
                    if (j >= PAD - L && j <= PAD + L && k > key) key = k;
                }
                key = warp_max_i64(key);
                const int j1 = 127 - (int)(key & 127);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int j = 8 * ((i >> 1) ? row_hi : row_lo) + 2 * t + (i & 1);
                    if (j >= PAD - L && j <= PAD + L && j != j1 && c9[i] > second) second = c9[i];
                }
                second = warp_max_i64(second);
                const int xc = pr == 2 ? 1 : 0, yc = pr == 0 ? 1 : 2;
                const long long bound = (long long)sqrt_prod_up(sl[xc], sl[yc]) + 1;
                sure = sure && (key >> 7) - second > bound && (key >> 7) >= 2048;
                b3[pr] = j1 - PAD;
            }
            if (sure) {
                const int4 e = __ldg(&p.peak_tab[((b3[0] + L) * (2 * L + 1) + (b3[1] + L)) * (2 * L + 1) + (b3[2] + L)]);
                if (e.x >= 0 || !(p.cell || p.xy)) {
                    if (lane < 3 && p.lags) p.lags[f * 3 + lane] = lane == 0 ? b3[0] : (lane == 1 ? b3[1] : b3[2]);
                    if (lane == 0) {
                        if (p.cell) p.cell[f] = e.x;
                        if (p.xy) reinterpret_cast<float2 *>(p.xy)[f] = make_float2(__int_as_float(e.y), __int_as_float(e.z));
                        if (p.gate) p.gate[f] = (b3[0] * b3[0] + b3[1] * b3[1] + b3[2] * b3[2]) > 4 ? 1 : 0;   // sample_compute.h:124-134
                        if (p.stats) { atomicAdd(&p.stats[3], 1ull); atomicAdd(&p.stats[4], 1ull); }
                    }
                    __syncwarp();   // planes are rewritten by the next frame
                    continue;
                }
            }
            imma_kloop<PLANE, UNROLL, 2, 32>(acc, ya_s, ya_4, xb, bsh, nsteps);
        } else {
            imma_kloop<PLANE, UNROLL, 3, PRUNE ? 32 : 64>(acc, ya_s, ya_4, xb, bsh, nsteps);
        }

        __syncwarp();   // every lane is done reading the planes: their data regions now hold the curves
        constexpr int CSTRIDE = PLANE / 8;                       // int64 stride between planes 0, 1, 2
        long long *const curve_base = reinterpret_cast<long long *>(pl + PAD);
        // ---- recombine in int64, arg-max per pair (correlations.c:20-23): key = value * 128 + (127 - j),
        //      so the 64-bit maximum is the largest value and, among equals, the lowest lag
        int best3[3];
        long long peak[3];
#pragma unroll
        for (int pr = 0; pr < 3; pr++) {
            long long key = LLONG_MIN;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int j = 8 * ((i >> 1) ? row_hi : row_lo) + 2 * t + (i & 1);
                const long long v = 65536LL * acc[pr][0][i] + 256LL * acc[pr][1][i] + (long long)acc[pr][2][i];
                const long long k = v * 128 + (127 - j);
                if (j >= PAD - L && j <= PAD + L && k > key) key = k;
            }
            key = warp_max_i64(key);
            best3[pr] = 127 - (int)(key & 127) - PAD;
            peak[pr] = key >> 7;
        }
        if (lane < 3 && p.lags) p.lags[f * 3 + lane] = lane == 0 ? best3[0] : (lane == 1 ? best3[1] : best3[2]);
        // position products only (cell / xy / highest / gate): consistent peaks are settled by one table load
        bool settled = !extras;
        if (extras && !(p.raw || p.corr || p.classes))
            settled = peak_tuple_lookup<L>(p, f, lane, best3[0], best3[1], best3[2], peak);
        if (!settled) {   // whole curves wanted, or peaks that are no LUT tuple: curves to shared memory, warp-scope epilogue
#pragma unroll
            for (int pr = 0; pr < 3; pr++)
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int j = 8 * ((i >> 1) ? row_hi : row_lo) + 2 * t + (i & 1);
                    AT_CHECK(pr * CSTRIDE + G::NJ <= 3 * CSTRIDE && PAD + 8 * (pr * CSTRIDE + G::NJ) <= 2 * PLANE * 3);   // curve scratch stays inside this warp's planes
                    if (j < G::NJ)
                        curve_base[pr * CSTRIDE + j] = 65536LL * acc[pr][0][i] + 256LL * acc[pr][1][i] + (long long)acc[pr][2][i];
                }
            __syncwarp();
            epilogue_warp<L, PAD, G::NJ, CSTRIDE>(curve_base, best3[0], best3[1], best3[2], s.gauss, p, f, lane);
        }
        __syncwarp();   // planes and scratch are rewritten by the next frame
    }
}

template <int NBITS, int L, int WARPS, int CTAS_PER_SM, bool PRUNE, int UNROLL = 3>
static cudaError_t launch_imma(const AtFusedParams &p, int sm_count, cudaStream_t st)
{
    using S = ImmaSmem<NBITS, L, WARPS>;
    auto kern = at_fused_imma_kernel<NBITS, L, WARPS, CTAS_PER_SM, UNROLL, PRUNE>;
    const int smem = (int)sizeof(S);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    unsigned long long grid = (unsigned long long)sm_count * per_sm;
    const unsigned long long need = (p.n_frames + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    if (grid == 0) return cudaSuccess;
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(p);
    at_count_launch();
    return cudaGetLastError();
}

} // namespace atk

bool at_fused_imma_supports(const AtShape &sh)
{
    return sh.n_mics == 3 && ((sh.n_bits == 10 && (sh.max_shift == 46 || sh.max_shift == 44)) ||
                              (sh.n_bits == 12 && sh.max_shift == 46));
}

cudaError_t at_launch_fused_imma(const AtShape &sh, const AtFusedParams &p, int sm_count, cudaStream_t st)
{
    if (p.sig16 || sh.n_mics != 3) return cudaErrorInvalidValue;
    if (sh.n_bits == 10 && sh.max_shift == 46) {
        static const int prune_env = getenv("AT_IMMA_PRUNE") ? atoi(getenv("AT_IMMA_PRUNE")) : 1;   // 0: always all twelve products
        // the certified nine-product shortcut serves lags / cell / xy / gate; whole curves go through the plain kernel
        const bool prune = prune_env && p.peak_tab && !(p.raw || p.corr || p.classes || p.highest);
        // k-loop unrolled by 3: 11 and 33 measured 7-12 % slower
        return prune ? atk::launch_imma<10, 46, 4, 4, true>(p, sm_count, st) : atk::launch_imma<10, 46, 4, 4, false>(p, sm_count, st);
    }
    if (sh.n_bits == 10 && sh.max_shift == 44) {
        const bool prune = p.peak_tab && !(p.raw || p.corr || p.classes || p.highest);
        return prune ? atk::launch_imma<10, 44, 4, 4, true>(p, sm_count, st) : atk::launch_imma<10, 44, 4, 4, false>(p, sm_count, st);
    }
    if (sh.n_bits == 12 && sh.max_shift == 46) return atk::launch_imma<12, 46, 7, 1, false>(p, sm_count, st);   // 7 warps x 25.5 KB of planes fill the SM
    return cudaErrorInvalidValue;
}
