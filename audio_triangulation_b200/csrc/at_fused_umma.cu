// at_fused_umma.cu -- fused localization kernel on the 5th-generation tensor cores (tcgen05 / UMMA), AT_KERNEL_UMMA.
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
// One tcgen05.mma kind::i8 (K = 32) covers 512 samples, so a frame needs 2 K-steps instead of 33 mma.sync steps.
// (tools/probes/umma_probe.cu is the stand-alone proof of this operand trick.)
//
// int16 samples are split into balanced signed digits w = 256 h + l, h and l both in [-128, 127] (possible because
// the windowed samples stay inside [-32767, 32511]); corr = 65536 (h.h) + 256 (h.l + l.h) + (l.l) as in the
// mma.sync kernel, each of the four digit products in its own 16 TMEM columns, int32 (|sum| < 2^26), recombined in
// int64: bit-exact.  An MMA of this shape costs about (A bytes + B bytes) / 128 cycles whatever N is (measured,
// tools/probes/umma_probe.cu: 46 cycles at N = 16..64), so the x planes are batched into one wide B operand.
//
// The diagonal sums are the CUDA cores' job.  To halve them every plane is stored twice, the second copy advanced by
// 8 bytes, and both copies are accumulated into the same tile:  D2[m][phi] = D[m][phi] + D[m+8][phi+8]  for
// phi = 0..7, i.e. two entries of the same diagonal; corr[s] = sum_{phi<8} D2[s + PAD + phi][phi].
//
// CTA = 16 warps, one CTA per SM, persistent, warp-specialised:
//   warp 0      one lane issues, frame after frame, the 16 MMAs of a frame (2 K-steps x 2 copies x 4 y planes; the
//               B operand is the four x planes a.h a.l b.h b.l side by side, N = 64) and commits them to two mbarriers
//               (accumulators ready / planes free);
//   warps 1-7   prep: one frame each -- coalesced 16-byte loads, DC removal, <<8, window (as imma_prep16), digit
//               planes to shared memory (both copies), mbarrier arrive;
//   warps 8-15  two epilogue sets of four warps, set k bound to TMEM slot k (192 columns = 12 tiles): TMEM ->
//               registers (warp w reads lane quadrant w % 4), transposing scatter through shared memory, diagonal
//               sums, int64 recombination, first-max arg-max (correlations.c:20-23), then peak-tuple look-up or the
//               warp-scope epilogue of at_imma_common.cuh for everything else.
//
// Status (measured on the B200): bit-exact on every parity test, 152 M frames/s -- 0.71x the mma.sync
// kernel, which therefore stays the default.  Two things bound it: an MMA of this shape costs ~63-73 cycles whatever
// N <= 64 is (4 KB of A operand per instruction, fetched at ~64-128 B/clk: 16 MMAs = ~1 170 cycles per frame), and the
// 9 x 8 x 128 diagonal terms per frame have to be added by the CUDA cores.
#include <limits.h>
#include <stdlib.h>

#include "at_imma_common.cuh"
#include "at_umma_common.cuh"

namespace atk {

template <int L>
struct UmmaGeo {
    static constexpr int N = 1024, NBITS = 10;
    static constexpr int PAD = 48;                      // lag index j = s + PAD; also the left zero pad of a plane
    static constexpr int PLANE = 1152;                  // bytes per plane buffer: 48 zeros, 1024 samples, 80 zeros
    static constexpr int NPLANES = 12;                  // [copy][channel a,b,c][h,l]; copy 1 = copy 0 advanced by 8 bytes
    static constexpr int FRAME = NPLANES * PLANE;       // 13 824 bytes of planes per frame
    static constexpr int NJ = 96;                       // lag slots kept (j = 0..95), j in [PAD-L, PAD+L] are real
    static constexpr int NL = 2 * L + 1;
    static constexpr int ZP = 136;                      // words per (pair, class, phase) column of the transposing scratch
    static constexpr int TCOLS = 192;                   // TMEM columns per frame: 12 tiles (3 pairs x {hh, hl, lh, ll}) x 16
    static constexpr int PREP_WARPS = 7, SETS = 2;      // warp 0 issues the MMAs, warps 1-7 prepare, warps 8-15 = 2 epilogue sets
    static_assert(PAD >= L && PAD + L + 15 < 128 && PAD + L < NJ, "lag window must fit the 128-row tile");
    static_assert(127 + 16 * 63 + 8 < PLANE, "A operand reads stay inside a plane buffer");
};

template <int L>
struct UmmaSmem {
    using G = UmmaGeo<L>;
    alignas(128) uint8_t planes[G::PREP_WARPS][G::FRAME];
    alignas(16) int z[G::SETS][3][3][8][G::ZP];         // [set][pair][class][phase][row - phase + 7]
    alignas(16) long long curve[G::SETS][3][G::NJ];     // raw curves by lag index (input of epilogue_warp)
    alignas(16) long long part[G::SETS][3][4];          // per-block arg-max keys
    alignas(16) uint32_t win2[G::N];
    float gauss[2 * L + 1];
    alignas(8) uint64_t full[G::SETS], empty[G::SETS], ready[G::PREP_WARPS], sfree[G::PREP_WARPS];
    uint32_t tmem_base;
};

template <int L>
__global__ void __launch_bounds__(512, 1) at_fused_umma_kernel(const AtFusedParams p)
{
    using G = UmmaGeo<L>;
    using S = UmmaSmem<L>;
    constexpr int N = G::N, PAD = G::PAD, PLANE = G::PLANE, P = G::PREP_WARPS;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    S &s = *reinterpret_cast<S *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- one-time CTA set-up: zero the planes (pads stay zero), window, Gaussian factors, barriers, TMEM
    for (int i = tid; i < (int)(sizeof(s.planes) / 16); i += 512)
        reinterpret_cast<uint4 *>(&s.planes[0][0])[i] = make_uint4(0, 0, 0, 0);
    imma_win_fill(s.win2, p.window, N, tid, 512);
    for (int i = tid; i < 2 * L + 1; i += 512) s.gauss[i] = p.gauss[i];
    if (tid == 0) {
        for (int k = 0; k < G::SETS; k++) { mbar_init(&s.full[k], 1); mbar_init(&s.empty[k], 4); }
        for (int w = 0; w < P; w++) { mbar_init(&s.ready[w], 1); mbar_init(&s.sfree[w], 1); }
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
    // frames of this CTA: f = blockIdx.x + gridDim.x * i, i = 0, 1, 2, ...; frame i is prepared by prep warp i % P,
    // accumulated in TMEM slot i % SETS and finished by epilogue set i % SETS.  Every mbarrier is waited on in phase
    // order by exactly one party (a parity wait only distinguishes the current from the preceding phase).

    if (warp == 0) {
        // =================================================================== MMA issue (one elected lane, frames in order)
        constexpr uint32_t I64 = umma_idesc(64), I32 = umma_idesc(32);
        constexpr uint32_t LBO = (128u >> 4) << 16;                     // K groups of 8 rows are 128 bytes apart
        constexpr uint32_t HI_A = (16u >> 4) | 0x4000u;                 // A: MN chunks 16 bytes apart (Hankel), version 1
        constexpr uint32_t HI_B = ((uint32_t)PLANE >> 4) | 0x4000u;     // B: one 16-phase chunk per x plane
        for (unsigned long long i = 0;; i++) {
            if (blockIdx.x + gstride * i >= nf) break;
            const unsigned w = (unsigned)(i % P), v = (unsigned)(i / P), slot = (unsigned)(i % G::SETS), u = (unsigned)(i / G::SETS);
            mbar_wait(&s.ready[w], v & 1);                          // planes of frame i are in shared memory
            if (u >= 1) mbar_wait(&s.empty[slot], (u - 1) & 1);     // the epilogue set has drained the slot
            tc_fence_after();
            if (elect_one()) {
                const uint32_t b16 = (smem_u32(&s.planes[w][0]) >> 4) + LBO;   // start-address field of plane a.h, copy 0
                const uint32_t cb = tmem + slot * G::TCOLS;
                if (!(p.debug_skip & 2))
#pragma unroll
                for (int kk = 0; kk < 2; kk++)
#pragma unroll
                    for (int copy = 0; copy < 2; copy++) {
                        const uint32_t acc = (kk | copy) ? 1u : 0u;
                        const uint32_t pl = b16 + (uint32_t)((copy * 6 * PLANE + 512 * kk) >> 4);   // plane a.h of this copy, K step kk
                        const uint32_t xb = pl + (PAD >> 4);            // B: x planes a.h a.l b.h b.l side by side
                        // A: one y plane, 128 Hankel rows.  Tiles: [x.h | x.l] per x channel.
                        umma_i8_lohi(cb + 0, pl + 4 * (PLANE >> 4), HI_A, xb, HI_B, I64, acc);     // c.h: ac hh, ac hl, bc hh, bc hl
                        umma_i8_lohi(cb + 64, pl + 5 * (PLANE >> 4), HI_A, xb, HI_B, I64, acc);    // c.l: ac lh, ac ll, bc lh, bc ll
                        umma_i8_lohi(cb + 128, pl + 2 * (PLANE >> 4), HI_A, xb, HI_B, I32, acc);   // b.h: ab hh, ab hl
                        umma_i8_lohi(cb + 160, pl + 3 * (PLANE >> 4), HI_A, xb, HI_B, I32, acc);   // b.l: ab lh, ab ll
                    }
                umma_commit(&s.full[slot]);
                umma_commit(&s.sfree[w]);
            }
            __syncwarp();
        }
    } else if (warp <= P) {
        // =================================================================== prep warps
        const int w = warp - 1;
        uint8_t *const buf = &s.planes[w][0];
        auto plane = [&](int copy, int ch, int hl) -> uint8_t * { return buf + ((copy * 3 + ch) * 2 + hl) * PLANE; };
        for (unsigned long long i = w;; i += P) {
            const unsigned long long f = blockIdx.x + gstride * i;
            if (f >= nf) break;
            const unsigned v = (unsigned)(i / P);
            const uint8_t *src = p.adc + f * (unsigned long long)(3 * N);
            const int head = p.heads ? (p.heads[f] & (N - 1)) : 0;

            // channel sums -> floor mean (rolling_buffer.c:48-64); the sum is rotation invariant
            uint4 raw[6];
            int mean[3];
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                unsigned sum = 0;
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const uint4 x = ldg_stream(src + ch * N + q * 512 + lane * 16);
                    raw[ch * 2 + q] = x;
                    sum = __dp4a(x.x, 0x01010101u, sum); sum = __dp4a(x.y, 0x01010101u, sum);
                    sum = __dp4a(x.z, 0x01010101u, sum); sum = __dp4a(x.w, 0x01010101u, sum);
                }
                sum = __reduce_add_sync(0xffffffffu, sum);
                mean[ch] = (int)(sum >> 10);
            }
            {   // next frame of this warp -> L1/L2
                const unsigned long long fn = f + gstride * P;
                if (fn < nf && lane * 128 < 3 * N) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.adc + fn * (unsigned long long)(3 * N) + lane * 128));
            }
            // the tensor core must be done with this warp's previous frame
            if (v >= 1) mbar_wait(&s.sfree[w], (v - 1) & 1);
            if (!(p.debug_skip & 1))
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int j0 = q * 512 + lane * 16;
                    const uint4 x = raw[ch * 2 + q];
                    const uint32_t rw[4] = {x.x, x.y, x.z, x.w};
                    if ((head & 15) == 0) {
                        const int i0 = (j0 - head) & (N - 1);
                        uint32_t hi[4], lo[4];
                        umma_prep16(rw, mean[ch], s.win2, i0, hi, lo);
                        *reinterpret_cast<uint4 *>(plane(0, ch, 0) + PAD + i0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4 *>(plane(0, ch, 1) + PAD + i0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        // second copy, advanced by 8 bytes: sample i sits at PAD - 8 + i
                        *reinterpret_cast<uint2 *>(plane(1, ch, 0) + PAD - 8 + i0) = make_uint2(hi[0], hi[1]);
                        *reinterpret_cast<uint2 *>(plane(1, ch, 0) + PAD + i0) = make_uint2(hi[2], hi[3]);
                        *reinterpret_cast<uint2 *>(plane(1, ch, 1) + PAD - 8 + i0) = make_uint2(lo[0], lo[1]);
                        *reinterpret_cast<uint2 *>(plane(1, ch, 1) + PAD + i0) = make_uint2(lo[2], lo[3]);
                    } else {   // ring head not 16-aligned: scalar stores (rare; capture heads are arbitrary)
#pragma unroll
                        for (int e = 0; e < 16; e++) {
                            const int ii = (j0 + e - head) & (N - 1);
                            const int q24 = imma_prep1(rw[e >> 2] >> (8 * (e & 3)), mean[ch], s.win2, ii) + 0x8000;
                            const uint8_t hb = (uint8_t)(q24 >> 16), lb = (uint8_t)((q24 >> 8) ^ 0x80);
                            plane(0, ch, 0)[PAD + ii] = hb; plane(0, ch, 1)[PAD + ii] = lb;
                            plane(1, ch, 0)[PAD - 8 + ii] = hb; plane(1, ch, 1)[PAD - 8 + ii] = lb;
                        }
                    }
                }
                if (p.power) {   // rolling_buffer.c:68-70
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
                    const int ch = idx / N, ii = idx % N;
                    p.windowed[f * (unsigned long long)(3 * N) + idx] =
                        (int16_t)((int)(signed char)plane(0, ch, 0)[PAD + ii] * 256 + (int)(signed char)plane(0, ch, 1)[PAD + ii]);
                }
            fence_proxy_async();          // this lane's plane bytes -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.ready[w]);
        }
    } else {
        // =================================================================== epilogue sets
        const int set = (warp - 1 - P) >> 2, wq = warp & 3, tset = wq * 32 + lane;   // tset = TMEM lane = tile row m
        int *const z = &s.z[set][0][0][0][0];
        long long *const curve = &s.curve[set][0][0];
        // TMEM column of the hh tile of each pair (x, y) = (a,b), (a,c), (b,c); hl follows at +16
        // and (lh, ll) sit OFF_L columns further (the tiles of the y.l plane)
        for (unsigned long long i = set;; i += G::SETS) {
            const unsigned long long f = blockIdx.x + gstride * i;
            if (f >= nf) break;
            const unsigned u = (unsigned)(i / G::SETS);
            mbar_wait(&s.full[set], u & 1);
            tc_fence_after();
            if (p.debug_skip & 4) {       // timing experiments: drain the slot without looking at it
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.empty[set]);
                continue;
            }
            const uint32_t ta = tmem + ((uint32_t)(wq * 32) << 16) + set * G::TCOLS;
#pragma unroll
            for (int pr = 0; pr < 3; pr++) {
                const uint32_t c_hh = pr == 0 ? 128 : (pr == 1 ? 0 : 32), off_l = pr == 0 ? 32 : 64;
                uint32_t hh[8], hl[8], lh[8], ll[8];
                tmem_ld8(ta + c_hh, hh); tmem_ld8(ta + c_hh + 16, hl);
                tmem_ld8(ta + c_hh + off_l, lh); tmem_ld8(ta + c_hh + off_l + 16, ll);
                tmem_ld_wait();
                // transposing scatter: entry (m, phi) belongs to lag index j = m - phi
                int *const zp = z + pr * 3 * 8 * G::ZP + tset + 7;
#pragma unroll
                for (int ph = 0; ph < 8; ph++) {
                    zp[(0 * 8 + ph) * G::ZP - ph] = (int)hh[ph];
                    zp[(1 * 8 + ph) * G::ZP - ph] = (int)hl[ph] + (int)lh[ph];
                    zp[(2 * 8 + ph) * G::ZP - ph] = (int)ll[ph];
                }
            }
            tc_fence_before();            // accumulators are out of TMEM: hand the slot back to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.empty[set]);
            named_bar(1 + set, 128);
            // diagonal sums: 9 blocks of 32 lags (3 pairs x 3), block b -> warp b % 4
            for (int b = wq; b < 9; b += 4) {
                const int pr = b / 3, j = (b % 3) * 32 + lane;
                const int *zr = z + pr * 3 * 8 * G::ZP + j + 7;
                int hh = 0, mid = 0, ll = 0;
#pragma unroll
                for (int ph = 0; ph < 8; ph++) {
                    hh += zr[(0 * 8 + ph) * G::ZP];
                    mid += zr[(1 * 8 + ph) * G::ZP];
                    ll += zr[(2 * 8 + ph) * G::ZP];
                }
                const long long val = 65536LL * hh + 256LL * mid + (long long)ll;
                curve[pr * G::NJ + j] = val;
                long long key = LLONG_MIN;
                if (j >= PAD - L && j <= PAD + L) key = val * 128 + (127 - j);   // largest value, then lowest lag
                key = warp_max_i64(key);
                if (lane == 0) s.part[set][pr][b % 3] = key;
            }
            named_bar(1 + set, 128);
            if (wq == 0) {
                int best3[3];
                long long peak[3];
#pragma unroll
                for (int pr = 0; pr < 3; pr++) {
                    long long key = s.part[set][pr][0];
                    if (s.part[set][pr][1] > key) key = s.part[set][pr][1];
                    if (s.part[set][pr][2] > key) key = s.part[set][pr][2];
                    best3[pr] = 127 - (int)(key & 127) - PAD;
                    peak[pr] = key >> 7;
                }
                if (lane < 3 && p.lags) p.lags[f * 3 + lane] = lane == 0 ? best3[0] : (lane == 1 ? best3[1] : best3[2]);
                const bool extras = p.gate || p.raw || p.corr || p.cell || p.highest || p.xy || p.classes;
                bool settled = !extras;
                if (extras && !(p.raw || p.corr || p.classes))
                    settled = peak_tuple_lookup<L>(p, f, lane, best3[0], best3[1], best3[2], peak);
                if (!settled) epilogue_warp<L, PAD, G::NJ, G::NJ>(curve, best3[0], best3[1], best3[2], s.gauss, p, f, lane);
            }
            named_bar(1 + set, 128);      // curve / part / z are rewritten by the next frame of this set
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

} // namespace atk

bool at_fused_umma_supports(const AtShape &sh)
{
    return sh.n_mics == 3 && sh.n_bits == 10 && (sh.max_shift == 46 || sh.max_shift == 44);
}

template <int L>
static cudaError_t launch_umma(const AtFusedParams &p, int sm_count, cudaStream_t st)
{
    auto kern = atk::at_fused_umma_kernel<L>;
    const int smem = (int)sizeof(atk::UmmaSmem<L>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    unsigned long long grid = (unsigned long long)sm_count;
    if (grid > p.n_frames) grid = p.n_frames;
    if (grid == 0) return cudaSuccess;
    kern<<<(unsigned)grid, 512, smem, st>>>(p);
    at_count_launch();
    return cudaGetLastError();
}

cudaError_t at_launch_fused_umma(const AtShape &sh, const AtFusedParams &p, int sm_count, cudaStream_t st)
{
    if (p.sig16 || !at_fused_umma_supports(sh)) return cudaErrorInvalidValue;
    return sh.max_shift == 46 ? launch_umma<46>(p, sm_count, st) : launch_umma<44>(p, sm_count, st);
}
