// at_fused_imma3.cu -- tensor-core localization kernel, low-instruction-count variant for the
// reference frame length (1024 samples).  Same mathematics and mapping as at_fused_imma.cu (one warp
// per frame, byte-split Toeplitz x Hankel IMMA tiles, bit-exact); what changes is how the fragments
// reach the registers.  On B200 a legacy mma.sync holds the issue port for its whole duration, so
// the loop is bound by the NUMBER of instructions issued around the 396 IMMA:
//
//   A (Hankel, y side): one ldmatrix.x4 per byte plane and k-step loads the whole 16x32 fragment
//     straight into its four registers.  ldmatrix wants 16-byte aligned rows and the Hankel rows
//     are 8 bytes apart, so every y plane is stored twice: E (as is) and O (advanced by 8 bytes);
//     even rows read E, odd rows read O, both at multiples of 16.
//   B (Toeplitz, x side): column n is the plane shifted by n bytes.  Every x plane is stored in
//     four copies delayed by 0..3 bytes; lane (g, t) reads copy g & 3, where its 4 bytes are word
//     aligned: two plain 32-bit loads per plane and k-step, no funnel shifts.
//
// 4 ldmatrix + 8 LDS per k-step instead of 28 LDS + 8 SHF (+ moves); paid for by ~170 extra prep
// instructions per frame (shifted copies) and 26 KB of shared memory per warp (8 warps per SM).
#include <limits.h>
#include <stdlib.h>

#include "at_imma_common.cuh"

namespace atk {

template <int L>
struct Imma3Geo {
    static constexpr int NBITS = 10, N = 1024;
    static constexpr int PAD = round_up(L, 16);                 // 48
    static constexpr int KSTEPS = ceil_div(N + 8, 32);          // 33
    static constexpr int PLANE = 1184;                          // bytes per plane; = 32 (mod 128): see banks below
    static constexpr int NJ = 96, NL = 2 * L + 1;
    // per-warp slice, byte offsets.  x planes xp = 0..3 (a.hi, a.lo, b.hi, b.lo), copy c = 0..3 delayed by c
    // bytes, laid out [xp][c] so the four copies a warp reads together sit 32 banks-bytes apart (conflict-free).
    static __host__ __device__ constexpr int XC(int xp, int c) { return (xp * 4 + c) * PLANE; }
    // y planes yp = 0..3 (b.hi, b.lo, c.hi, c.lo): E copy and O copy (O[q] = E[q + 8]).  b's E copies are the
    // undelayed x copies.  O bases are 64 (mod 128) away from their E so that the 8 rows of an ldmatrix 8x8
    // (4 from E, 4 from O, same offsets) touch 8 distinct 16-byte bank groups.
    static constexpr int YBASE = 16 * PLANE;                    // 18944 = 0 (mod 128)
    static constexpr int YSTRIDE = 1216;                        // = 64 (mod 128)
    static __host__ __device__ constexpr int YE(int yp) { return yp == 0 ? XC(2, 0) : yp == 1 ? XC(3, 0) : YBASE + (yp - 2) * 2 * YSTRIDE; }
    static __host__ __device__ constexpr int YO(int yp) { return yp < 2 ? YBASE + 4 * YSTRIDE + yp * 2 * YSTRIDE + 64 : YBASE + (yp - 2) * 2 * YSTRIDE + YSTRIDE; }
    static constexpr int SLICE = round_up(YO(1) + PLANE, 128);   // 27520 bytes per warp
    static_assert(PLANE % 128 == 32 && YBASE % 128 == 0 && YSTRIDE % 128 == 64, "bank layout");
    static_assert((YO(0) - YE(0)) % 128 == 64 && (YO(1) - YE(1)) % 128 == 64 && (YO(2) - YE(2)) % 128 == 64 &&
                  (YO(3) - YE(3)) % 128 == 64, "E/O copies must be 64 (mod 128) apart");
    static_assert(YO(1) + PLANE <= SLICE && YO(3) + PLANE <= YO(0), "planes overlap");
    static_assert(32 * (KSTEPS - 1) + 16 + 8 * 14 + 16 <= PLANE, "fragment reads stay inside a plane");
    static_assert(PAD + L < NJ && NJ * 8 <= N && PAD % 16 == 0, "lag range / curve scratch");
};

template <int L, int WARPS>
struct Imma3Smem {
    using G = Imma3Geo<L>;
    alignas(16) uint32_t win2[G::N];                 // pre-masked 2*W, chunk-interleaved (see at_fused_imma.cu)
    float gauss[2 * L + 1];
    alignas(128) uint8_t slice[WARPS][G::SLICE];
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&a)[4], uint32_t smem_addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(smem_addr));
}
__device__ __forceinline__ uint32_t lds32(uint32_t smem_addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_addr));
    return v;
}

template <int L, int WARPS, int CTAS_PER_SM>
__global__ void __launch_bounds__(WARPS * 32, CTAS_PER_SM) at_fused_imma3_kernel(const AtFusedParams p)
{
    using G = Imma3Geo<L>;
    using S = Imma3Smem<L, WARPS>;
    constexpr int N = G::N, PAD = G::PAD, PLANE = G::PLANE;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    S &s = *reinterpret_cast<S *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    // one-time CTA setup: zero every slice (pads stay zero), window, Gaussian factors
    for (int i = tid; i < (int)(sizeof(s.slice) / 16); i += WARPS * 32)
        reinterpret_cast<uint4 *>(&s.slice[0][0])[i] = make_uint4(0, 0, 0, 0);
    imma_win_fill(s.win2, p.window, N, tid, WARPS * 32);
    for (int i = tid; i < 2 * L + 1; i += WARPS * 32) s.gauss[i] = p.gauss[i];
    __syncthreads();

    uint8_t *const sl = &s.slice[warp][0];
    const uint32_t sl_s = smem_u32(sl);
    // A (ldmatrix): this lane supplies the address of row r of 8x8 block mi (see header comment)
    const int mi = lane >> 3, rr = lane & 7, arow = rr + 8 * (mi & 1), akofs = 16 * (mi >> 1);
    uint32_t a_addr[4];
#pragma unroll
    for (int yp = 0; yp < 4; yp++)
        a_addr[yp] = sl_s + ((arow & 1) ? G::YO(yp) : G::YE(yp)) + akofs + 8 * (arow & ~1);
    // B: copy g & 3 of each x plane, word-aligned at k = 4t (and 4t + 16)
    const uint32_t b_addr = sl_s + (g & 3) * PLANE + PAD + 4 * t - 4 * (g >> 2);
    const bool extras = p.gate || p.raw || p.corr || p.cell || p.highest || p.xy || p.classes;

    const unsigned long long stride = (unsigned long long)gridDim.x * WARPS;
    for (unsigned long long f = (unsigned long long)blockIdx.x * WARPS + warp; f < p.n_frames; f += stride) {
        const uint8_t *src = p.adc + f * (unsigned long long)(3 * N);
        const int head = p.heads ? (p.heads[f] & (N - 1)) : 0;

        // ---- channel sums -> floor mean (rolling_buffer.c:48-64); the sum is rotation invariant
        uint4 raw[6];
        int mean[3];
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            unsigned sum = 0;
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const uint4 v = ldg_stream(src + ch * N + q * 512 + lane * 16);
                raw[ch * 2 + q] = v;
                sum = __dp4a(v.x, 0x01010101u, sum); sum = __dp4a(v.y, 0x01010101u, sum);
                sum = __dp4a(v.z, 0x01010101u, sum); sum = __dp4a(v.w, 0x01010101u, sum);
            }
            sum = __reduce_add_sync(0xffffffffu, sum);
            mean[ch] = (int)(sum >> 10);
        }

        // ---- DC removal, <<8, window -> byte planes (rolling_buffer.c:66, buffer.c:16, :8-9), then the
        //      delayed x copies and the advanced y copies
        if ((head & 15) == 0) {
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                uint32_t carry_hi = 0, carry_lo = 0;          // last word of the previous chunk (lane 31), for lane 0
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int j0 = q * 512 + lane * 16;
                    const int i0 = (j0 - head) & (N - 1);
                    const uint4 v = raw[ch * 2 + q];
                    const uint32_t rw[4] = {v.x, v.y, v.z, v.w};
                    uint32_t hi[4], lo[4];
                    imma_prep16(rw, mean[ch], s.win2, i0, hi, lo);
                    uint8_t *const at = sl + PAD + i0;           // plane-relative position of this chunk
                    if (ch < 2) {       // x side (a, b): copies delayed by 0..3 bytes
                        // chronological predecessor word: lane-1's last word; for the first chunk of the frame: zero
                        uint32_t ph = __shfl_up_sync(0xffffffffu, hi[3], 1), plo = __shfl_up_sync(0xffffffffu, lo[3], 1);
                        if (head == 0) {
                            if (lane == 0) { ph = carry_hi; plo = carry_lo; }
                            carry_hi = __shfl_sync(0xffffffffu, hi[3], 31); carry_lo = __shfl_sync(0xffffffffu, lo[3], 31);
                        } else {        // rotated frame: the predecessor chunk lives in another (q, lane); fetch it via smem later
                            if (i0 == 0) { ph = 0; plo = 0; }
                        }
                        *reinterpret_cast<uint4 *>(at + G::XC(2 * ch, 0)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4 *>(at + G::XC(2 * ch + 1, 0)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
#pragma unroll
                        for (int c = 1; c < 4; c++) {
                            *reinterpret_cast<uint4 *>(at + G::XC(2 * ch, c)) =
                                make_uint4(__funnelshift_l(ph, hi[0], 8 * c), __funnelshift_l(hi[0], hi[1], 8 * c),
                                           __funnelshift_l(hi[1], hi[2], 8 * c), __funnelshift_l(hi[2], hi[3], 8 * c));
                            *reinterpret_cast<uint4 *>(at + G::XC(2 * ch + 1, c)) =
                                make_uint4(__funnelshift_l(plo, lo[0], 8 * c), __funnelshift_l(lo[0], lo[1], 8 * c),
                                           __funnelshift_l(lo[1], lo[2], 8 * c), __funnelshift_l(lo[2], lo[3], 8 * c));
                            if (i0 == N - 16) {   // the delayed copies run c bytes past the frame end
                                *reinterpret_cast<uint32_t *>(at + 16 + G::XC(2 * ch, c)) = __funnelshift_l(hi[3], 0u, 8 * c);
                                *reinterpret_cast<uint32_t *>(at + 16 + G::XC(2 * ch + 1, c)) = __funnelshift_l(lo[3], 0u, 8 * c);
                            }
                        }
                    }
                    if (ch >= 1) {      // y side (b, c): E copy (b's is its undelayed x copy, written above) and O copy
                        const int yp = 2 * (ch - 1);
                        if (ch == 2) {
                            *reinterpret_cast<uint4 *>(at + G::YE(yp)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                            *reinterpret_cast<uint4 *>(at + G::YE(yp + 1)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        }
                        *reinterpret_cast<uint2 *>(at - 8 + G::YO(yp)) = make_uint2(hi[0], hi[1]);
                        *reinterpret_cast<uint2 *>(at + G::YO(yp)) = make_uint2(hi[2], hi[3]);
                        *reinterpret_cast<uint2 *>(at - 8 + G::YO(yp + 1)) = make_uint2(lo[0], lo[1]);
                        *reinterpret_cast<uint2 *>(at + G::YO(yp + 1)) = make_uint2(lo[2], lo[3]);
                    }
                }
            }
            if (head != 0) {
                // rotated (16-aligned) frame: the word before each chunk came from a different lane; patch the first
                // word of every delayed-copy chunk from the undelayed copy now that all of it is in shared memory
                __syncwarp();
#pragma unroll
                for (int xp = 0; xp < 4; xp++)
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const int i0 = q * 512 + lane * 16;
                        const uint8_t *base = sl + G::XC(xp, 0) + PAD + i0;
                        const uint32_t prev = *reinterpret_cast<const uint32_t *>(base - 4), cur = *reinterpret_cast<const uint32_t *>(base);
#pragma unroll
                        for (int c = 1; c < 4; c++)
                            *reinterpret_cast<uint32_t *>(sl + G::XC(xp, c) + PAD + i0) = __funnelshift_l(prev, cur, 8 * c);
                    }
            }
        } else {
            // ring head not 16-aligned: byte-wise stores into every copy (rare; capture heads are arbitrary).
            // The first c bytes of a delayed copy's data region belong to the zero pad; the previous frame's
            // epilogue may have used that region as scratch, so clear them first.
            if (lane < 12) *reinterpret_cast<uint32_t *>(sl + G::XC(lane / 3, 1 + lane % 3) + PAD) = 0u;
            __syncwarp();
            for (int ch = 0; ch < 3; ch++)
                for (int q = 0; q < 2; q++) {
                    const uint4 v = raw[ch * 2 + q];
                    const uint32_t rw[4] = {v.x, v.y, v.z, v.w};
                    for (int e = 0; e < 16; e++) {
                        const int i = (q * 512 + lane * 16 + e - head) & (N - 1);
                        const int pr = imma_prep1(rw[e >> 2] >> (8 * (e & 3)), mean[ch], s.win2, i);
                        const uint8_t bh = (uint8_t)(pr >> 16), bl = (uint8_t)(pr >> 8);
                        if (ch < 2)
                            for (int c = 0; c < 4; c++) { sl[G::XC(2 * ch, c) + PAD + i + c] = bh; sl[G::XC(2 * ch + 1, c) + PAD + i + c] = bl; }
                        if (ch == 2) { sl[G::YE(2) + PAD + i] = bh; sl[G::YE(3) + PAD + i] = bl; }
                        if (ch >= 1) { sl[G::YO(2 * (ch - 1)) + PAD + i - 8] = bh; sl[G::YO(2 * (ch - 1) + 1) + PAD + i - 8] = bl; }
                    }
                }
        }
        if (p.power) {   // rolling_buffer.c:68-70: power of the DC-removed samples (9-bit differences)
            for (int ch = 0; ch < 3; ch++) {
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
                const uint8_t *ph = sl + (ch < 2 ? G::XC(2 * ch, 0) : G::YE(2)) + PAD + i;
                const uint8_t *plo = sl + (ch < 2 ? G::XC(2 * ch + 1, 0) : G::YE(3)) + PAD + i;
                p.windowed[f * (unsigned long long)(3 * N) + idx] = (int16_t)(((int)(signed char)*ph << 8) | *plo);
            }

        // next frame of this warp -> L1/L2 while the tensor pipe works
        if (f + stride < p.n_frames) {
            const uint8_t *nx = p.adc + (f + stride) * (unsigned long long)(3 * N) + lane * 128;
            if (lane * 128 < 3 * N) asm volatile("prefetch.global.L1 [%0];" ::"l"(nx));
        }

        // ---- 33 k-steps x (4 ldmatrix + 8 LDS + 12 IMMA): pairs (a,b), (a,c), (b,c); x = first, y = second mic
        int acc[3][3][4];
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[a][b][c] = 0;
        const int nsteps = (p.debug_skip & 2) ? 0 : G::KSTEPS;
#pragma unroll 3
        for (int ks = 0; ks < nsteps; ks++) {
            const int k0 = 32 * ks;
            uint32_t Y[4][4], X[4][2];
#pragma unroll
            for (int yp = 0; yp < 4; yp++) ldmatrix_x4(Y[yp], a_addr[yp] + k0);
#pragma unroll
            for (int xp = 0; xp < 4; xp++) {
                X[xp][0] = lds32(b_addr + G::XC(xp, 0) + k0);
                X[xp][1] = lds32(b_addr + G::XC(xp, 0) + k0 + 16);
            }
            // Y[0] = b.hi, Y[1] = b.lo, Y[2] = c.hi, Y[3] = c.lo;  X[0] = a.hi, X[1] = a.lo, X[2] = b.hi, X[3] = b.lo
            mma_s8_s8(acc[0][0], Y[0], X[0]); mma_s8_u8(acc[0][1], Y[0], X[1]); mma_u8_u8(acc[0][2], Y[1], X[1]);
            mma_s8_s8(acc[1][0], Y[2], X[0]); mma_s8_u8(acc[1][1], Y[2], X[1]); mma_u8_u8(acc[1][2], Y[3], X[1]);
            mma_s8_s8(acc[2][0], Y[2], X[2]); mma_s8_u8(acc[2][1], Y[2], X[3]); mma_u8_u8(acc[2][2], Y[3], X[3]);
            mma_u8_s8(acc[0][1], Y[1], X[0]); mma_u8_s8(acc[1][1], Y[3], X[0]); mma_u8_s8(acc[2][1], Y[3], X[2]);
        }

        __syncwarp();   // every lane is done reading the planes: the data regions of three of them now hold the curves
        constexpr int CSTRIDE = PLANE / 8;                       // int64 stride between consecutive planes
        long long *const curve_base = reinterpret_cast<long long *>(sl + PAD);   // planes XC(0,0), XC(0,1), XC(0,2)
        // ---- recombine in int64, arg-max per pair (correlations.c:20-23): key = value * 128 + (127 - j)
        int best3[3];
#pragma unroll
        for (int pr = 0; pr < 3; pr++) {
            long long key = LLONG_MIN;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int j = 8 * (g + 8 * (i >> 1)) + 2 * t + (i & 1);
                const long long v = 65536LL * acc[pr][0][i] + 256LL * acc[pr][1][i] + (long long)acc[pr][2][i];
                if (extras && j < G::NJ) curve_base[pr * CSTRIDE + j] = v;
                const long long k = v * 128 + (127 - j);
                if (j >= PAD - L && j <= PAD + L && k > key) key = k;
            }
            key = warp_max_i64(key);
            best3[pr] = 127 - (int)(key & 127) - PAD;
        }
        if (lane < 3 && p.lags) p.lags[f * 3 + lane] = lane == 0 ? best3[0] : (lane == 1 ? best3[1] : best3[2]);
        if (extras && !(p.debug_skip & 4)) {
            __syncwarp();
            epilogue_warp<L, PAD, G::NJ, CSTRIDE>(curve_base, best3[0], best3[1], best3[2], s.gauss, p, f, lane);
        }
        __syncwarp();   // planes and scratch are rewritten by the next frame
    }
}

template <int L, int WARPS, int CTAS_PER_SM>
static cudaError_t launch_imma3(const AtFusedParams &p, int sm_count, cudaStream_t st)
{
    using S = Imma3Smem<L, WARPS>;
    auto kern = at_fused_imma3_kernel<L, WARPS, CTAS_PER_SM>;
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

bool at_fused_imma3_supports(const AtShape &sh)
{
    return sh.n_mics == 3 && sh.n_bits == 10 && (sh.max_shift == 46 || sh.max_shift == 44);
}

cudaError_t at_launch_fused_imma3(const AtShape &sh, const AtFusedParams &p, int sm_count, cudaStream_t st)
{
    if (p.sig16 || !at_fused_imma3_supports(sh)) return cudaErrorInvalidValue;
    if (sh.max_shift == 46) return atk::launch_imma3<46, 4, 2>(p, sm_count, st);
    return atk::launch_imma3<44, 4, 2>(p, sm_count, st);
}
