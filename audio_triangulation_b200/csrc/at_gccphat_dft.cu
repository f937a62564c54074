// at_gccphat_dft.cu -- the inverse side of the GCC-PHAT variant as ONE dense tensor-core contraction.
// Only the 2L+1 lags of the admissible window are wanted from each pair's inverse transform, so instead of 14 inverse
// FFTs per frame (at_gccphat.cu, gcc_pair_kernel) the wanted outputs are evaluated directly,
//     y_p[s] = G_p[0] + (-1)^s G_p[N] + sum_{m=1}^{N-1} 2 (Re G_p[m] cos(pi m s / N) - Im G_p[m] sin(pi m s / N)),
// which is a GEMM with a CONSTANT left operand:  Y[128 x cols] = A[128 x 2N] . B[2N x cols],  A = the cos / -sin table
// (rows = lags, fp16), B = the whitened cross-spectra of (frame, pair) columns, 256 columns per CTA (a group of frames).
// tcgen05.mma kind::f16 (fp16 operands, fp32 accumulation in TMEM), M128 x N256 x K16, 4 per chunk of 32 bins.
// CTA = 18 warps: warp 0 streams A tiles (pre-tiled in the UMMA K-major canonical layout) and the group's whitened
// spectra by bulk copies; warps 1-16 (two threads per column) form G = conj(U_a) U_b for their (frame, pair) and write the
// B tile; warp 17 issues the MMAs; at the end warps 1-16 read the accumulators (lane = lag) and take the first-max arg-max.
// Not a reference algorithm (see at_gccphat.cu); checked against the FFT form and a float64 restatement.
#include <cuda_fp16.h>
#include <math.h>

#include "at_umma_common.cuh"

namespace atk {

struct GccDftGeo {
    static constexpr int COLS = 256;                 // (frame, pair) columns of a CTA = TMEM columns
    static constexpr int CB = 32;                    // bins per chunk: K = 64 = 4 MMAs of K16
    static constexpr int A_TILE = 128 * 2 * CB * 2;  // bytes: 128 lags x 64 k x fp16
    static constexpr int B_TILE = COLS * 2 * CB * 2; // bytes
    static constexpr int UROW = 144;                 // bytes per (frame, mic) row of a spectra tile: 32 half2 + 16 bytes of padding
    static constexpr int CONV_WARPS = 16;              // two threads per column, half of a chunk's bins each
    static constexpr int MMA_WARP = 1 + CONV_WARPS;
    static constexpr int THREADS = 32 * (2 + CONV_WARPS);
};

struct GccDftSmem {
    static constexpr int MAX_IN = 8;
    alignas(128) uint8_t b[2][GccDftGeo::B_TILE];
    uint2 part[4][GccDftGeo::COLS];                  // per lane quarter and column: (ordered key of the maximum, its row)
    alignas(8) uint64_t in_full[MAX_IN], in_empty[MAX_IN], b_full[2], b_empty[2], acc_full;
    uint32_t tmem_base;
    // followed by n_in input stages of (A tile, spectra tile): the bulk copies run several chunks ahead of the tensor
    // core (their latency is a few chunk times); the spectra tile's size depends on the frames per group
};

__host__ __device__ constexpr uint32_t umma_idesc_f16(int n)     // F32 accumulators, F16 A and B, both K-major, M = 128
{
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ uint32_t fkey(float v)      // order-preserving map float -> uint32, > 0 for every finite v
{
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k); }

// spec: [group][chunk][frame in group][mic][36] half2 (32 bins + padding; one contiguous tile per chunk); nyq: [frame][mic] half2 (bin N)
template <int NBITS>
__global__ void __launch_bounds__(GccDftGeo::THREADS, 1) gcc_dft_kernel(const __half *__restrict__ a_tiles, const __half2 *__restrict__ spec,
                                                                         const __half2 *__restrict__ nyq, int n_mics, int L, int ps_log2, int n_in,
                                                                         unsigned n_frames, int32_t *lags, float *peak)
{
    using G = GccDftGeo;
    constexpr int N = 1 << NBITS, N2 = 2 * N, NCH = N / G::CB;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    GccDftSmem &s = *reinterpret_cast<GccDftSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = n_mics * (n_mics - 1) / 2, PS = 1 << ps_log2, FG = G::COLS >> ps_log2;     // pair slots per frame, frames per group
    const uint32_t u_bytes = (uint32_t)(FG * n_mics * G::UROW);
    uint8_t *const in_tiles = smem_raw + ((sizeof(GccDftSmem) + 127) & ~(size_t)127);
    const uint32_t in_bytes = G::A_TILE + u_bytes;                  // one input stage: A tile, then the spectra tile
    const unsigned g = blockIdx.x;

    if (tid == 0) {
        for (int k = 0; k < n_in; k++) { mbar_init(&s.in_full[k], 1); mbar_init(&s.in_empty[k], 1 + G::CONV_WARPS); }   // the MMA commit + the converter warps
        for (int k = 0; k < 2; k++) { mbar_init(&s.b_full[k], G::CONV_WARPS); mbar_init(&s.b_empty[k], 1); }
        mbar_init(&s.acc_full, 1);
        fence_barrier_init();
    }
    if (warp == G::MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s.tmem_base)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;

    if (warp == 0) {
        // =================================================================== loader: A tile + the group's spectra tile per chunk
        if (lane == 0) {
            const uint8_t *ug = reinterpret_cast<const uint8_t *>(spec) + (size_t)g * NCH * u_bytes;
            for (int c = 0, si = 0, use = 0; c < NCH; c++) {
                if (use >= 1) mbar_wait(&s.in_empty[si], (use - 1) & 1);
                mbar_expect_tx(&s.in_full[si], in_bytes);
                bulk_g2s(in_tiles + si * in_bytes, reinterpret_cast<const uint8_t *>(a_tiles) + (size_t)c * G::A_TILE, G::A_TILE, &s.in_full[si]);
                bulk_g2s(in_tiles + si * in_bytes + G::A_TILE, ug + (size_t)c * u_bytes, u_bytes, &s.in_full[si]);
                if (++si == n_in) { si = 0; use++; }
            }
        }
    } else if (warp == G::MMA_WARP) {
        // =================================================================== MMA issue
        constexpr uint32_t IDESC = umma_idesc_f16(G::COLS);
        constexpr uint32_t LBO_A = 16 * 128, LBO_B = (G::COLS / 8) * 128, SBO = 128;    // K-major, no swizzle: 8 x 16 B core matrices
        for (int c = 0, si = 0, use = 0; c < NCH; c++) {
            const int st = c & 1;
            mbar_wait(&s.in_full[si], use & 1);
            mbar_wait(&s.b_full[st], (c >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a0 = smem_u32(in_tiles + si * in_bytes), b0 = smem_u32(&s.b[st][0]);
#pragma unroll
                for (int j = 0; j < 4; j++)
                    umma_f16(tmem, umma_desc(a0 + j * 2 * LBO_A, LBO_A, SBO), umma_desc(b0 + j * 2 * LBO_B, LBO_B, SBO), IDESC, (c | j) ? 1u : 0u);
                umma_commit(&s.b_empty[st]);
                umma_commit(&s.in_empty[si]);
                if (c == NCH - 1) umma_commit(&s.acc_full);
            }
            __syncwarp();
            if (++si == n_in) { si = 0; use++; }
        }
    } else {
        // =================================================================== one thread per column: G = conj(U_a) U_b -> B tile; arg-max
        const int n = (tid - 32) & (G::COLS - 1), qh = (tid - 32) >> 8, fi = n >> ps_log2, p = n & (PS - 1);      // column, half of the chunk
        const unsigned f = g * (unsigned)FG + (unsigned)fi;
        const bool valid = p < P && f < n_frames;
        int ma = 0, mb = 1;
        if (valid) {
            int rem = p;
            while (rem >= n_mics - 1 - ma) { rem -= n_mics - 1 - ma; ma++; }
            mb = ma + 1 + rem;
        }
        for (int c = 0, si = 0, use = 0; c < NCH; c++) {
            const int st = c & 1;
            mbar_wait(&s.in_full[si], use & 1);
            if (c >= 2) mbar_wait(&s.b_empty[st], ((c >> 1) - 1) & 1);
            if (valid) {
                const uint8_t *ut = in_tiles + si * in_bytes + G::A_TILE;
                const uint4 *ua = reinterpret_cast<const uint4 *>(ut + (size_t)(fi * n_mics + ma) * G::UROW);
                const uint4 *ub = reinterpret_cast<const uint4 *>(ut + (size_t)(fi * n_mics + mb) * G::UROW);
                uint8_t *dst = &s.b[st][(n >> 3) * 128 + (n & 7) * 16];
#pragma unroll
                for (int q = qh * (G::CB / 8); q < (qh + 1) * (G::CB / 8); q++) {      // 4 bins = 8 k values = one 16-byte core-matrix row
                    const uint4 xa = ua[q], xb = ub[q];
                    const uint32_t wa[4] = {xa.x, xa.y, xa.z, xa.w}, wb[4] = {xb.x, xb.y, xb.z, xb.w};
                    uint32_t o[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&wa[e]));
                        const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&wb[e]));
                        const __half2 gq = __floats2half2_rn(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);     // conj(a) b
                        o[e] = *reinterpret_cast<const uint32_t *>(&gq);
                    }
                    *reinterpret_cast<uint4 *>(dst + q * ((G::COLS / 8) * 128)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
            fence_proxy_async();          // the tile bytes -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) { mbar_arrive(&s.b_full[st]); mbar_arrive(&s.in_empty[si]); }
            if (++si == n_in) { si = 0; use++; }
        }
        // ---- accumulators: lane = row = lag index, columns = (frame, pair); first-max arg-max per column
        float gn = 0.f;                   // Nyquist bin of this column: real, enters with (-1)^s
        if (valid) {
            const float2 a = __half22float2(nyq[(size_t)f * n_mics + ma]), b = __half22float2(nyq[(size_t)f * n_mics + mb]);
            gn = a.x * b.x + a.y * b.y;
        }
        float *const gnyq = reinterpret_cast<float *>(&s.b[0][0]);      // the B tiles are free by now (acc_full below)
        mbar_wait(&s.acc_full, 0);
        tc_fence_after();
        if (qh == 0) gnyq[n] = gn;
        named_bar(1, 32 * G::CONV_WARPS);
        const int wq = warp & 3, part = (warp - 1) >> 2, row = wq * 32 + lane;      // lane quarter, quarter of the columns
        const bool rvalid = row <= 2 * L;
        const float sgn = ((row - L) & 1) ? -1.f : 1.f;
        for (int cb = 0; cb < 64; cb += 16) {
            const int c0 = part * 64 + cb;
            uint32_t v[16];
            tmem_ld16(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0, v);
            tmem_ld_wait();
            uint32_t mykey = 0; int myrow = 0;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const float y = __uint_as_float(v[j]) + sgn * gnyq[c0 + j];
                const uint32_t key = rvalid ? fkey(y) : 0u;
                const uint32_t mx = __reduce_max_sync(0xffffffffu, key);
                const int first = __ffs(__ballot_sync(0xffffffffu, key == mx)) - 1;
                if (lane == j) { mykey = mx; myrow = wq * 32 + first; }
            }
            if (lane < 16) s.part[wq][c0 + lane] = make_uint2(mykey, (uint32_t)myrow);
        }
        tc_fence_before();
        named_bar(1, 32 * G::CONV_WARPS);
        if (valid && qh == 0) {
            uint32_t bk = 0; int br = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) { const uint2 e = s.part[q][n]; if (e.x > bk) { bk = e.x; br = (int)e.y; } }     // strict: first maximum
            lags[(size_t)f * P + p] = br - L;
            if (peak) peak[(size_t)f * P + p] = fkey_inv(bk) / (float)N2;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == G::MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256));
}

} // namespace atk

using namespace atk;

// A tiles for (n_bits, L): chunk c holds bins 32c .. 32c+31 as 64 k values (2 per bin: cos, -sin), rows = lag index i = s + L
// (rows > 2L zero), in the UMMA K-major no-swizzle canonical layout [k / 8][row / 8][row % 8][k % 8].
void at_gccphat_dft_tiles(int n_bits, int L, uint16_t *h_out /* fp16 bits, (N / 32) * 8192 entries */)
{
    const int N = 1 << n_bits, N2 = 2 * N;
    for (int c = 0; c < N / 32; c++)
        for (int k = 0; k < 64; k++) {
            const int m = 32 * c + k / 2;
            const double cm = m == 0 ? 1.0 : 2.0;
            for (int i = 0; i < 128; i++) {
                double v = 0.0;
                if (i <= 2 * L) {
                    const int s = i - L;
                    const long long ph = ((long long)m * s) % N2;          // exact phase index
                    const double ang = 2.0 * M_PI * (double)ph / (double)N2;
                    v = (k & 1) ? -cm * sin(ang) : cm * cos(ang);
                }
                const __half h = __float2half_rn((float)v);
                h_out[(size_t)c * 8192 + (size_t)(((k >> 3) * 16 + (i >> 3)) * 8 + (i & 7)) * 8 + (k & 7)] = *reinterpret_cast<const uint16_t *>(&h);
            }
        }
}

int at_gccphat_dft_ps_log2(int n_mics)      // pair slots per frame (power of two, at least 4)
{
    const int P = n_mics * (n_mics - 1) / 2;
    int l = 2;
    while ((1 << l) < P) l++;
    return l;
}

cudaError_t at_launch_gccphat_dft(int n_mics, int n_bits, int L, size_t n_frames, const void *d_a_tiles, const void *d_spec_tiled,
                                  const void *d_nyq, int32_t *d_lags, float *d_peak, cudaStream_t st)
{
    if (!n_frames) return cudaSuccess;
    if (L > 63 || n_mics < 2 || n_mics > 8) return cudaErrorInvalidValue;
    const int psl = at_gccphat_dft_ps_log2(n_mics), FG = GccDftGeo::COLS >> psl;
    const unsigned groups = (unsigned)((n_frames + FG - 1) / FG);
    const size_t fixed = (sizeof(GccDftSmem) + 127) & ~(size_t)127, stage = GccDftGeo::A_TILE + (size_t)FG * n_mics * GccDftGeo::UROW;
    int n_in = (int)(((size_t)224 * 1024 - fixed) / stage);                // as many input stages as fit
    if (n_in > GccDftSmem::MAX_IN) n_in = GccDftSmem::MAX_IN;
    if (n_in < 2) return cudaErrorInvalidValue;
    const int smem = (int)(fixed + (size_t)n_in * stage);
    cudaError_t e;
#define AT_DFT(NB)                                                                                                        \
    {                                                                                                                     \
        if ((e = cudaFuncSetAttribute(gcc_dft_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e; \
        gcc_dft_kernel<NB><<<groups, GccDftGeo::THREADS, smem, st>>>((const __half *)d_a_tiles, (const __half2 *)d_spec_tiled,         \
                                                                    (const __half2 *)d_nyq, n_mics, L, psl, n_in, (unsigned)n_frames, d_lags, d_peak); \
    }
    if (n_bits == 10) AT_DFT(10)
    else if (n_bits == 12) AT_DFT(12)
    else return cudaErrorInvalidValue;
#undef AT_DFT
    at_count_launch(1);
    return cudaGetLastError();
}
