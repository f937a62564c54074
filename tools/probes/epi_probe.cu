// epi_probe.cu -- round-2 probe for the tcgen05 form of the reference shape (3 mics x 1024 samples):
//   (1) TMEM read throughput (tcgen05.ld 32x32b.x16) with 4..16 warps per SM,
//   (2) the register butterfly that turns a 128 x 16 polyphase tile into its diagonal sums (correctness + cycles),
//   (3) the ten-MMA frame (Hankel A, N = 64/32/16) issued back to back from rotating plane slots, alone and with
//       other warps generating shared-memory traffic.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o epi_probe epi_probe.cu ; run on a B200.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- diagonal sums of a 32-row x 16-phase block held one row per lane: entry (row, phi) belongs to lag row - phi.
// Five exchange stages (lane ^ 1, 2, 4, 8, 16); after stage k a lane holds the lags congruent to it modulo 2^(k+1).
// Returns n0 = sum for lag (lane) and n1 = partial sum for lag (lane - 32) (non-zero for lanes >= 17 only).
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

// mode 0: TMEM read throughput: reps x (LOADS x ld16, wait).  mode 1: reps x (2 x ld16, pack 256 a + b, butterfly).
// mode 2: correctness of the butterfly: the tile is written with tcgen05.st from gin[128][16], result -> gout[2][128].
template <int LOADS>
__global__ void __launch_bounds__(512) epi_kernel(int mode, int reps, const int *gin, int *gout, long long *cycles)
{
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wq = warp & 3;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    const uint32_t ta = tmem + ((uint32_t)(wq * 32) << 16);
    {   // initialise the first 128 columns of this warp's lanes (each quarter is written by its first warp)
        if (warp < 4) {
            uint32_t v[16];
            for (int c = 0; c < 512; c += 16) {
#pragma unroll
                for (int j = 0; j < 16; j++) v[j] = (mode == 2 && c < 16) ? (uint32_t)gin[(wq * 32 + lane) * 16 + j] : (uint32_t)(c + j + lane);
                tmem_st16(ta + c, v);
            }
            tmem_st_wait();
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    int sink = 0;
    const long long t0 = clock64();
    if (mode == 0) {
        for (int r = 0; r < reps; r++) {
            uint32_t v[LOADS][16];
#pragma unroll
            for (int k = 0; k < LOADS; k++) tmem_ld16(ta + ((r * LOADS + k) * 16 & 511), v[k]);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < LOADS; k++) sink += (int)(v[k][0] ^ v[k][7] ^ v[k][15]);
        }
    } else if (mode == 1) {
        for (int r = 0; r < reps; r++) {
            uint32_t h[16], m[16];
            tmem_ld16(ta + ((r * 32) & 511), h);
            tmem_ld16(ta + ((r * 32 + 16) & 511), m);
            tmem_ld_wait();
            Arr<16> u;
#pragma unroll
            for (int j = 0; j < 16; j++) u.v[j] = (int)h[j] * 256 + (int)m[j];
            int n0, n1;
            diag_butterfly(u, lane, n0, n1);
            sink += n0 ^ n1;
        }
    } else {
        uint32_t h[16];
        tmem_ld16(ta, h);
        tmem_ld_wait();
        Arr<16> u;
#pragma unroll
        for (int j = 0; j < 16; j++) u.v[j] = (int)h[j];
        int n0, n1;
        diag_butterfly(u, lane, n0, n1);
        if (warp < 4) { gout[wq * 32 + lane] = n0; gout[128 + wq * 32 + lane] = n1; }
    }
    const long long t1 = clock64();
    __shared__ long long tmax;
    if (tid == 0) tmax = 0;
    __syncthreads();
    atomicMax((unsigned long long *)&tmax, (unsigned long long)(t1 - t0));
    __syncthreads();
    if (tid == 0 && cycles) cycles[blockIdx.x] = tmax;
    if (sink == 0x12345678 && gout) gout[300] = sink;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

// ---------------------------------------------------------------- MMA sequence
__device__ __forceinline__ void umma_i8_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
                 "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, {%7, %7, %7, %7}, p;\n\t}\n"
                 :: "r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__host__ __device__ constexpr uint32_t umma_idesc(int n)
{
    return (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}\n"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

constexpr int PLANE = 1152, SLOTS = 7, FRAMEB = 6 * PLANE;
// variant 0: the ten-MMA frame (12 products, mid accumulated in TMEM); variant 1: eight MMAs (no l.l: 9 products);
// variant 2: eight MMAs with four separate tiles per pair (N = 64/64/32/32, no accumulate-onto)
// traffic: the other (blockDim/32 - 1) warps copy 16-byte words through shared memory all the time (0 = idle)
__global__ void __launch_bounds__(512) mma_kernel(int variant, int frames, int traffic, long long *cycles)
{
    extern __shared__ __align__(128) uint8_t dyn[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    __shared__ int stop;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (SLOTS * FRAMEB + 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(dyn)[i] = (uint32_t)i * 2654435761u;
    if (tid == 0) { mbar_init(&bar, 1); stop = 0; }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (warp == 0) {
        const long long t0 = clock64();
        if (tid == 0) {
            constexpr uint32_t LBO = (128u >> 4) << 16;
            constexpr uint32_t HI_A = (16u >> 4) | 0x4000u;
            constexpr uint32_t HI_B1 = ((uint32_t)PLANE >> 4) | 0x4000u, HI_B2 = ((uint32_t)(2 * PLANE) >> 4) | 0x4000u;
            constexpr uint32_t I64 = umma_idesc(64), I32 = umma_idesc(32), I16 = umma_idesc(16);
            for (int f = 0; f < frames; f++) {
                const uint32_t b16 = (smem_u32(dyn + (f % SLOTS) * FRAMEB) >> 4) + LBO;
                const uint32_t cb = tmem + (uint32_t)(f % 3) * 160;
                auto pl = [&](int k, int kk) { return b16 + (uint32_t)((k * PLANE + 512 * kk) >> 4); };
                auto xb = [&](int k, int kk) { return b16 + (uint32_t)((k * PLANE + 48 + 512 * kk) >> 4); };
                // planes: 0 a.h, 1 b.h, 2 a.l, 3 b.l, 4 c.h, 5 c.l
                if (variant == 0) {
                    umma_i8_lohi(cb + 0, pl(4, 0), HI_A, xb(0, 0), HI_B1, I64, 0);
                    umma_i8_lohi(cb + 32, pl(5, 0), HI_A, xb(0, 0), HI_B1, I32, 1);
                    umma_i8_lohi(cb + 64, pl(5, 0), HI_A, xb(2, 0), HI_B1, I32, 0);
                    umma_i8_lohi(cb + 96, pl(1, 0), HI_A, xb(0, 0), HI_B2, I32, 0);
                    umma_i8_lohi(cb + 112, pl(3, 0), HI_A, xb(0, 0), HI_B2, I16, 1);
                    umma_i8_lohi(cb + 128, pl(3, 0), HI_A, xb(2, 0), HI_B2, I16, 0);
                    umma_i8_lohi(cb + 0, pl(4, 1), HI_A, xb(0, 1), HI_B1, I64, 1);
                    umma_i8_lohi(cb + 32, pl(5, 1), HI_A, xb(0, 1), HI_B1, I64, 1);
                    umma_i8_lohi(cb + 96, pl(1, 1), HI_A, xb(0, 1), HI_B2, I32, 1);
                    umma_i8_lohi(cb + 112, pl(3, 1), HI_A, xb(0, 1), HI_B2, I32, 1);
                } else if (variant == 1) {
                    for (int kk = 0; kk < 2; kk++) {
                        umma_i8_lohi(cb + 0, pl(4, kk), HI_A, xb(0, kk), HI_B1, I64, kk);
                        umma_i8_lohi(cb + 32, pl(5, kk), HI_A, xb(0, kk), HI_B1, I32, 1);
                        umma_i8_lohi(cb + 96, pl(1, kk), HI_A, xb(0, kk), HI_B2, I32, kk);
                        umma_i8_lohi(cb + 112, pl(3, kk), HI_A, xb(0, kk), HI_B2, I16, 1);
                    }
                } else {
                    for (int kk = 0; kk < 2; kk++) {
                        umma_i8_lohi(cb + 0, pl(4, kk), HI_A, xb(0, kk), HI_B1, I64, kk);
                        umma_i8_lohi(cb + 64, pl(5, kk), HI_A, xb(0, kk), HI_B1, I64, kk);
                        umma_i8_lohi(cb + 128, pl(1, kk), HI_A, xb(0, kk), HI_B2, I32, kk);
                        umma_i8_lohi(cb + 160, pl(3, kk), HI_A, xb(0, kk), HI_B2, I32, kk);
                    }
                }
            }
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (tid == 0) { cycles[blockIdx.x] = t1 - t0; *(volatile int *)&stop = 1; }
    } else if (traffic) {
        // every other warp: read 16 bytes, write 16 bytes in a scratch region after the planes (conflict-free)
        uint4 *scr = reinterpret_cast<uint4 *>(dyn + SLOTS * FRAMEB);
        const int idx = (tid - 32) & 1023;
        uint4 v = scr[idx];
        while (!*(volatile int *)&stop) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (traffic & 1) { const uint4 w = scr[(idx + 32 * k) & 1023]; v.x ^= w.x; v.y += w.y; v.z ^= w.z; v.w += w.w; }
                if (traffic & 2) scr[(idx + 32 * k + 512) & 1023] = v;
            }
        }
        if (v.x == 0x12345678u) cycles[blockIdx.x + 148] = v.x;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

static double avg_cycles(long long *dc, int n)
{
    std::vector<long long> c(n);
    CK(cudaMemcpy(c.data(), dc, sizeof(long long) * n, cudaMemcpyDeviceToHost));
    double a = 0; for (auto v : c) a += (double)v;
    return a / n;
}

int main()
{
    long long *dc; int *gin, *gout;
    CK(cudaMalloc(&dc, sizeof(long long) * 296)); CK(cudaMalloc(&gin, sizeof(int) * 128 * 16)); CK(cudaMalloc(&gout, sizeof(int) * 512));
    // (2) correctness
    std::vector<int> D(128 * 16);
    uint32_t st = 777;
    for (auto &v : D) { st = st * 1664525u + 1013904223u; v = (int)(st >> 8) - (1 << 23); }
    CK(cudaMemcpy(gin, D.data(), sizeof(int) * D.size(), cudaMemcpyHostToDevice));
    epi_kernel<1><<<1, 128>>>(2, 1, gin, gout, nullptr);
    CK(cudaDeviceSynchronize());
    std::vector<int> out(256);
    CK(cudaMemcpy(out.data(), gout, sizeof(int) * 256, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int j = -15; j < 128; j++) {
        long ref = 0;
        for (int phi = 0; phi < 16; phi++) { const int m = j + phi; if (m >= 0 && m < 128) ref += D[m * 16 + phi]; }
        long got = 0;
        if (j >= 0) got += out[j];
        if (j + 32 < 128 && ((j + 32) & 31) >= 17) got += out[128 + j + 32];
        if ((int)ref != (int)got) { if (bad < 8) printf("lag %d: got %ld want %ld\n", j, got, ref); bad++; }
    }
    printf("butterfly diagonal sums: %ld mismatches of 143 lags\n", bad);
    // (1), (2) timing
    for (int nw : {4, 8, 12, 16}) {
        const int reps = 2000;
        epi_kernel<4><<<148, nw * 32>>>(0, reps, gin, gout, dc);
        CK(cudaDeviceSynchronize());
        const double c0 = avg_cycles(dc, 148);
        printf("LDTM  %2d warps/SM: %7.1f cycles per (4 x ld16 + wait) per warp-iteration -> %6.1f B/clk/SM\n", nw, c0 / reps,
               (double)nw * 4 * 2048 * reps / c0);
        epi_kernel<8><<<148, nw * 32>>>(0, reps, gin, gout, dc);
        CK(cudaDeviceSynchronize());
        const double c2 = avg_cycles(dc, 148);
        printf("LDTM  %2d warps/SM: %7.1f cycles per (8 x ld16 + wait) per warp-iteration -> %6.1f B/clk/SM\n", nw, c2 / reps,
               (double)nw * 8 * 2048 * reps / c2);
        epi_kernel<1><<<148, nw * 32>>>(1, reps, gin, gout, dc);
        CK(cudaDeviceSynchronize());
        const double c1 = avg_cycles(dc, 148);
        printf("BFLY  %2d warps/SM: %7.1f cycles per (2 x ld16, pack, butterfly) per warp-iteration -> %6.1f cycles per warp-tile-pair per SM\n",
               nw, c1 / reps, c1 / reps / nw);
    }
    // (3) MMA sequences
    CK(cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SLOTS * FRAMEB + 16384));
    const char *vn[] = {"10 MMAs (12 products, mid accumulated)", "8 MMAs (9 products)", "8 MMAs (4 tiles per pair)"};
    for (int variant = 0; variant < 3; variant++)
        for (int traffic = 0; traffic < 4; traffic++)
            for (int nw : {1, 8, 16}) {
                if (traffic == 0 && nw != 1) continue;
                if (traffic != 0 && nw == 1) continue;
                const int frames = 2000;
                mma_kernel<<<148, nw * 32, SLOTS * FRAMEB + 16384>>>(variant, frames, traffic, dc);
                CK(cudaDeviceSynchronize());
                printf("MMA   %-40s traffic %d (%2d warps): %7.1f cycles per frame\n", vn[variant], traffic, nw, avg_cycles(dc, 148) / frames);
            }
    return 0;
}
