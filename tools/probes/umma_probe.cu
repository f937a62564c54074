// umma_probe.cu -- feasibility probe for a tcgen05 (UMMA) form of the lagged cross-correlation.
//
// Polyphase formulation: with time split as i = 16 q + phi, the int8 product
//     D[m][phi] = sum_q Y[m + 16 q] * X[phi + 16 q]          (m = 0..127, phi = 0..15, q = 0..63)
// is a GEMM whose operands are the PLAIN byte planes in shared memory, read as MN-major UMMA operands:
//   A[m][q] = Y[m + 16 q]   -- a Hankel matrix: MN chunks of 16 bytes 16 bytes apart (SBO = 16), K rows 16 bytes apart,
//                              i.e. the chunks OVERLAP in memory; nothing is materialised,
//   B[n][q] = X[n + 16 q]   -- 16 phases per plane, several planes side by side (SBO = plane stride).
// The correlation at lag s is the diagonal sum  corr[s] = sum_phi D[s + PAD + phi][phi]  (done by CUDA cores).
// One tcgen05.mma kind::i8 (K = 32) covers 512 samples: 2 MMAs per (y plane, x plane group) and frame.
//
// The probe (1) checks D against a CPU evaluation for M = 128, N = 64 (4 x planes), K = 64, which validates MN-major
// int8 operands, overlapping descriptors and the TMEM read-back; (2) times back-to-back MMAs of that shape, the TMEM
// read-out and a shared-memory atomic diagonal reduction, the pieces of a performance model.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu ; run on a B200.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int PLANE = 1152;          // bytes per plane buffer (multiple of 16)
constexpr int NXP = 4;               // x planes side by side -> N = 64
constexpr int M = 128, N = 16 * NXP;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, SWIZZLE_NONE; fields in 16-byte units (cute/arch/mma_sm100_desc.hpp, SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // version 1 (sm_100)
    return d;
}
// instruction descriptor (InstrDescriptor): S32 accumulate, signed int8 A and B, both MN-major
__host__ __device__ constexpr uint32_t make_idesc(int m, int n)
{
    return (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u));
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}\n"
                 :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}

struct Smem {
    alignas(128) int8_t y[PLANE + 128];     // zero-padded y plane (A operand reads up to y[127 + 16*63 + 15])
    alignas(128) int8_t x[NXP][PLANE];      // x planes
    alignas(8) uint64_t bar;
    uint32_t tmem_base;
    alignas(16) int diag[3][128];           // diagonal-sum scratch for the timing mode
};

// mode 0: correctness (D -> global).  mode 1: time `reps` x (2 MMAs, commit, wait).  mode 2: + TMEM read-out of 64 columns.
// mode 3: + shared-memory atomic diagonal reduction of the 64 columns.
__global__ void __launch_bounds__(128) probe(const int8_t *gy, const int8_t *gx, int *gd, int mode, int reps, long long *cycles)
{
    __shared__ Smem s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < PLANE + 128; i += 128) s.y[i] = gy[i];
    for (int i = tid; i < NXP * PLANE; i += 128) (&s.x[0][0])[i] = gx[i];
    for (int i = tid; i < 3 * 128; i += 128) (&s.diag[0][0])[i] = 0;
    if (tid == 0) mbar_init(smem_u32(&s.bar), 1);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s.tmem_base)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s.tmem_base;
    const uint32_t bar = smem_u32(&s.bar);
    constexpr uint32_t idesc = make_idesc(M, N);
    const uint32_t ya = smem_u32(s.y), xa = smem_u32(&s.x[0][0]);

    long long t0 = clock64();
    uint32_t parity = 0;
    int sink = 0;
    for (int r = 0; r < reps; r++) {
        if (tid == 0) {
#pragma unroll
            for (int kk = 0; kk < 2; kk++) {          // K = 64 = 2 MMAs; 32 K rows of 16 bytes = 512 bytes per step
                const uint64_t ad = make_desc(ya + 512 * kk, /*LBO: K groups of 8 rows*/ 128, /*SBO: MN chunks*/ 16);
                const uint64_t bd = make_desc(xa + 512 * kk, 128, PLANE);
                umma_i8(tmem, ad, bd, idesc, kk);
            }
            umma_commit(bar);
        }
        mbar_wait(bar, parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (mode == 0 || mode >= 2) {
#pragma unroll
            for (int c = 0; c < N; c += 16) {
                uint32_t v[16];
                tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (mode == 0) {
                    for (int j = 0; j < 16; j++) gd[(warp * 32 + lane) * N + c + j] = (int)v[j];
                } else if (mode == 2) {
                    for (int j = 0; j < 16; j++) sink += (int)v[j];
                } else {
                    const int m = warp * 32 + lane;   // diagonal sum: entry (m, phi) belongs to lag index m - phi
                    for (int j = 0; j < 16; j++)
                        if (m - j >= 0) atomicAdd(&s.diag[(c >> 4) % 3][m - j], (int)v[j]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        if (mode == 3) __syncthreads();
    }
    long long t1 = clock64();
    if (tid == 0 && cycles) cycles[blockIdx.x] = t1 - t0;
    if (mode >= 2 && gd) gd[M * N + tid] = sink + s.diag[0][tid];
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(64));
}

// Throughput: `reps` back-to-back MMAs of shape M128 x N x K32 from one thread, one commit, one wait.
// sbo_a = 16: the overlapping Hankel operand; sbo_a = 128: a conventional (non-overlapping) MN-major operand.
template <int NN>
__global__ void __launch_bounds__(128) tput(const int8_t *gy, int reps, uint32_t sbo_a, uint32_t lbo_a, long long *cycles)
{
    extern __shared__ __align__(128) uint8_t dyn[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 32768; i += 128) dyn[i] = (uint8_t)gy[i % 1024];
    if (tid == 0) mbar_init(smem_u32(&bar), 1);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    constexpr uint32_t idesc = make_idesc(M, NN);
    const uint32_t a0 = smem_u32(dyn), b0 = smem_u32(dyn) + 8192;
    long long t0 = clock64();
    if (tid == 0) {
        const uint64_t ad = make_desc(a0, lbo_a, sbo_a), bd = make_desc(b0, 128, 1152);
        if (lbo_a & 1) {     // odd LBO flag: walk through different A and B addresses (no operand re-use between MMAs)
            const uint32_t lbo = lbo_a & ~1u;
            for (int r = 0; r < reps; r++)
                umma_i8(tmem, make_desc(a0 + (uint32_t)(r & 7) * 1024, lbo, sbo_a), make_desc(b0 + (uint32_t)(r & 3) * 4608, 128, 1152), idesc, 1);
        } else
        for (int r = 0; r < reps; r++) umma_i8(tmem, ad, bd, idesc, 1);
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256));
}
template <int NN>
static void run_tput(const int8_t *dy, long long *dc, const char *what, uint32_t sbo_a, uint32_t lbo_a)
{
    const int reps = 4000;
    CK(cudaFuncSetAttribute(tput<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    tput<NN><<<148, 128, 65536>>>(dy, reps, sbo_a, lbo_a, dc);
    CK(cudaDeviceSynchronize());
    std::vector<long long> c(148);
    CK(cudaMemcpy(c.data(), dc, sizeof(long long) * 148, cudaMemcpyDeviceToHost));
    double avg = 0; for (auto v : c) avg += (double)v; avg /= 148.0 * reps;
    printf("throughput  M128 N%-3d K32  %-34s %7.1f cycles per MMA  (floor 128*N/256 = %d)\n", NN, what, avg, 128 * NN / 256);
}

int main()
{
    std::vector<int8_t> y(PLANE + 128, 0), x(NXP * PLANE, 0);
    uint32_t st = 12345;
    auto rnd = [&]() { st = st * 1664525u + 1013904223u; return (int8_t)(st >> 24); };
    for (int i = 48; i < 48 + 1024; i++) y[i] = rnd();          // data region with zero pads around it
    for (int p = 0; p < NXP; p++) for (int i = 0; i < 1024; i++) x[p * PLANE + i] = rnd();
    int8_t *dy, *dx; int *dd; long long *dc;
    CK(cudaMalloc(&dy, y.size())); CK(cudaMalloc(&dx, x.size())); CK(cudaMalloc(&dd, sizeof(int) * (M * N + 128))); CK(cudaMalloc(&dc, sizeof(long long) * 148));
    CK(cudaMemcpy(dy, y.data(), y.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dx, x.data(), x.size(), cudaMemcpyHostToDevice));
    probe<<<1, 128>>>(dy, dx, dd, 0, 1, nullptr);
    CK(cudaDeviceSynchronize());
    std::vector<int> d(M * N);
    CK(cudaMemcpy(d.data(), dd, sizeof(int) * M * N, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            long e = 0;
            for (int q = 0; q < 64; q++) e += (long)y[m + 16 * q] * (long)x[(n >> 4) * PLANE + (n & 15) + 16 * q];
            if (e != d[m * N + n]) { if (bad < 8) printf("mismatch m=%d n=%d got %d want %ld\n", m, n, d[m * N + n], e); bad++; }
        }
    printf("correctness: %ld mismatches of %d (M=%d N=%d K=64, MN-major int8, overlapping Hankel A)\n", bad, M * N, M, N);
    // correlation through the diagonal sums, against the direct form (PAD = 48)
    if (!bad) {
        long worst = 0;
        for (int s = -46; s <= 46; s++) {
            long viaD = 0, direct = 0;
            for (int phi = 0; phi < 16; phi++) viaD += d[(s + 48 + phi) * N + phi];
            for (int i = 0; i < 1024; i++) { const int j = i + s; if (j >= 0 && j < 1024) direct += (long)x[i] * (long)y[48 + j]; }
            if (viaD != direct) worst++;
        }
        printf("diagonal sums vs direct correlation of plane 0 over lags -46..46: %ld mismatches\n", worst);
    }
    const char *names[] = {"", "2 MMAs (M128 N64 K32) + commit + wait", "+ TMEM read-out of 64 columns", "+ shared-atomic diagonal reduction"};
    for (int mode = 1; mode <= 3; mode++) {
        const int reps = 2000;
        probe<<<148, 128>>>(dy, dx, dd, mode, reps, dc);
        CK(cudaDeviceSynchronize());
        std::vector<long long> c(148);
        CK(cudaMemcpy(c.data(), dc, sizeof(long long) * 148, cudaMemcpyDeviceToHost));
        double avg = 0; for (auto v : c) avg += (double)v; avg /= 148.0 * reps;
        printf("mode %d  %-45s %8.1f cycles per repetition (one CTA per SM, serialised)\n", mode, names[mode], avg);
    }
    run_tput<16>(dy, dc, "A overlapping Hankel (SBO 16)", 16, 128);
    run_tput<32>(dy, dc, "A overlapping Hankel (SBO 16)", 16, 128);
    run_tput<64>(dy, dc, "A overlapping Hankel (SBO 16)", 16, 128);
    run_tput<128>(dy, dc, "A overlapping Hankel (SBO 16)", 16, 128);
    run_tput<256>(dy, dc, "A overlapping Hankel (SBO 16)", 16, 128);
    run_tput<16>(dy, dc, "A disjoint chunks (SBO 512, LBO 128)", 512, 128);
    run_tput<64>(dy, dc, "A disjoint chunks (SBO 512, LBO 128)", 512, 128);
    run_tput<256>(dy, dc, "A disjoint chunks (SBO 512, LBO 128)", 512, 128);
    run_tput<64>(dy, dc, "Hankel A, new A and B every MMA", 16, 129);
    run_tput<32>(dy, dc, "Hankel A, new A and B every MMA", 16, 129);
    run_tput<128>(dy, dc, "Hankel A, new A and B every MMA", 16, 129);
    run_tput<16>(dy, dc, "A canonical (SBO 128, LBO 1024)", 128, 1024);
    run_tput<64>(dy, dc, "A canonical (SBO 128, LBO 1024)", 128, 1024);
    return 0;
}
