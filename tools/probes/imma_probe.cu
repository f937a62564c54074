// IMMA issue-rate probe for sm_100a: how fast can legacy mma.sync m16n8k32 int8 issue under the
// conditions of the localization loop (few warps, distinct operands, mixed signedness, chains)?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_probe imma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#define MMA(TA, TB, c, a, b)                                                                                    \
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32." #TA "." #TB ".s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "     \
                 "{%8,%9}, {%0,%1,%2,%3};"                                                                      \
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])                                               \
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]))

template <int MODE>
__global__ void __launch_bounds__(128) probe(int iters, int seed, int *sink)
{
    unsigned A[4][4], B[4][2];
    int c[12][4];
    int x[8]; float xf[8];
    __shared__ int sbuf[1024];
    sbuf[threadIdx.x] = seed; sbuf[threadIdx.x + 128] = seed;
    __syncthreads();
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(&sbuf[threadIdx.x & 31]);
    const unsigned saddr16 = (unsigned)__cvta_generic_to_shared(&sbuf[4 * (threadIdx.x & 31)]);
    const unsigned saddr8 = (unsigned)__cvta_generic_to_shared(&sbuf[2 * (threadIdx.x & 31)]);
    for (int i = 0; i < 8; i++) { x[i] = seed + i; xf[i] = seed + i; }
    for (int i = 0; i < 4; i++) {
        for (int j = 0; j < 4; j++) A[i][j] = seed * (i * 4 + j + 1) + threadIdx.x;
        for (int j = 0; j < 2; j++) B[i][j] = seed * (i * 7 + j + 3) + threadIdx.x * 3;
    }
    for (int i = 0; i < 12; i++)
        for (int j = 0; j < 4; j++) c[i][j] = 0;
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {          // 12 independent tiles, same operands, s8.s8
#pragma unroll
            for (int i = 0; i < 12; i++) MMA(s8, s8, c[i], A[0], B[0]);
        } else if (MODE == 1) {   // 12 independent tiles, 4 x 4 distinct operands, s8.s8
#pragma unroll
            for (int i = 0; i < 12; i++) MMA(s8, s8, c[i], A[i & 3], B[(i >> 2) + (i & 1)]);
        } else if (MODE == 2) {   // the kernel's pattern: 9 tiles, mixed signs, MID tiles hit twice
            MMA(s8, s8, c[0], A[0], B[0]); MMA(s8, u8, c[1], A[0], B[1]); MMA(u8, u8, c[2], A[1], B[1]);
            MMA(s8, s8, c[3], A[2], B[0]); MMA(s8, u8, c[4], A[2], B[1]); MMA(u8, u8, c[5], A[3], B[1]);
            MMA(s8, s8, c[6], A[2], B[2]); MMA(s8, u8, c[7], A[2], B[3]); MMA(u8, u8, c[8], A[3], B[3]);
            MMA(u8, s8, c[1], A[1], B[0]); MMA(u8, s8, c[4], A[3], B[0]); MMA(u8, s8, c[7], A[3], B[2]);
        } else if (MODE == 3) {   // same pattern, all s8.s8
            MMA(s8, s8, c[0], A[0], B[0]); MMA(s8, s8, c[1], A[0], B[1]); MMA(s8, s8, c[2], A[1], B[1]);
            MMA(s8, s8, c[3], A[2], B[0]); MMA(s8, s8, c[4], A[2], B[1]); MMA(s8, s8, c[5], A[3], B[1]);
            MMA(s8, s8, c[6], A[2], B[2]); MMA(s8, s8, c[7], A[2], B[3]); MMA(s8, s8, c[8], A[3], B[3]);
            MMA(s8, s8, c[1], A[1], B[0]); MMA(s8, s8, c[4], A[3], B[0]); MMA(s8, s8, c[7], A[3], B[2]);
        } else if (MODE == 4) {   // 12 tiles, mixed signs, no double hits
            MMA(s8, s8, c[0], A[0], B[0]); MMA(s8, u8, c[1], A[0], B[1]); MMA(u8, u8, c[2], A[1], B[1]);
            MMA(s8, s8, c[3], A[2], B[0]); MMA(s8, u8, c[4], A[2], B[1]); MMA(u8, u8, c[5], A[3], B[1]);
            MMA(s8, s8, c[6], A[2], B[2]); MMA(s8, u8, c[7], A[2], B[3]); MMA(u8, u8, c[8], A[3], B[3]);
            MMA(u8, s8, c[9], A[1], B[0]); MMA(u8, s8, c[10], A[3], B[0]); MMA(u8, s8, c[11], A[3], B[2]);
        }
        if (MODE >= 5) {          // kernel pattern + 24 extra integer instructions per 12 IMMA
            MMA(s8, s8, c[0], A[0], B[0]); MMA(s8, u8, c[1], A[0], B[1]); MMA(u8, u8, c[2], A[1], B[1]);
            MMA(s8, s8, c[3], A[2], B[0]); MMA(s8, u8, c[4], A[2], B[1]); MMA(u8, u8, c[5], A[3], B[1]);
            MMA(s8, s8, c[6], A[2], B[2]); MMA(s8, u8, c[7], A[2], B[3]); MMA(u8, u8, c[8], A[3], B[3]);
            MMA(u8, s8, c[1], A[1], B[0]); MMA(u8, s8, c[4], A[3], B[0]); MMA(u8, s8, c[7], A[3], B[2]);
#pragma unroll
            for (int k = 0; k < 24; k++) {
                if (MODE == 5) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[k & 7]) : "r"(seed), "r"(it));       // FMA pipe (IMAD)
                if (MODE == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k & 7]) : "r"(seed), "r"(it)); // ALU pipe
                if (MODE == 7) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(xf[k & 7]) : "f"(1.0001f), "f"(0.5f));  // FFMA
                if (MODE == 8) asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(x[k & 7]) : "r"(saddr + 4 * (k & 7)));
            }
            if (MODE == 9) {     // 4 ldmatrix.x4 per 12 IMMA
#pragma unroll
                for (int k = 0; k < 4; k++)
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]) : "r"(saddr16 + 512 * k));
            }
            if (MODE == 10) {    // 12 LDS.64 per 12 IMMA
#pragma unroll
                for (int k = 0; k < 12; k++)
                    asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x[k & 3]), "=r"(x[4 + (k & 3)]) : "r"(saddr8 + 256 * (k & 7)));
            }
            if (MODE == 11) {    // 6 LDS.128 per 12 IMMA
#pragma unroll
                for (int k = 0; k < 6; k++)
                    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]) : "r"(saddr16 + 512 * k));
            }
            if (MODE == 12) {    // 24 MOV-like (IMAD.MOV is what ptxas emits) -> use mov via prmt to stay on ALU
#pragma unroll
                for (int k = 0; k < 24; k++) asm volatile("prmt.b32 %0, %1, %2, 0x3210;" : "=r"(x[k & 7]) : "r"(x[(k + 1) & 7]), "r"(it));
            }
        }
        if (MODE == 13) {    // 48 x m8n8k16 (same MACs as 12 x m16n8k32)
#pragma unroll
            for (int i = 0; i < 48; i++)
                asm volatile("mma.sync.aligned.m8n8k16.row.col.s32.s8.s8.s32 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+r"(c[i % 12][(i / 12) & 1]), "+r"(c[i % 12][2 + ((i / 12) & 1)]) : "r"(A[i & 3][i & 3]), "r"(B[i & 3][i & 1]));
        }
        if (MODE == 14) {    // 24 x m16n8k16 (same MACs as 12 x m16n8k32)
#pragma unroll
            for (int i = 0; i < 24; i++)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+r"(c[i % 12][0]), "+r"(c[i % 12][1]), "+r"(c[i % 12][2]), "+r"(c[i % 12][3])
                             : "r"(A[i & 3][0]), "r"(A[i & 3][1]), "r"(B[i & 3][i & 1]));
        }
        // operands change every iteration, like freshly loaded fragments
        A[it & 3][it & 3] += it; B[(it + 1) & 3][it & 1] ^= it;
    }
    int s = 0;
    for (int i = 0; i < 12; i++)
        for (int j = 0; j < 4; j++) s += c[i][j];
    for (int i = 0; i < 8; i++) s += x[i] + (int)xf[i];
    if (s == 0x1234567) sink[0] = s;
}

template <int MODE>
void run(const char *name, int warps_per_sm, int sms)
{
    int *d;
    cudaMalloc(&d, 4);
    const int iters = 2000, blocks = sms * warps_per_sm / 4;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    probe<MODE><<<blocks, 128>>>(10, 1, d);
    cudaEventRecord(a);
    probe<MODE><<<blocks, 128>>>(iters, 1, d);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double immas = 12.0 * iters * blocks * 4;
    printf("%-46s warps/SM %2d: %.2f clk per IMMA per SM (at 1.965 GHz)  %s\n", name, warps_per_sm,
           ms * 1e-3 * 1.965e9 / (immas / sms), cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    for (int w : {16}) {
        run<0>("12 tiles, same operands, s8.s8", w, sms);
        run<1>("12 tiles, distinct operands, s8.s8", w, sms);
        run<3>("kernel pattern (9 tiles, 3 hit twice), s8.s8", w, sms);
        run<2>("kernel pattern, mixed signs", w, sms);
        run<4>("12 tiles, mixed signs, no double hit", w, sms);
        run<5>("kernel pattern + 24 IMAD", w, sms);
        run<6>("kernel pattern + 24 LOP3 (ALU)", w, sms);
        run<7>("kernel pattern + 24 FFMA", w, sms);
        run<8>("kernel pattern + 24 LDS.32", w, sms);
        run<9>("kernel pattern + 4 LDSM.x4", w, sms);
        run<10>("kernel pattern + 12 LDS.64", w, sms);
        run<11>("kernel pattern + 6 LDS.128", w, sms);
        run<12>("kernel pattern + 24 PRMT", w, sms);
        run<13>("48 x m8n8k16 (= MACs of 12 x m16n8k32; per 12-group)", w, sms);
        run<14>("24 x m16n8k16 (= MACs of 12 x m16n8k32; per 12-group)", w, sms);
    }
    return 0;
}
