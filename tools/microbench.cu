// tools/microbench.cu -- issue-rate microbenchmarks for the integer ops the kernels lean on (sm_100a).
// Prints lane-ops per clock per SM for each op.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
template <int OP>
__global__ void __launch_bounds__(1024) k(unsigned* out, unsigned seed)
{
    unsigned a[8], b = seed + threadIdx.x, c = seed * 3 + 1;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 17 + i + seed;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) asm volatile("vabsdiff.s32.s32.s32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 1) asm volatile("add.s16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == 2) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 3) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, 0x6420;" : "+r"(a[i]) : "r"(b));
            if (OP == 5) asm volatile("shr.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(c & 3));
            if (OP == 6) asm volatile("abs.s32 %0, %0;" : "+r"(a[i]));
            if (OP == 7) asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 8) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 9) asm volatile("max.s16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == 10) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(1024) k_imma(int* out, unsigned seed)
{
    int d[4][4] = {};
    unsigned a[4] = { seed, seed + 1, seed + 2, seed + 3 }, b0 = seed * 5, b1 = seed * 7;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(d[i][0]), "+r"(d[i][1]), "+r"(d[i][2]), "+r"(d[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
    int s = 0;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float timeit(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, ctas = sms * 2;
    unsigned* out; cudaMalloc(&out, (size_t)ctas * 1024 * 4);
    const char* names[] = { "VABSDIFF(+acc)", "VIADD.16x2", "IMAD", "IADD", "PRMT", "SHF.R.S32", "IABS", "IDP.4A", "IDP.2A", "VIMNMX.S16x2", "LOP3" };
    printf("%s: %d SMs, max clock %d MHz (rates below assume the max clock; real clock may be lower)\n", p.name, sms, clk / 1000);
#define RUN(OP) { float ms = timeit([&] { k<OP><<<ctas, 1024>>>(out, 1234u); }); \
      double ops = (double)ctas * 1024 * ITERS * 8; printf("%-16s %7.3f ms  %6.1f lane-ops/clk/SM\n", names[OP], ms, ops / (ms * 1e-3) / (clk * 1e3) / sms); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10)
    {
        float ms = timeit([&] { k_imma<<<ctas, 1024>>>((int*)out, 77u); });
        double mmas = (double)ctas * 32 * ITERS * 4;
        printf("IMMA.16832       %7.3f ms  %6.1f MAC/clk/SM  (%.1f cycles per IMMA per SMSP)\n", ms, mmas * 4096 / (ms * 1e-3) / (clk * 1e3) / sms,
               (ms * 1e-3) * (clk * 1e3) * sms * 4 / mmas);
    }
    return 0;
}
