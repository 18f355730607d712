// tools/mix_bench2.cu -- which form of "sum_k max(T[k], C[k])" issues fastest on sm_100a?  All variants read C from shared
// memory with the kernel's access pattern (LDS.128, 8 lanes share an address), two positions per thread, 15 warps per SM.
//   ACC 0: (m0+m1)+(m2+m3) -> IDP.2A       ACC 1: IDP.2A per word      ACC 2: m0+m1+m2 (IADD3) -> IDP.2A
//   ACC 3: m0+m1+m2 -> IMMA (ones matrix)  MAXOP 0: VIMNMX.S16x2       MAXOP 1: HMNMX2 (max.f16x2 on the same bit patterns)
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../x266_b200/csrc/satd_packed.h"
using namespace x266::s3;
#define ITERS 2048

template <int MAXOP> __device__ __forceinline__ unsigned mx(unsigned a, unsigned b)
{
    if (MAXOP == 0) return vmax2(a, b);
    unsigned r;
    asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}

__device__ __forceinline__ void imma(int (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int ACC, int MAXOP>
__global__ void __launch_bounds__(160, 3) mix(unsigned* out, unsigned* cyc, unsigned seed)
{
    __shared__ __align__(16) unsigned tc[16][36];
    for (int i = threadIdx.x; i < 16 * 36; i += blockDim.x) (&tc[0][0])[i] = (seed * i) & 0x3FFF3FFFu;
    __syncthreads();
    unsigned T[2][32];
#pragma unroll
    for (int i = 0; i < 32; i++) { T[0][i] = (threadIdx.x * 17 + i * seed) & 0x3FFF3FFFu; T[1][i] = (threadIdx.x * 29 + i * seed) & 0x3FFF3FFFu; }
    unsigned acc[2] = {0, 0};
    int d[2][4] = {};
    const unsigned b0 = seed & 0x01010101u, b1 = (seed >> 1) & 0x01010101u;
    const long long t0 = clock64();
    const int g = (threadIdx.x >> 3) & 3;
    for (int it = 0; it < ITERS; it++) {
        const unsigned* row = tc[(2 * g + it) & 15];
        unsigned m[2][32];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint4 c = *reinterpret_cast<const uint4*>(row + 4 * k);
#pragma unroll
            for (int p = 0; p < 2; p++) {
                m[p][4 * k] = mx<MAXOP>(T[p][4 * k], c.x); m[p][4 * k + 1] = mx<MAXOP>(T[p][4 * k + 1], c.y);
                m[p][4 * k + 2] = mx<MAXOP>(T[p][4 * k + 2], c.z); m[p][4 * k + 3] = mx<MAXOP>(T[p][4 * k + 3], c.w);
            }
        }
#pragma unroll
        for (int p = 0; p < 2; p++) {
            if (ACC == 0) {
#pragma unroll
                for (int k = 0; k < 8; k++) acc[p] = fold2((m[p][4 * k] + m[p][4 * k + 1]) + (m[p][4 * k + 2] + m[p][4 * k + 3]), acc[p]);
            } else if (ACC == 1) {
#pragma unroll
                for (int k = 0; k < 32; k++) acc[p] = fold2(m[p][k], acc[p]);
            } else {
                unsigned s[12];
#pragma unroll
                for (int k = 0; k < 10; k++) s[k] = m[p][3 * k] + m[p][3 * k + 1] + m[p][3 * k + 2];
                s[10] = m[p][30] + m[p][31];
                s[11] = 0;
                if (ACC == 2) {
#pragma unroll
                    for (int k = 0; k < 11; k++) acc[p] = fold2(s[k], acc[p]);
                } else {
                    const unsigned a0[4] = {s[0], s[1], s[2], s[3]}, a1[4] = {s[4], s[5], s[6], s[7]}, a2[4] = {s[8], s[9], s[10], s[11]};
                    imma(d[p], a0, b0, b1); imma(d[p], a1, b0, b1); imma(d[p], a2, b0, b1);
                }
            }
        }
        if (ACC == 3) {                      // per candidate: fold the accumulator fragment (own columns) and reset
#pragma unroll
            for (int p = 0; p < 2; p++) { acc[p] += (d[p][0] + d[p][2]) + 256 * (d[p][1] + d[p][3]); d[p][0] = d[p][1] = d[p][2] = d[p][3] = 0; }
        }
    }
    unsigned s = acc[0] + acc[1];
#pragma unroll
    for (int i = 0; i < 32; i++) s += T[0][i] ^ T[1][i];
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 5 + (threadIdx.x >> 5)] = (unsigned)(t1 - t0);
}

int main(int argc, char** argv)
{
    const int only = argc > 1 ? atoi(argv[1]) : -1; int idx = 0;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, ctas = sms * 3;
    unsigned* out; cudaMalloc(&out, (size_t)ctas * 160 * 4);
    unsigned* cyc; cudaMallocManaged(&cyc, (size_t)ctas * 5 * 4);
#define RUN(A, M, name) if (only < 0 || only == idx++) { for (int r = 0; r < 3; r++) mix<A, M><<<ctas, 160>>>(out, cyc, 1234u); cudaDeviceSynchronize(); \
      double c = 0; for (int i = 0; i < ctas * 5; i++) c += cyc[i]; c /= ctas * 5; \
      printf("%-44s warp: %7.1f clk/iter  = %.1f SMSP cycles per 32 candidates (4 warps per SMSP)\n", name, c / ITERS, c / ITERS * 4 / 2); }
    RUN(0, 0, "VIMNMX, pairs + IDP (shipped)")
    RUN(1, 0, "VIMNMX, IDP per word")
    RUN(2, 0, "VIMNMX, IADD3 triples + IDP")
    RUN(3, 0, "VIMNMX, IADD3 triples + 3 IMMA")
    RUN(0, 1, "HMNMX2, pairs + IDP")
    RUN(2, 1, "HMNMX2, IADD3 triples + IDP")
    RUN(3, 1, "HMNMX2, IADD3 triples + 3 IMMA")
    return 0;
}
