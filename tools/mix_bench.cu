// tools/mix_bench.cu -- issue-rate check of the search-v3 inner loop instruction mix (sm_100a).
// Variants: 0 = max+add+idp from registers; 1 = same with the T(cur) words read from shared memory (LDS.128);
// 2 = only the VIMNMX.S16x2; 3 = VIMNMX + adds (no IDP).  Prints warp-instructions per clock per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
#include "../x266_b200/csrc/satd_packed.h"
using namespace x266::s3;
#define ITERS 2048

template <int V>
__global__ void __launch_bounds__(160, 3) mix(unsigned* out, unsigned* cyc, unsigned seed)
{
    __shared__ __align__(16) unsigned tc[16][36];
    __shared__ __align__(16) unsigned rep[16][8][32];        // variant 4/7: chunk k of a block replicated 8x side by side (one 128 B row)
    for (int i = threadIdx.x; i < 16 * 8 * 32; i += blockDim.x) (&rep[0][0][0])[i] = (seed * (i >> 2)) & 0x3FFF3FFFu;
    for (int i = threadIdx.x; i < 16 * 36; i += blockDim.x) (&tc[0][0])[i] = (seed * i) & 0x3FFF3FFFu;
    __syncthreads();
    unsigned TA[32], TB[32];
#pragma unroll
    for (int i = 0; i < 32; i++) { TA[i] = (threadIdx.x * 17 + i * seed) & 0x3FFF3FFFu; TB[i] = (threadIdx.x * 29 + i * seed) & 0x3FFF3FFFu; }
    unsigned TC[V == 8 ? 32 : 1];
    if (V == 8) {
#pragma unroll
        for (int i = 0; i < 32; i++) TC[i] = (threadIdx.x * 31 + i * seed) & 0x3FFF3FFFu;
    }
    unsigned accA = 0, accB = 0, accC = 0;
    const long long t0 = clock64();
    const int g = (threadIdx.x >> 3) & 3;
    uint4 cr[8];
#pragma unroll
    for (int k = 0; k < 8; k++) cr[k] = make_uint4(seed * k, seed + k, threadIdx.x + k, seed ^ k);
    for (int it = 0; it < ITERS; it++) {
        if (V == 9) {                       // rotate: 8 moves per iteration, every c differs from the last iteration's
            const uint4 t = cr[0];
#pragma unroll
            for (int k = 0; k < 7; k++) cr[k] = cr[k + 1];
            cr[7] = make_uint4(t.y, t.z, t.w, t.x + 1);
        }
        const unsigned* row = tc[(2 * g + it) & 15];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint4 c;
            if (V == 1 || V == 6 || V == 8) c = *reinterpret_cast<const uint4*>(row + 4 * k);
            else if (V == 4 || V == 7) c = *reinterpret_cast<const uint4*>(&rep[(2 * g + it) & 15][k][4 * (threadIdx.x & 7)]);
            else if (V == 9) c = cr[k];
            else if (V == 5) c = *reinterpret_cast<const uint4*>(&tc[it & 15][4 * k]);
            else c = make_uint4(accA + k, accB ^ k, seed + k, seed * k);
            if (V == 6 || V == 7) { accA += c.x ^ c.y; accB += c.z ^ c.w; }
            else if (V == 2) {
                TA[4 * k] = vmax2(TA[4 * k], c.x); TA[4 * k + 1] = vmax2(TA[4 * k + 1], c.y); TA[4 * k + 2] = vmax2(TA[4 * k + 2], c.z); TA[4 * k + 3] = vmax2(TA[4 * k + 3], c.w);
                TB[4 * k] = vmax2(TB[4 * k], c.x); TB[4 * k + 1] = vmax2(TB[4 * k + 1], c.y); TB[4 * k + 2] = vmax2(TB[4 * k + 2], c.z); TB[4 * k + 3] = vmax2(TB[4 * k + 3], c.w);
            } else {
                const unsigned ma = (vmax2(TA[4 * k], c.x) + vmax2(TA[4 * k + 1], c.y)) + (vmax2(TA[4 * k + 2], c.z) + vmax2(TA[4 * k + 3], c.w));
                const unsigned mb = (vmax2(TB[4 * k], c.x) + vmax2(TB[4 * k + 1], c.y)) + (vmax2(TB[4 * k + 2], c.z) + vmax2(TB[4 * k + 3], c.w));
                if (V == 8) {
                    const unsigned mc = (vmax2(TC[4 * k], c.x) + vmax2(TC[4 * k + 1], c.y)) + (vmax2(TC[4 * k + 2], c.z) + vmax2(TC[4 * k + 3], c.w));
                    accC = fold2(mc, accC);
                }
                if (V == 3) { accA += ma; accB += mb; }
                else { accA = fold2(ma, accA); accB = fold2(mb, accB); }
            }
        }
    }
    unsigned s = accA + accB + accC;
    if (V == 8) {
#pragma unroll
        for (int i = 0; i < 32; i++) s += TC[i];
    }
#pragma unroll
    for (int i = 0; i < 32; i++) s += TA[i] ^ TB[i];
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 5 + (threadIdx.x >> 5)] = (unsigned)(t1 - t0);
}

template <typename F> static float timeit(F f)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, ctas = sms * 3;
    unsigned* out; cudaMalloc(&out, (size_t)ctas * 160 * 4);
    unsigned* cyc; cudaMallocManaged(&cyc, (size_t)ctas * 5 * 4);
    const char* names[] = {"max+add+idp (regs)", "max+add+idp (LDS.128 bcast/8)", "max only", "max+add", "max+add+idp (LDS.128 replicated)",
                           "max+add+idp (LDS.128 warp-uniform)", "LDS.128 bcast/8 only", "LDS.128 replicated only", "3 positions: max+add+idp (LDS.128 bcast/8)", "max+add+idp (rotating regs, no LDS)"};
    const double instr[] = {64 + 48 + 16, 64 + 48 + 16 + 8, 64, 64 + 48 + 16, 64 + 48 + 16 + 8, 64 + 48 + 16 + 8, 8 + 32, 8 + 32, 96 + 72 + 24 + 8, 64 + 48 + 16};   // warp instructions per iteration (adds: 3 per group, x2)
#define RUN(V) { float ms = timeit([&] { mix<V><<<ctas, 160>>>(out, cyc, 1234u); }); cudaDeviceSynchronize(); \
      double c = 0; for (int i = 0; i < ctas * 5; i++) c += cyc[i]; c /= ctas * 5;   /* mean cycles a warp needed, 15 warps per SM resident */ \
      printf("%-36s %7.3f ms  warp: %.1f clk/iter -> %.3f warp-instr/clk/SMSP (nominal count), %.1f SMSP cycles per 2x32 candidates\n", names[V], ms, \
             c / ITERS, instr[V] * 3.75 / (c / ITERS), (c / ITERS) / 3.75); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
    return 0;
}
