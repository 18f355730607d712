// dct_frag.cuh -- the byte-plane IMMA transform's fragment helpers, shared by dct_imma.cu and encode.cu: the s8 matrix, the index
// permutations of the register-resident two-pass choreography (see the header comment of dct_imma.cu) and the packers.
#pragma once
#include "common.cuh"

namespace x266 {

struct G8 { int8_t v[32][32]; };
constexpr G8 make_g8()
{
    G8 g{};
    for (int k = 0; k < 32; k++)
        for (int n = 0; n < 32; n++) g.v[k][n] = (int8_t)g32(k, n);
    return g;
}
// global (not __constant__) memory: every lane gathers different bytes of the table for its MMA fragments, which the constant
// cache serialises 32 ways (measured: 20 us of prologue per launch); plain cached loads take about 2 us.
static __device__ const G8 c_g8 = make_g8();


__device__ __forceinline__ int perm_sigma(int mu)  { return 8 * ((mu >> 1) & 3) + 2 * (mu >> 3) + (mu & 1); }
__device__ __forceinline__ int perm_pi(int kappa)  { return 8 * ((kappa & 15) >> 2) + 4 * (kappa >> 4) + (kappa & 3); }
// kappa2 = 16*hi + 4*q + i  ->  j = 8*(2*hi + (i>>1)) + 2*q + (i&1)
__device__ __forceinline__ int perm_pi2(int kappa2)
{
    const int hi = kappa2 >> 4, q = (kappa2 >> 2) & 3, i = kappa2 & 3;
    return 8 * (2 * hi + (i >> 1)) + 2 * q + (i & 1);
}

__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d)
{
    return (uint32_t)(a & 0xFF) | ((uint32_t)(b & 0xFF) << 8) | ((uint32_t)(c & 0xFF) << 16) | ((uint32_t)(d & 0xFF) << 24);
}

// pack two s32 into s16x2 with saturation (the standard's Clip3 to int16) in one I2IP.S16.S32.SAT: lo = sat(a), hi = sat(b)
__device__ __forceinline__ uint32_t pack_sat16(int a, int b)
{
    uint32_t r;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(r) : "r"(b), "r"(a));
    return r;
}

} // namespace x266
