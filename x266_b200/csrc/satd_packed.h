// satd_packed.h -- packed 16-bit, biased 2-D Hadamard of 8-bit pixel blocks and the max-sum cost form used by
// the full-search kernel v3 (satd_search3.cu).  Host+device: the same inline functions are compiled by g++ in
// tests/c/satd_packed_model.cpp, so the arithmetic is checked on the CPU against satd8x8 of the difference without a GPU.
//
// What is computed (src_tb/satd.c:31-118 applied to diff = cur - ref(mv)):  the Hadamard transform is linear, so
// T(diff) = T(cur) - T(ref) and the cost is (sum_k |Tcur[k] - Tref[k]| + 2) >> 2 -- exact because a transform of
// 8-bit pixels cannot reach the int16 wrap of satd.c:35 (|coef| <= 16320).
//
// Packing: two coefficients per 32-bit word, every half kept NON-NEGATIVE by a constant bias, so that a plain
// 32-bit add / subtract is two independent 16-bit add / subtracts (no carry or borrow crosses bit 16):
//   s = a + b            bias doubles
//   d = a - b + K        bias becomes K (the input biases are equal and cancel); K >= max |a - b|
// Vertical pass on pixels (two columns per word), K = 256, 512, 1024  -> output row 0 has bias 0, rows 1..7 bias 1024.
// Horizontal pass (two rows per word),            K = 2048, 4096, 8192 -> DC has bias 0, every AC coefficient 8192.
// AC coefficients of an 8-bit block lie in [-8160, 8160], DC in [0, 16320], so every biased value is in [0, 16352]
// and the sum of four of them still fits 16 bits.  Both operands of the cost go through the same code, so biases
// cancel in |a - b|.
//
// Cost without a subtraction per coefficient:  |a - b| = 2 max(a,b) - a - b, and the 64 coefficients of a Hadamard
// transform sum to 64 * x[0][0] (every basis row but the first sums to zero), so
//   sum_k |a_k - b_k| = 2 * sum_k max(a'_k, b'_k) - 64 * (ref[0][0] + cur[0][0]) - 2 * 63 * 8192.
// One VIMNMX.S16x2 handles two coefficients; four words are added with plain adds and folded into a 32-bit
// accumulator by one IDP.2A (dot product of the two halves with (1, 1)).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define S3_HD __host__ __device__ __forceinline__
#else
#define S3_HD inline
#endif

namespace x266 {
namespace s3 {

constexpr uint32_t PAIR = 0x00010001u;
constexpr int BIAS_SUM = 63 * 8192;              // sum of the 64 coefficient biases (DC carries none)

S3_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
#endif
}

S3_HD uint32_t vmax2(uint32_t a, uint32_t b)      // per-half signed max (values are < 2^15, so signed == unsigned)
{
#if defined(__CUDA_ARCH__)
    return __vmaxs2(a, b);
#else
    const uint32_t al = a & 0xFFFF, bl = b & 0xFFFF, ah = a >> 16, bh = b >> 16;
    return (al > bl ? al : bl) | ((ah > bh ? ah : bh) << 16);
#endif
}

S3_HD uint32_t fold2(uint32_t w, uint32_t acc)    // acc + low half + high half (both unsigned)
{
#if defined(__CUDA_ARCH__)
    return __dp2a_lo(w, 0x0101u, acc);
#else
    return acc + (w & 0xFFFF) + (w >> 16);
#endif
}

// in-place 8-point Hadamard on packed words, partner distances 4, 2, 1; element stride STRIDE words
template <uint32_t K1, uint32_t K2, uint32_t K3, int STRIDE>
S3_HD void had8p(uint32_t* v)
{
#pragma unroll
    for (int dist = 4; dist >= 1; dist >>= 1) {
        const uint32_t K = (dist == 4 ? K1 : dist == 2 ? K2 : K3) * PAIR;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (!(i & dist)) {
                const uint32_t a = v[i * STRIDE], b = v[(i + dist) * STRIDE];
                v[i * STRIDE] = a + b;
                v[(i + dist) * STRIDE] = a - b + K;
            }
        }
    }
}

// Vertical pass of four adjacent columns: px[i] = the 4 pixels (bytes, column x..x+3) of row i.
// out[c][j] = word of column x+c holding output rows (2j, 2j+1) in its (low, high) half.
S3_HD void vertical4(const uint32_t px[8], uint32_t out[4][4])
{
    uint32_t a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = prmt(px[i], 0u, 0x4140);          // (col x, col x+1)
        b[i] = prmt(px[i], 0u, 0x4342);          // (col x+2, col x+3)
    }
    had8p<256, 512, 1024, 1>(a);
    had8p<256, 512, 1024, 1>(b);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        out[0][j] = prmt(a[2 * j], a[2 * j + 1], 0x5410);
        out[1][j] = prmt(a[2 * j], a[2 * j + 1], 0x7632);
        out[2][j] = prmt(b[2 * j], b[2 * j + 1], 0x5410);
        out[3][j] = prmt(b[2 * j], b[2 * j + 1], 0x7632);
    }
}

// Horizontal pass at one position: T[8*j + c] for the four row-pair words j of 8 adjacent columns, in place
// (T[8*j + c] holds the vertical output of column c on entry).
S3_HD void horizontal8(uint32_t T[32])
{
#pragma unroll
    for (int j = 0; j < 4; j++) had8p<2048, 4096, 8192, 1>(&T[8 * j]);
}

// acc + sum of max(T[j], c_j) over four packed words.  FORM 0: two levels of packed adds, one IDP.2A; FORM 1: one IDP.2A
// per word (no packed adds: the integer ALU pipe only sees the VIMNMX); FORM 2: one packed add, two IDP.2A.
template <int FORM>
S3_HD uint32_t maxsum4(const uint32_t* T, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t acc)
{
    const uint32_t m0 = vmax2(T[0], c0), m1 = vmax2(T[1], c1), m2 = vmax2(T[2], c2), m3 = vmax2(T[3], c3);
    if (FORM == 0) return fold2((m0 + m1) + (m2 + m3), acc);
    if (FORM == 1) return fold2(m3, fold2(m2, fold2(m1, fold2(m0, acc))));
    return fold2(m2 + m3, fold2(m0 + m1, acc));
}

// sum_k max(T[k], C[k]) over the 64 packed coefficients
S3_HD uint32_t maxsum(const uint32_t T[32], const uint32_t C[32])
{
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t s = (vmax2(T[4 * k], C[4 * k]) + vmax2(T[4 * k + 1], C[4 * k + 1])) +
                           (vmax2(T[4 * k + 2], C[4 * k + 2]) + vmax2(T[4 * k + 3], C[4 * k + 3]));
        acc = fold2(s, acc);
    }
    return acc;
}

// (sum |Tcur - Tref| + 2) >> 2 from the max-sum and the two top-left pixels (satd.c:113)
S3_HD uint32_t cost_from_maxsum(uint32_t acc, uint32_t refPix, uint32_t curPix)
{
    return (2u * acc - 64u * (refPix + curPix) - 2u * (uint32_t)BIAS_SUM + 2u) >> 2;
}

} // namespace s3
} // namespace x266
