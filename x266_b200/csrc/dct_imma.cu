// dct_imma.cu -- 32x32 forward transform on the int8 tensor cores, one transform block per warp,
// fully register resident between the two passes (no transpose through memory at all).
//
// Reference behaviour: src_tb/dct32.c:197-198 (two partialButterfly32 passes, shifts s1/s2).  The
// dense formulation dst[k][j] = round(sum_n g_t32[k][n]*src[j][n]) is bit-identical to the butterfly
// because all arithmetic is exact 32-bit integer (|sum| <= 2048*32768 = 2^26).
//
// Byte-plane split: an int16 sample x = hi*256 + lo with lo in [0,255] (u8) and hi in [-128,127] (s8),
// the matrix entries fit s8 (|g| <= 90), so
//     sum_n g*x = sum_n g*lo  +  256 * sum_n g*hi
// is two m16n8k32 integer MMAs (s8 x u8 and s8 x s8) with s32 accumulators -- exact.
//
// Data flow per block (warp = 32 lanes, g = lane>>2, q = lane&3):
//   pass 1:  D1[mu][j]  = sum_kappa  A1[mu][kappa] * B1[kappa][j]
//            A1[mu][kappa] = g_t32[sigma(mu)][pi(kappa)]          (constant fragments, registers)
//            B1[kappa][j]  = src[j][pi(kappa)]                    (col-major B == row-major src)
//            -> thread holds coef[sigma(mu)][j] for mu in {16m+8h+g}, j in {8t+2q+e}
//   pass 2:  D2[k2][nu]  = sum_kappa2 A2[k2][kappa2] * B2[kappa2][nu],  nu = mu (pass-1 row position)
//            B2[kappa2][nu] = coef[sigma(nu)][pi2(kappa2)]        (taken straight from the pass-1
//                                                                  accumulator registers)
//            A2[k2][kappa2] = g_t32[k2][pi2(kappa2)]
//            -> thread holds dct[k2][sigma(nu)], nu in {8t2+2q+e}: sigma is chosen so that these are
//               the 8 contiguous columns 8q..8q+7 -> one 128-bit store per output row.
//   pi, pi2, sigma are pure index permutations; the contraction order is free (exact integers) and
//   the row order of pass 1 is free, so they cost nothing.
//
// Staging: each warp owns a ring of STAGES 2 KiB shared-memory slots filled by the TMA engine with
// 1-D bulk async copies (cp.async.bulk ... mbarrier::complete_tx), so HBM latency is covered by
// WARPS*STAGES*2 KiB in flight per CTA independent of register pressure.  Loads from the slot are
// linear 128-bit (lane*16 + t*512): conflict free.  Stores are 512 contiguous bytes per instruction.
#include "common.cuh"
#include "kernels.h"
#include "dct_frag.cuh"

namespace x266 {

// DIRECT = true replaces the shared-memory ring by register double-buffering with 128-bit global loads
// (kept as a measured alternative; see DESIGN.md for the sweep).
template <int IMMA_WARPS, int IMMA_STAGES, int MIN_CTAS, bool DIRECT>
__global__ void __launch_bounds__(IMMA_WARPS * 32, MIN_CTAS)
dct32_imma_kernel(const int16_t* __restrict__ src, int16_t* __restrict__ dst, size_t nBlocks, int shift1, int shift2)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    const uint32_t ring = smem_u32(smem) + warp * (IMMA_STAGES * 2048);
    const uint32_t bars = smem_u32(smem) + IMMA_WARPS * IMMA_STAGES * 2048 + warp * (IMMA_STAGES * 8);

    // ---- constant A fragments (m16n8k32 .row layout: a0 (g, 4q+i) a1 (g+8, 4q+i) a2 (g, 16+4q+i) a3 (g+8, 16+4q+i))
    uint32_t A1[2][4], A2[2][4];
#pragma unroll
    for (int m = 0; m < 2; m++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int row = 16 * m + g + 8 * (r & 1);
            const int kb = 16 * (r >> 1) + 4 * q;
            const int k1 = perm_sigma(row);
            A1[m][r] = pack4(c_g8.v[k1][perm_pi(kb + 0)], c_g8.v[k1][perm_pi(kb + 1)],
                             c_g8.v[k1][perm_pi(kb + 2)], c_g8.v[k1][perm_pi(kb + 3)]);
            A2[m][r] = pack4(c_g8.v[row][perm_pi2(kb + 0)], c_g8.v[row][perm_pi2(kb + 1)],
                             c_g8.v[row][perm_pi2(kb + 2)], c_g8.v[row][perm_pi2(kb + 3)]);
        }
    }

    const size_t first = (size_t)blockIdx.x * IMMA_WARPS + warp;
    const size_t stride = (size_t)gridDim.x * IMMA_WARPS;
    uint64_t policy = 0;
    uint4 nxt[DIRECT ? IMMA_STAGES : 1][4] = {};      // DIRECT: blocks b+stride .. b+STAGES*stride in flight

    if constexpr (!DIRECT) {
        policy = policy_evict_first();
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < IMMA_STAGES; s++) mbar_init(bars + 8 * s, 1);
            fence_mbar_init();
            fence_proxy_async();
#pragma unroll
            for (int s = 0; s < IMMA_STAGES; s++) {
                const size_t b = first + (size_t)s * stride;
                if (b < nBlocks) {
                    mbar_arrive_expect_tx(bars + 8 * s, 2048);
                    bulk_g2s(ring + s * 2048, src + b * 1024, 2048, bars + 8 * s, policy);
                }
            }
        }
        __syncwarp();
    } else {
#pragma unroll
        for (int s = 0; s < IMMA_STAGES; s++) {
            const size_t b = first + (size_t)s * stride;
            if (b < nBlocks) {
#pragma unroll
                for (int t = 0; t < 4; t++) nxt[s][t] = ld_global_stream(src + b * 1024 + t * 256 + lane * 8);
            }
        }
    }

    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);
    const int cAdd1[4] = { add1, add1, add1, add1 };
    const int cAdd2[4] = { add2, add2, add2, add2 };
    const int cZero[4] = { 0, 0, 0, 0 };

    int stage = 0;
    uint32_t parity = 0;
    for (size_t b = first; b < nBlocks; b += stride) {
        // ---- B1 fragments: 16-byte chunk q of row j = 8t+g sits at byte t*512 + lane*16 of the block
        uint32_t BL[4][2], BH[4][2];
        if constexpr (!DIRECT) {
            mbar_wait(bars + 8 * stage, parity);
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const uint4 w = ld_shared_v4(ring + stage * 2048 + t * 512 + lane * 16);
                BL[t][0] = prmt(w.x, w.y, 0x6420); BH[t][0] = prmt(w.x, w.y, 0x7531);
                BL[t][1] = prmt(w.z, w.w, 0x6420); BH[t][1] = prmt(w.z, w.w, 0x7531);
            }
            __syncwarp();
            if (lane == 0) {
                const size_t nb = b + (size_t)IMMA_STAGES * stride;
                if (nb < nBlocks) {
                    fence_proxy_async();    // order the generic-proxy reads above before the async-proxy refill
                    mbar_arrive_expect_tx(bars + 8 * stage, 2048);
                    bulk_g2s(ring + stage * 2048, src + nb * 1024, 2048, bars + 8 * stage, policy);
                }
            }
        } else {
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const uint4 w = nxt[0][t];
                BL[t][0] = prmt(w.x, w.y, 0x6420); BH[t][0] = prmt(w.x, w.y, 0x7531);
                BL[t][1] = prmt(w.z, w.w, 0x6420); BH[t][1] = prmt(w.z, w.w, 0x7531);
            }
#pragma unroll
            for (int s = 0; s + 1 < IMMA_STAGES; s++)
#pragma unroll
                for (int t = 0; t < 4; t++) nxt[s][t] = nxt[s + 1][t];
            const size_t nb = b + (size_t)IMMA_STAGES * stride;
            if (nb < nBlocks) {
#pragma unroll
                for (int t = 0; t < 4; t++) nxt[IMMA_STAGES - 1][t] = ld_global_stream(src + nb * 1024 + t * 256 + lane * 8);
            }
        }

        // ---- pass 1 ---------------------------------------------------------------------------
        uint32_t B2L[4][2], B2H[4][2];
#pragma unroll
        for (int m = 0; m < 2; m++) {
            int r[4][4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                int dl[4], dh[4];
                mma_s8u8(dl, A1[m], BL[t][0], BL[t][1], cAdd1);
                mma_s8s8(dh, A1[m], BH[t][0], BH[t][1], cZero);
#pragma unroll
                for (int c = 0; c < 4; c++) r[t][c] = (dl[c] + dh[c] * 256) >> shift1;
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t p0 = prmt(r[0][2 * h], r[0][2 * h + 1], 0x5140);
                const uint32_t p1 = prmt(r[1][2 * h], r[1][2 * h + 1], 0x5140);
                const uint32_t p2 = prmt(r[2][2 * h], r[2][2 * h + 1], 0x5140);
                const uint32_t p3 = prmt(r[3][2 * h], r[3][2 * h + 1], 0x5140);
                B2L[2 * m + h][0] = prmt(p0, p1, 0x5410); B2H[2 * m + h][0] = prmt(p0, p1, 0x7632);
                B2L[2 * m + h][1] = prmt(p2, p3, 0x5410); B2H[2 * m + h][1] = prmt(p2, p3, 0x7632);
            }
        }

        // ---- pass 2 + store -------------------------------------------------------------------
        int16_t* d = dst + b * 1024;
#pragma unroll
        for (int m2 = 0; m2 < 2; m2++) {
            int r[4][4];
#pragma unroll
            for (int t2 = 0; t2 < 4; t2++) {
                int dl[4], dh[4];
                mma_s8u8(dl, A2[m2], B2L[t2][0], B2L[t2][1], cAdd2);
                mma_s8s8(dh, A2[m2], B2H[t2][0], B2H[t2][1], cZero);
#pragma unroll
                for (int c = 0; c < 4; c++) r[t2][c] = (dl[c] + dh[c] * 256) >> shift2;
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint4 o;
                o.x = prmt(r[0][2 * h], r[0][2 * h + 1], 0x5410);
                o.y = prmt(r[1][2 * h], r[1][2 * h + 1], 0x5410);
                o.z = prmt(r[2][2 * h], r[2][2 * h + 1], 0x5410);
                o.w = prmt(r[3][2 * h], r[3][2 * h + 1], 0x5410);
                st_global_stream(d + (16 * m2 + 8 * h + g) * 32 + q * 8, o);
            }
        }

        if constexpr (!DIRECT) {
            if (++stage == IMMA_STAGES) { stage = 0; parity ^= 1; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 16x16 forward transform on the tensor cores: same choreography as the 32x32 kernel with the k16 MMA
// (A = G16 is exactly one m16 tile; G16[k][n] = g_t32[2k][n], src_tb/dct32.c:138-143 / mkDct32.bsv).
// One warp transforms 4 blocks (2 KiB) per iteration; lane (g,q) loads the 8-byte chunk q of rows g and
// g+8 of each block (bytes t*256 + lane*8: linear, coalesced) and ends with 8 contiguous output bytes per
// row (sigma16 below), i.e. 256 contiguous bytes per warp store.
// ------------------------------------------------------------------------------------------------
constexpr int D16_WARPS = 8;

__device__ __forceinline__ int perm_sigma16(int mu) { return 4 * ((mu >> 1) & 3) + 2 * (mu >> 3) + (mu & 1); }
// kappa2 = 4q + 2t + e  ->  j = 8t + 2q + e
__device__ __forceinline__ int perm_pi2_16(int k2) { return 8 * ((k2 >> 1) & 1) + 2 * (k2 >> 2) + (k2 & 1); }

__global__ void __launch_bounds__(D16_WARPS * 32, 2)
dct16_imma_kernel(const int16_t* __restrict__ src, int16_t* __restrict__ dst, size_t nBlocks, int shift1, int shift2)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    uint32_t A1[2], A2[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int row = g + 8 * r;
        const int k1 = perm_sigma16(row);
        A1[r] = pack4(c_g8.v[2 * k1][4 * q + 0], c_g8.v[2 * k1][4 * q + 1], c_g8.v[2 * k1][4 * q + 2], c_g8.v[2 * k1][4 * q + 3]);
        A2[r] = pack4(c_g8.v[2 * row][perm_pi2_16(4 * q + 0)], c_g8.v[2 * row][perm_pi2_16(4 * q + 1)],
                      c_g8.v[2 * row][perm_pi2_16(4 * q + 2)], c_g8.v[2 * row][perm_pi2_16(4 * q + 3)]);
    }
    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);
    const int cAdd1[4] = { add1, add1, add1, add1 };
    const int cAdd2[4] = { add2, add2, add2, add2 };
    const int cZero[4] = { 0, 0, 0, 0 };

    const size_t nUnits = (nBlocks + 3) / 4;
    const size_t first = (size_t)blockIdx.x * D16_WARPS + warp;
    const size_t stride = (size_t)gridDim.x * D16_WARPS;

    auto load_unit = [&](size_t u, uint2 (&w)[4][2]) {
#pragma unroll
        for (int bb = 0; bb < 4; bb++) {
            size_t b = u * 4 + bb;
            b = b < nBlocks ? b : nBlocks - 1;                 // ragged tail: re-read the last block, never stored
#pragma unroll
            for (int t = 0; t < 2; t++) w[bb][t] = ld_global_stream_v2(src + b * 256 + t * 128 + lane * 4);
        }
    };

    uint2 nxt[4][2] = {};
    if (first < nUnits) load_unit(first, nxt);

    for (size_t u = first; u < nUnits; u += stride) {
        uint32_t BL[4][2], BH[4][2];
#pragma unroll
        for (int bb = 0; bb < 4; bb++)
#pragma unroll
            for (int t = 0; t < 2; t++) {
                BL[bb][t] = prmt(nxt[bb][t].x, nxt[bb][t].y, 0x6420);
                BH[bb][t] = prmt(nxt[bb][t].x, nxt[bb][t].y, 0x7531);
            }
        if (u + stride < nUnits) load_unit(u + stride, nxt);

#pragma unroll
        for (int bb = 0; bb < 4; bb++) {
            int r[2][4];
#pragma unroll
            for (int t = 0; t < 2; t++) {
                int dl[4], dh[4];
                mma16_s8u8(dl, A1, BL[bb][t], cAdd1);
                mma16_s8s8(dh, A1, BH[bb][t], cZero);
#pragma unroll
                for (int c = 0; c < 4; c++) r[t][c] = (dl[c] + dh[c] * 256) >> shift1;
            }
            uint32_t B2L[2], B2H[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t p0 = prmt(r[0][2 * h], r[0][2 * h + 1], 0x5140);
                const uint32_t p1 = prmt(r[1][2 * h], r[1][2 * h + 1], 0x5140);
                B2L[h] = prmt(p0, p1, 0x5410);
                B2H[h] = prmt(p0, p1, 0x7632);
            }
            int r2[2][4];
#pragma unroll
            for (int t2 = 0; t2 < 2; t2++) {
                int dl[4], dh[4];
                mma16_s8u8(dl, A2, B2L[t2], cAdd2);
                mma16_s8s8(dh, A2, B2H[t2], cZero);
#pragma unroll
                for (int c = 0; c < 4; c++) r2[t2][c] = (dl[c] + dh[c] * 256) >> shift2;
            }
            const size_t b = u * 4 + bb;
            if (b < nBlocks) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    uint2 o;
                    o.x = prmt(r2[0][2 * h], r2[0][2 * h + 1], 0x5410);
                    o.y = prmt(r2[1][2 * h], r2[1][2 * h + 1], 0x5410);
                    st_global_stream_v2(dst + b * 256 + (g + 8 * h) * 16 + q * 4, o);
                }
            }
        }
    }
}

cudaError_t launch_dct16_imma(const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2, cudaStream_t st)
{
    if (nBlocks == 0) return cudaSuccess;
    const size_t want = ((nBlocks + 3) / 4 + D16_WARPS - 1) / D16_WARPS;
    const size_t cap = (size_t)sm_count() * 2;
    dct16_imma_kernel<<<(int)(want < cap ? want : cap), D16_WARPS * 32, 0, st>>>(src, dst, nBlocks, s1, s2);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// 8x8 forward transform on the tensor cores: two blocks (alpha, beta) per k16 MMA through the block-
// diagonal operand diag(G8, G8) (G8[k][n] = g_t32[4k][n], src_tb/dct32.c:132-136).  Rows 0-7 / K 0-7 of the
// MMA belong to alpha, rows 8-15 / K 8-15 to beta; the 8 columns are the 8 rows j of both blocks.
// Between the passes each lane owns 2 of the 4 j-values it needs for its K slice, the other 2 sit in lane
// q^2 of the same quad: one SHFL per block pair.  One warp transforms 16 blocks (2 KiB) per iteration.
// ------------------------------------------------------------------------------------------------
constexpr int D8_WARPS = 8;

// (Measured and not kept: the rounding shifts as the high word of a multiply by 2^(32-shift) -- IMAD.HI on the idle FMA pipe instead of SHF
// on the busy integer-ALU pipe -- 0.996 -> 0.93 of the roofline: IMAD.HI is a multi-pass instruction.)
__global__ void __launch_bounds__(D8_WARPS * 32, 2)
dct8_imma_kernel(const int16_t* __restrict__ src, int16_t* __restrict__ dst, size_t nBlocks, int shift1, int shift2)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    const bool lowHalf = q < 2;                // this lane's K slice belongs to block alpha (else beta)
    // pass 1: K position 4q+i <-> sample n = 4(q&1)+i of block (q>>1); pass 2: byte i of the K slice <-> j = pi2(i)
    //   pi2(i) = 2(q&1) + i (i<2),  2(q&1) + 4 + (i-2) (i>=2)
    uint32_t A1[2], A2[2];
    {
        const int n0 = 4 * (q & 1);
        const uint32_t a1 = pack4(c_g8.v[4 * g][n0 + 0], c_g8.v[4 * g][n0 + 1], c_g8.v[4 * g][n0 + 2], c_g8.v[4 * g][n0 + 3]);
        const int j0 = 2 * (q & 1);
        const uint32_t a2 = pack4(c_g8.v[4 * g][j0 + 0], c_g8.v[4 * g][j0 + 1], c_g8.v[4 * g][j0 + 4], c_g8.v[4 * g][j0 + 5]);
        A1[0] = lowHalf ? a1 : 0u; A1[1] = lowHalf ? 0u : a1;      // a0: row g (alpha rows), a1: row g+8 (beta rows)
        A2[0] = lowHalf ? a2 : 0u; A2[1] = lowHalf ? 0u : a2;
    }
    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);
    const int cAdd1[4] = { add1, add1, add1, add1 };
    const int cAdd2[4] = { add2, add2, add2, add2 };
    const int cZero[4] = { 0, 0, 0, 0 };

    const size_t nPairs = (nBlocks + 1) / 2;
    const size_t nUnits = (nPairs + 7) / 8;
    const size_t first = (size_t)blockIdx.x * D8_WARPS + warp;
    const size_t stride = (size_t)gridDim.x * D8_WARPS;
    // a unit = 16 consecutive blocks = 2 KiB: every unit but a ragged last one is addressed from ONE base pointer with immediate offsets
    // (the per-block clamps and 64-bit store predicates of the general form were a quarter of the kernel's integer-ALU instructions)
    const int offL = (q >> 1) * 64 + g * 8 + (q & 1) * 4, offS = g * 8 + q * 2;
    auto load_unit = [&](size_t u, uint2 (&w)[8]) {
        if ((u + 1) * 16 <= nBlocks) {
            const int16_t* p = src + u * 1024 + offL;
#pragma unroll
            for (int pp = 0; pp < 8; pp++) w[pp] = ld_global_stream_v2(p + pp * 128);
            return;
        }
#pragma unroll
        for (int pp = 0; pp < 8; pp++) {
            size_t blk = (u * 8 + pp) * 2 + (q >> 1);
            blk = blk < nBlocks ? blk : nBlocks - 1;               // ragged tail: clamp, never stored
            w[pp] = ld_global_stream_v2(src + blk * 64 + g * 8 + (q & 1) * 4);
        }
    };

    uint2 nxt[8] = {};
    if (first < nUnits) load_unit(first, nxt);

    for (size_t u = first; u < nUnits; u += stride) {
        uint32_t BL[8], BH[8];
#pragma unroll
        for (int pp = 0; pp < 8; pp++) {
            BL[pp] = prmt(nxt[pp].x, nxt[pp].y, 0x6420);
            BH[pp] = prmt(nxt[pp].x, nxt[pp].y, 0x7531);
        }
        if (u + stride < nUnits) load_unit(u + stride, nxt);
        const size_t left = nBlocks - u * 16;
        const int rem = left < 16 ? (int)left : 16;             // blocks of this unit that exist: one 32-bit compare per store
        int16_t* dU = dst + u * 1024 + offS;

#pragma unroll
        for (int pp = 0; pp < 8; pp++) {
            int dl[4], dh[4], r[4];
            mma16_s8u8(dl, A1, BL[pp], cAdd1);
            mma16_s8s8(dh, A1, BH[pp], cZero);
#pragma unroll
            for (int c = 0; c < 4; c++) r[c] = (dl[c] + dh[c] * 256) >> shift1;
            // r0,r1: coef_alpha[k=g][j=2q,2q+1];  r2,r3: coef_beta[k=g][j=2q,2q+1]
            const uint32_t mineA = prmt(r[0], r[1], 0x5410), mineB = prmt(r[2], r[3], 0x5410);
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, lowHalf ? mineB : mineA, 2);
            const uint32_t firstW = lowHalf ? mineA : recv;         // j = 2(q&1), 2(q&1)+1
            const uint32_t secondW = lowHalf ? recv : mineB;        // j = 2(q&1)+4, 2(q&1)+5
            const uint32_t B2L = prmt(firstW, secondW, 0x6420), B2H = prmt(firstW, secondW, 0x7531);
            mma16_s8u8(dl, A2, B2L, cAdd2);
            mma16_s8s8(dh, A2, B2H, cZero);
#pragma unroll
            for (int c = 0; c < 4; c++) r[c] = (dl[c] + dh[c] * 256) >> shift2;
            // r0,r1: dct_alpha[k2=g][k=2q,2q+1];  r2,r3: dct_beta[k2=g][k=2q,2q+1]
            if (2 * pp < rem) *reinterpret_cast<uint32_t*>(dU + pp * 128) = prmt(r[0], r[1], 0x5410);
            if (2 * pp + 1 < rem) *reinterpret_cast<uint32_t*>(dU + pp * 128 + 64) = prmt(r[2], r[3], 0x5410);
        }
    }
}

static std::atomic<int> g_dct8Ctas{0};       // tuning/diagnostic: CTAs per SM of the persistent grid (0 = 2)
void set_dct8_ctas(int v) { g_dct8Ctas = v; }

cudaError_t launch_dct8_imma(const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2, cudaStream_t st)
{
    if (nBlocks == 0) return cudaSuccess;
    const size_t want = ((nBlocks + 15) / 16 + D8_WARPS - 1) / D8_WARPS;
    const size_t cap = (size_t)sm_count() * (g_dct8Ctas > 0 ? g_dct8Ctas.load() : 2);
    dct8_imma_kernel<<<(int)(want < cap ? want : cap), D8_WARPS * 32, 0, st>>>(src, dst, nBlocks, s1, s2);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// "Next" row N2/N1: residual formation + 32x32 transform straight from the encoder's tiled
// frame stores.  Frames are rasters of ref_block_t tiles (src/x266.cpp:56-63: 512 bytes = m_Y[16*16] |
// m_C[2*8*8] | m_I[128]; stride m_frames_strd = width/16 tiles, x266.cpp:503), written by xConvInputFmt
// (x266.cpp:415-453).  For every 32x32 luma block (2x2 tiles) of `cur` and `pred`:
//     coef = DCT32( cur - pred )            (dct32.c:197-198 on the 9-bit residual)
// Linearity makes the residual free: G*(c - p) = G*c + (-G)*p, both operands are raw u8 pixels, so pass 1
// is two s8 x u8 MMAs into one accumulator per tile -- no subtraction, no byte-plane split, no combine.
// Pass 2 and the store are those of dct32_imma_kernel.  The int16 residual never exists in memory.
// ------------------------------------------------------------------------------------------------
constexpr int FR_WARPS = 8;

// DEPTH = blocks whose pixels are in flight (registers) ahead of the one being transformed, MINB = CTAs per SM
template <int DEPTH, int MINB>
__global__ void __launch_bounds__(FR_WARPS * 32, MINB)
frame_resi_dct32_kernel(const uint8_t* __restrict__ cur, const uint8_t* __restrict__ pred, int tilesPerRow, int blocksPerRow,
                        size_t nBlocks, int16_t* __restrict__ dst, int shift1, int shift2)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    uint32_t A1[2][4], A1n[2][4], A2[2][4];
#pragma unroll
    for (int m = 0; m < 2; m++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int row = 16 * m + g + 8 * (r & 1);
            const int kb = 16 * (r >> 1) + 4 * q;
            const int k1 = perm_sigma(row);
            const int a = c_g8.v[k1][perm_pi(kb + 0)], b = c_g8.v[k1][perm_pi(kb + 1)];
            const int c = c_g8.v[k1][perm_pi(kb + 2)], d = c_g8.v[k1][perm_pi(kb + 3)];
            A1[m][r] = pack4(a, b, c, d);
            A1n[m][r] = pack4(-a, -b, -c, -d);
            A2[m][r] = pack4(c_g8.v[row][perm_pi2(kb + 0)], c_g8.v[row][perm_pi2(kb + 1)],
                             c_g8.v[row][perm_pi2(kb + 2)], c_g8.v[row][perm_pi2(kb + 3)]);
        }
    }
    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);
    const int cAdd1[4] = { add1, add1, add1, add1 };
    const int cAdd2[4] = { add2, add2, add2, add2 };
    const int cZero[4] = { 0, 0, 0, 0 };

    const size_t first = (size_t)blockIdx.x * FR_WARPS + warp;
    const size_t stride = (size_t)gridDim.x * FR_WARPS;

    // lane's 8 pixels of row j = 8t+g: columns 8q..8q+7 -> tile column q>>1, tile row j>>4 = t>>1, in-tile offset (j&15)*16 + 8*(q&1) =
    // (t&1)*128 + g*16 + 8*(q&1): ONE base address per block, the four rows at +0, +128, +rowStride, +rowStride+128.  The block's (bx, by)
    // is carried along the grid stride instead of divided out of the 64-bit block index every time.
    const unsigned bpr = (unsigned)blocksPerRow;
    const unsigned rowStride = (unsigned)tilesPerRow * 512u;
    const unsigned laneOff = (unsigned)(q >> 1) * 512u + (unsigned)g * 16u + 8u * (unsigned)(q & 1);
    const unsigned sby = (unsigned)(stride / bpr), sbx = (unsigned)(stride % bpr);
    unsigned pby = (unsigned)(first / bpr), pbx = (unsigned)(first % bpr);      // block the next prefetch loads
    auto load_next = [&](uint2 (&c)[4], uint2 (&p)[4]) {
        const size_t base = ((size_t)(2u * pby) * (unsigned)tilesPerRow + 2u * pbx) * 512u + laneOff;
        const uint8_t* cb = cur + base;
        const uint8_t* pb = pred + base;
        c[0] = ld_global_stream_v2(cb); c[1] = ld_global_stream_v2(cb + 128);
        c[2] = ld_global_stream_v2(cb + rowStride); c[3] = ld_global_stream_v2(cb + rowStride + 128);
        p[0] = ld_global_stream_v2(pb); p[1] = ld_global_stream_v2(pb + 128);
        p[2] = ld_global_stream_v2(pb + rowStride); p[3] = ld_global_stream_v2(pb + rowStride + 128);
        pbx += sbx; pby += sby;
        if (pbx >= bpr) { pbx -= bpr; pby++; }
    };

    uint2 nc[DEPTH][4] = {}, np[DEPTH][4] = {};
#pragma unroll
    for (int d = 0; d < DEPTH; d++)
        if (first + d * stride < nBlocks) load_next(nc[d], np[d]);

    for (size_t b = first; b < nBlocks; b += stride) {
        uint2 bc[4], bp[4];
#pragma unroll
        for (int t = 0; t < 4; t++) { bc[t] = nc[0][t]; bp[t] = np[0][t]; }
#pragma unroll
        for (int d = 0; d + 1 < DEPTH; d++)
#pragma unroll
            for (int t = 0; t < 4; t++) { nc[d][t] = nc[d + 1][t]; np[d][t] = np[d + 1][t]; }
        if (b + DEPTH * stride < nBlocks) load_next(nc[DEPTH - 1], np[DEPTH - 1]);

        uint32_t B2L[4][2], B2H[4][2];
#pragma unroll
        for (int m = 0; m < 2; m++) {
            int r[4][4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                int acc[4];
                mma_s8u8(acc, A1[m], bc[t].x, bc[t].y, cAdd1);
                mma_s8u8(acc, A1n[m], bp[t].x, bp[t].y, acc);
#pragma unroll
                for (int c = 0; c < 4; c++) r[t][c] = acc[c] >> shift1;
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t p0 = prmt(r[0][2 * h], r[0][2 * h + 1], 0x5140);
                const uint32_t p1 = prmt(r[1][2 * h], r[1][2 * h + 1], 0x5140);
                const uint32_t p2 = prmt(r[2][2 * h], r[2][2 * h + 1], 0x5140);
                const uint32_t p3 = prmt(r[3][2 * h], r[3][2 * h + 1], 0x5140);
                B2L[2 * m + h][0] = prmt(p0, p1, 0x5410); B2H[2 * m + h][0] = prmt(p0, p1, 0x7632);
                B2L[2 * m + h][1] = prmt(p2, p3, 0x5410); B2H[2 * m + h][1] = prmt(p2, p3, 0x7632);
            }
        }
        int16_t* d = dst + b * 1024;
#pragma unroll
        for (int m2 = 0; m2 < 2; m2++) {
            int r[4][4];
#pragma unroll
            for (int t2 = 0; t2 < 4; t2++) {
                int dl[4], dh[4];
                mma_s8u8(dl, A2[m2], B2L[t2][0], B2L[t2][1], cAdd2);
                mma_s8s8(dh, A2[m2], B2H[t2][0], B2H[t2][1], cZero);
#pragma unroll
                for (int c = 0; c < 4; c++) r[t2][c] = (dl[c] + dh[c] * 256) >> shift2;
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint4 o;
                o.x = prmt(r[0][2 * h], r[0][2 * h + 1], 0x5410);
                o.y = prmt(r[1][2 * h], r[1][2 * h + 1], 0x5410);
                o.z = prmt(r[2][2 * h], r[2][2 * h + 1], 0x5410);
                o.w = prmt(r[3][2 * h], r[3][2 * h + 1], 0x5410);
                st_global_stream(d + (16 * m2 + 8 * h + g) * 32 + q * 8, o);
            }
        }
    }
}

static std::atomic<int> g_frameResiCfg{0};   // tuning/diagnostic: 0 = two blocks in flight per warp (shipped), 1 = one (round 1)
void set_frame_resi_config(int v) { g_frameResiCfg = v; }

cudaError_t launch_frame_resi_dct32(const uint8_t* cur, const uint8_t* pred, int width, int height, int16_t* dst,
                                    int s1, int s2, cudaStream_t st)
{
    if (width <= 0 || height <= 0 || (width & 31) || (height & 31)) return cudaErrorInvalidValue;
    const size_t nBlocks = (size_t)(width / 32) * (height / 32);
    const size_t want = (nBlocks + FR_WARPS - 1) / FR_WARPS;
    auto go = [&](auto kern, int ctas) {
        const size_t cap = (size_t)sm_count() * ctas;
        kern<<<(int)(want < cap ? want : cap), FR_WARPS * 32, 0, st>>>(cur, pred, width / 16, width / 32, nBlocks, dst, s1, s2);
    };
    // measured on 16 stacked 8K frames (profiles/r02_frame_resi_sweep.log): depth 2 at 2 CTAs/SM 0.959 of the HBM roofline, depth 1 0.941-0.946,
    // depth 3 0.936 (126 -> 128 registers, nothing left for the MMA section), 3 or 4 CTAs/SM 0.69-0.85 (register spills)
    if (g_frameResiCfg.load() == 1) go(frame_resi_dct32_kernel<1, 2>, 2);
    else go(frame_resi_dct32_kernel<2, 2>, 2);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// "Next" row N3: inverse 32x32 transform with the same matrix, closing the reconstruction loop.
// Not in the reference C (PARITY UNPINNED); defined as the HEVC/VVC decoder does it and as the HM model
// partialButterflyInverse32 computes it: vertical pass first,
//     tmp[y][v] = clip16((sum_u g[u][y] * coef[u][v] + (1 << (s1-1))) >> s1)          (s1 = 7)
//     out[y][x] = clip16((sum_v tmp[y][v] * g[v][x] + (1 << (s2-1))) >> s2)          (s2 = 12 for 8-bit video)
// clip16 = saturation to int16 (the standard's Clip3), unlike the forward path's truncating store.
// Pass 1 contracts over the ROW index of the stored block, so the B fragments (4 consecutive K per byte
// register) need a 4x4 byte transpose: each lane loads 8-byte pieces of 8 rows and transposes them with 8 PRMT
// per 16 samples (that also performs the byte-plane split).  The pass-1 accumulators of an m16 tile are exactly
// one A fragment of pass 2 (rows y stay rows), so again nothing is transposed through memory.
// ------------------------------------------------------------------------------------------------
constexpr int IDCT_WARPS = 8;

__global__ void __launch_bounds__(IDCT_WARPS * 32, 2)
idct32_imma_kernel(const int16_t* __restrict__ src, int16_t* __restrict__ dst, size_t nBlocks, int shift1, int shift2)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    // pass 1: A1[mu = y][kappa = u] = g[u][y]
    uint32_t A1[2][4];
#pragma unroll
    for (int m = 0; m < 2; m++)
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int y = 16 * m + g + 8 * (r & 1);
            const int kb = 16 * (r >> 1) + 4 * q;
            A1[m][r] = pack4(c_g8.v[kb + 0][y], c_g8.v[kb + 1][y], c_g8.v[kb + 2][y], c_g8.v[kb + 3][y]);
        }
    // pass 2: B2[kappa2][nu] = g[v(kappa2)][x(nu)],  v = 8q + 4(i&1) + 2*hi + (i>>1),  x = sigma(8*tx + g)
    uint32_t B2[4][2];
#pragma unroll
    for (int tx = 0; tx < 4; tx++)
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            const int x = perm_sigma(8 * tx + g);
            int v[4];
#pragma unroll
            for (int i = 0; i < 4; i++) v[i] = c_g8.v[8 * q + 4 * (i & 1) + 2 * rr + (i >> 1)][x];
            B2[tx][rr] = pack4(v[0], v[1], v[2], v[3]);
        }
    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);
    const int cAdd1[4] = { add1, add1, add1, add1 };
    const int cAdd2[4] = { add2, add2, add2, add2 };
    const int cZero[4] = { 0, 0, 0, 0 };

    const size_t first = (size_t)blockIdx.x * IDCT_WARPS + warp;
    const size_t stride = (size_t)gridDim.x * IDCT_WARPS;

    // lane loads columns 4g..4g+3 (8 bytes) of rows 4q+i and 16+4q+i
    auto load_block = [&](size_t b, uint2 (&w)[2][4]) {
#pragma unroll
        for (int hh = 0; hh < 2; hh++)
#pragma unroll
            for (int i = 0; i < 4; i++) w[hh][i] = ld_global_stream_v2(src + b * 1024 + (16 * hh + 4 * q + i) * 32 + g * 4);
    };
    uint2 nxt[2][4] = {};
    if (first < nBlocks) load_block(first, nxt);

    for (size_t b = first; b < nBlocks; b += stride) {
        // 4x4 byte transposes -> B1 fragments (n-tile t <-> column v = 4g+t)
        uint32_t BL[4][2], BH[4][2];
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const uint32_t r0 = half ? nxt[hh][0].y : nxt[hh][0].x, r1 = half ? nxt[hh][1].y : nxt[hh][1].x;
                const uint32_t r2 = half ? nxt[hh][2].y : nxt[hh][2].x, r3 = half ? nxt[hh][3].y : nxt[hh][3].x;
                const uint32_t t0 = prmt(r0, r1, 0x5140), t1 = prmt(r2, r3, 0x5140);
                const uint32_t t2 = prmt(r0, r1, 0x7362), t3 = prmt(r2, r3, 0x7362);
                BL[2 * half][hh] = prmt(t0, t1, 0x5410);     BH[2 * half][hh] = prmt(t0, t1, 0x7632);
                BL[2 * half + 1][hh] = prmt(t2, t3, 0x5410); BH[2 * half + 1][hh] = prmt(t2, t3, 0x7632);
            }
        }
        if (b + stride < nBlocks) load_block(b + stride, nxt);

        int16_t* d = dst + b * 1024;
#pragma unroll
        for (int m = 0; m < 2; m++) {
            int r[4][4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                int dl[4], dh[4];
                mma_s8u8(dl, A1[m], BL[t][0], BL[t][1], cAdd1);
                mma_s8s8(dh, A1[m], BH[t][0], BH[t][1], cZero);
#pragma unroll
                for (int c = 0; c < 4; c++) r[t][c] = (dl[c] + dh[c] * 256) >> shift1;
            }
            // accumulators of this m16 tile -> pass-2 A fragment (rows y = 16m+g / +8; K slice of this lane)
            uint32_t AL[4], AH[4];
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int hi = 0; hi < 2; hi++) {
                    const uint32_t p0 = pack_sat16(r[2 * hi][2 * h], r[2 * hi][2 * h + 1]);          // clip16 + pack
                    const uint32_t p1 = pack_sat16(r[2 * hi + 1][2 * h], r[2 * hi + 1][2 * h + 1]);
                    AL[h + 2 * hi] = prmt(p0, p1, 0x6420);
                    AH[h + 2 * hi] = prmt(p0, p1, 0x7531);
                }
            int r2[4][4];
#pragma unroll
            for (int tx = 0; tx < 4; tx++) {
                int dl[4], dh[4];
                mma_u8s8(dl, AL, B2[tx][0], B2[tx][1], cAdd2);
                mma_s8s8(dh, AH, B2[tx][0], B2[tx][1], cZero);
#pragma unroll
                for (int c = 0; c < 4; c++) r2[tx][c] = (dl[c] + dh[c] * 256) >> shift2;
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint4 o;
                o.x = pack_sat16(r2[0][2 * h], r2[0][2 * h + 1]);
                o.y = pack_sat16(r2[1][2 * h], r2[1][2 * h + 1]);
                o.z = pack_sat16(r2[2][2 * h], r2[2][2 * h + 1]);
                o.w = pack_sat16(r2[3][2 * h], r2[3][2 * h + 1]);
                st_global_stream(d + (16 * m + 8 * h + g) * 32 + q * 8, o);
            }
        }
    }
}

cudaError_t launch_idct32_imma(const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2, cudaStream_t st)
{
    if (nBlocks == 0) return cudaSuccess;
    const size_t want = (nBlocks + IDCT_WARPS - 1) / IDCT_WARPS;
    const size_t cap = (size_t)sm_count() * 2;
    idct32_imma_kernel<<<(int)(want < cap ? want : cap), IDCT_WARPS * 32, 0, st>>>(src, dst, nBlocks, s1, s2);
    count_launch();
    return cudaGetLastError();
}

// ---- configuration table (index = tuning id).  Shipped default = 6 (8 warps, 2 CTAs/SM, register
// double-buffered 128-bit global loads): 95.7 % of the measured HBM roofline on B200 vs 86.7 % for the
// best TMA-ring instantiation (profiles/r01_tune_dct.log).
constexpr int IMMA_DEFAULT_CFG = 6;
static std::atomic<int> g_immaCfg{IMMA_DEFAULT_CFG};
void set_imma_config(int id) { g_immaCfg = id < 0 ? IMMA_DEFAULT_CFG : id; }

template <int W, int S, int B, bool D>
static cudaError_t launch_cfg(const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2, cudaStream_t st)
{
    constexpr int SMEM = D ? 0 : (W * S * 2048 + W * S * 8);
    static std::atomic<bool> attrSet[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (SMEM > 48 * 1024 && (dev < 0 || dev >= 64 || !attrSet[dev])) {
        cudaError_t e = cudaFuncSetAttribute(dct32_imma_kernel<W, S, B, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attrSet[dev] = true;
    }
    const size_t want = (nBlocks + W - 1) / W;
    const size_t cap = (size_t)sm_count() * B;
    dct32_imma_kernel<W, S, B, D><<<(int)(want < cap ? want : cap), W * 32, SMEM, st>>>(src, dst, nBlocks, s1, s2);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_dct32_imma(const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2, cudaStream_t st)
{
    if (nBlocks == 0) return cudaSuccess;
    // the instantiations the staging sweep left standing (profiles/r01_tune_dct.log; ids as in the sweep): per-warp TMA bulk-copy rings top out at
    // 0.875 of the copy roofline, register double-buffered 128-bit loads reach 0.99
    switch (g_immaCfg) {
    case 0: return launch_cfg<8, 4, 2, false>(src, dst, nBlocks, s1, s2, st);     // TMA ring, 8 warps x 4 stages
    case 3: return launch_cfg<8, 6, 2, false>(src, dst, nBlocks, s1, s2, st);     // TMA ring, 8 warps x 6 stages
    default:
    case 6: return launch_cfg<8, 1, 2, true>(src, dst, nBlocks, s1, s2, st);      // direct loads, one block ahead (shipped)
    case 7: return launch_cfg<8, 1, 3, true>(src, dst, nBlocks, s1, s2, st);      // direct loads, 3 CTAs per SM
    case 12: return launch_cfg<8, 2, 2, true>(src, dst, nBlocks, s1, s2, st);     // direct loads, two blocks ahead
    }
}

} // namespace x266
