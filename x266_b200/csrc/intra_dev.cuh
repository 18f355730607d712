// intra_dev.cuh -- device pieces of the intra predictor shared by intra.cu (prediction, mode decision) and encode.cu (the fused
// decide -> predict -> transform -> reconstruct kernel): mode constants and the SWAR row generators.  Spec: src/mkIntra32-wip.bsv.
#pragma once
#include "common.cuh"

namespace x266 {

static __constant__ int c_intraAngle[35] = { 0, 0, 32, 26, 21, 17, 13, 9, 5, 2, 0, -2, -5, -9, -13, -17, -21, -26,
                                      -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32 };
// 8192/|angle| rounded, for angle = -2,-5,-9,-13,-17,-21,-26,-32
static __constant__ int c_intraInv[8] = { 4096, 1638, 910, 630, 482, 390, 315, 256 };
// the same per mode (0 where the angle is not negative)
static __constant__ int c_intraInvMode[35] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4096, 1638, 910, 630, 482, 390, 315,
                                        256, 315, 390, 482, 630, 910, 1638, 4096, 0, 0, 0, 0, 0, 0, 0, 0, 0 };

__device__ __forceinline__ uint32_t intra_row4(uint32_t a, uint32_t b, int f)
{
    // 4 pixels: ((32-f)*a + f*b + 16) >> 5 per byte, two 16-bit lanes per multiply-add
    const uint32_t ae = a & 0x00FF00FFu, ao = (a >> 8) & 0x00FF00FFu;
    const uint32_t be = b & 0x00FF00FFu, bo = (b >> 8) & 0x00FF00FFu;
    const uint32_t pe = ((ae * (uint32_t)(32 - f) + be * (uint32_t)f + 0x00100010u) >> 5) & 0x00FF00FFu;
    const uint32_t po = ((ao * (uint32_t)(32 - f) + bo * (uint32_t)f + 0x00100010u) >> 5) & 0x00FF00FFu;
    return pe | (po << 8);
}

// D(16x8,s32) += A(16x32,u8,row) * B(32x8,u8,col): the angular / planar predictors' weight x Hankel product
__device__ __forceinline__ void mma_u8u8_16832(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(0));
}

// pixel = byte 1 of each 32-bit sum: the four sums of accumulator rows r of two adjacent n8 tiles -> one word of four pixels
__device__ __forceinline__ uint32_t intra_word(const int (&a)[4], const int (&b)[4], int r)
{
    return __byte_perm(__byte_perm((uint32_t)a[2 * r], (uint32_t)a[2 * r + 1], 0x5151),
                       __byte_perm((uint32_t)b[2 * r], (uint32_t)b[2 * r + 1], 0x5151), 0x5410);
}

// The same word with the shift-and-merge on the FMA pipe: every sum is < 2^16 (pixel in byte 1, remainder in byte 0), so a * 2^16 + b holds
// the pixel of a in byte 3 and the pixel of b in byte 1 with no carry between them -- two multiply-adds (k16 = 65536 in a register the
// compiler cannot fold, so that they stay IMADs) and ONE permute per word instead of three permutes on the integer ALU pipe, which is the
// busiest pipe of the fractional modes (ncu: 52-76 %, profiles/r02_ncu_intra_frac.md).
__device__ __forceinline__ uint32_t intra_word_fma(const int (&a)[4], const int (&b)[4], int r, uint32_t k16)
{
    const uint32_t x = (uint32_t)a[2 * r] * k16 + (uint32_t)b[2 * r], y = (uint32_t)a[2 * r + 1] * k16 + (uint32_t)b[2 * r + 1];
    return __byte_perm(x, y, 0x5173);
}
__device__ __forceinline__ uint32_t intra_k16()
{
    uint32_t k;
    asm volatile("mov.b32 %0, 0x10000;" : "=r"(k));
    return k;
}

constexpr int INTRA_WARPS = 8;
constexpr int INTRA_STRIP = 112;            // 32 (negative part) + 66 + padding, multiple of 16

// Angular modes, SWAR: lane l generates, for it = 0,1, the 16 pixels (row 16*it + (l>>1), columns 16*(l&1)..+15) of
// the vertical-family prediction P_v (distance = row, position along the main reference = column): one (idx, f)
// pair per row, 5 aligned words of the reference strip re-aligned with funnel shifts, 2 pixels per multiply-add.
// Vertical modes store the four words directly (512 contiguous bytes per warp store).  Horizontal modes are
// P_v^T of the left reference: the rows go through a padded per-warp tile and are read back as columns.
// 16 pixels of one prediction row from 17 reference bytes starting at byte `sh/8` of p[0], 9 instructions per 4 pixels:
// the weights are pre-scaled by 8 so that ((32-f)a + f b + 16) >> 5 is the HIGH byte of each 16-bit lane and the final
// shift-and-mask folds into the byte permute that interleaves the even and odd pixels.  With the aligned bytes a0..a4,
// E0 = (a0,a2), O = (a1,a3), E1 = (a2,a4) (16-bit lanes):  even pixels = E0*w0 + O*w1,  odd pixels = O*w0 + E1*w1,
// and E1 is one permute of this word's and the next word's E0.
__device__ __forceinline__ void intra_row16(const uint32_t* __restrict__ p, int sh, int f, uint32_t (&out)[4])
{
    const uint32_t x0 = p[0], x1 = p[1], x2 = p[2], x3 = p[3], x4 = p[4];
    uint32_t E[5];
    const uint32_t A0 = __funnelshift_r(x0, x1, sh), A1 = __funnelshift_r(x1, x2, sh), A2 = __funnelshift_r(x2, x3, sh),
                   A3 = __funnelshift_r(x3, x4, sh), A4 = x4 >> sh;
    E[0] = A0 & 0x00FF00FFu; E[1] = A1 & 0x00FF00FFu; E[2] = A2 & 0x00FF00FFu; E[3] = A3 & 0x00FF00FFu; E[4] = A4 & 0x00FF00FFu;
    const uint32_t A[4] = { A0, A1, A2, A3 };
    const uint32_t w0 = (uint32_t)(32 - f) * 8u, w1 = (uint32_t)f * 8u;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t O = __byte_perm(A[j], 0u, 0x4341);
        const uint32_t E1 = __byte_perm(E[j], E[j + 1], 0x5432);
        const uint32_t pe = E[j] * w0 + (O * w1 + 0x00800080u);
        const uint32_t po = O * w0 + (E1 * w1 + 0x00800080u);
        out[j] = __byte_perm(pe, po, 0x7351);
    }
}

// the same for 8 pixels (9 reference bytes): used by the mode-decision kernel, whose A fragments are 8-pixel row pieces
__device__ __forceinline__ void intra_row8(const uint32_t* __restrict__ p, int sh, int f, uint32_t& out0, uint32_t& out1)
{
    const uint32_t x0 = p[0], x1 = p[1], x2 = p[2];
    const uint32_t A0 = __funnelshift_r(x0, x1, sh), A1 = __funnelshift_r(x1, x2, sh), A2 = x2 >> sh;
    const uint32_t E0 = A0 & 0x00FF00FFu, E1 = A1 & 0x00FF00FFu, E2 = A2 & 0x00FF00FFu;
    const uint32_t w0 = (uint32_t)(32 - f) * 8u, w1 = (uint32_t)f * 8u;
    const uint32_t O0 = __byte_perm(A0, 0u, 0x4341), O1 = __byte_perm(A1, 0u, 0x4341);
    const uint32_t S0 = __byte_perm(E0, E1, 0x5432), S1 = __byte_perm(E1, E2, 0x5432);
    out0 = __byte_perm(E0 * w0 + (O0 * w1 + 0x00800080u), O0 * w0 + (S0 * w1 + 0x00800080u), 0x7351);
    out1 = __byte_perm(E1 * w0 + (O1 * w1 + 0x00800080u), O1 * w0 + (S1 * w1 + 0x00800080u), 0x7351);
}

// Angular modes: lane l generates, for it = 0,1, the 16 pixels (row 16*it + (l>>1), columns 16*(l&1)..+15) of the
// vertical-family prediction P_v (distance = row, position along the main reference = column): one (idx, f) pair per
// row.  Vertical modes store the four words directly (512 contiguous bytes per warp store).  Horizontal modes are
// P_v^T of the left reference: the rows go through a padded per-warp tile and come back as 4x4 byte blocks that are
// transposed in registers (8 PRMT per block).
__device__ __forceinline__ void intra_angular_rows(const uint32_t* __restrict__ strip32, int ref0, int ang, int lane, uint32_t (&w)[2][4])
{
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const int row = 16 * it + (lane >> 1), half = lane & 1;
        const int t = (row + 1) * ang, idx = t >> 5, f = t & 31;
        const int o = ref0 + 16 * half + idx + 1;                   // byte offset of ref[16*half + idx + 1] in the strip
        if ((ang & 31) == 0) {                                      // modes 2, 10, 18, 26, 34: every fraction is 0, rows are copies
            const uint32_t* p = strip32 + (o >> 2);
            const int sh = (o & 3) * 8;
            const uint32_t x0 = p[0], x1 = p[1], x2 = p[2], x3 = p[3], x4 = p[4];
            w[it][0] = __funnelshift_r(x0, x1, sh); w[it][1] = __funnelshift_r(x1, x2, sh);
            w[it][2] = __funnelshift_r(x2, x3, sh); w[it][3] = __funnelshift_r(x3, x4, sh);
        } else {
            intra_row16(strip32 + (o >> 2), (o & 3) * 8, f, w[it]);
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Mode decision of ONE block by one CTA of IDEC_WARPS warps (the body of intra32_decide_v2_kernel, also the first phase of the fused
// encode kernel in encode.cu); see the kernel's comment in intra.cu for the method.  On return (after a CTA barrier) sm.scost[0..34]
// holds the 35 costs.
// ------------------------------------------------------------------------------------------------
constexpr int IDEC_WARPS = 8;
constexpr int IDEC_NB = 2;                               // blocks decided per pass of a CTA: 70 (block, mode) items over 8 warps (35 leave 3 warps a mode short)

struct DecideSmem {
    __align__(16) uint8_t scur[IDEC_NB][2][1024];        // per block: [0] = block, [1] = its transpose
    __align__(16) int stc[IDEC_NB][2][32 * 32];          // T(cur), T(cur^T) in accumulator layout: [n-tile t][lane][register c], one 16-byte read per tile
    __align__(16) uint8_t sraw[IDEC_NB][144];
    __align__(16) uint8_t strip[IDEC_WARPS][IDEC_NB][2][INTRA_STRIP + 16];   // per warp and block: [0] the left-based working line (modes 2..17), [1] the top-based one
    uint32_t scost[IDEC_NB][35];
};

// the inputs of one pass in registers (thread tid: word tid of each current block, byte tid of each block's 129 reference bytes), loaded
// one pass ahead so that a CTA never waits on HBM at the top of a block
struct DecideIn { uint32_t w[IDEC_NB]; uint32_t r[IDEC_NB]; };

__device__ __forceinline__ void decide_load(DecideIn& in, const uint8_t* __restrict__ cur, const uint8_t* __restrict__ refs, size_t p, size_t n, int tid)
{
#pragma unroll
    for (int b = 0; b < IDEC_NB; b++) {
        in.w[b] = 0; in.r[b] = 0;
        if (p + b < n) {
            in.w[b] = __ldg(reinterpret_cast<const uint32_t*>(cur + (p + b) * 1024) + tid);
            if (tid < 129) in.r[b] = __ldg(refs + (p + b) * 129 + tid);
        }
    }
}

// +-1 matrix fragments, identical to satd8x8_imma_kernel: K position 16r+4q+i <-> sample 32s+8q+4r+i, column 8t+g
__device__ __forceinline__ void decide_hadamard_fragments(uint32_t (&B)[2][8][2], int g, int q)
{
#pragma unroll
    for (int s = 0; s < 2; s++)
#pragma unroll
        for (int t = 0; t < 8; t++)
#pragma unroll
            for (int r = 0; r < 2; r++) {
                uint32_t v = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int pix = 32 * s + 8 * q + 4 * r + i, nn = 8 * t + g;
                    v |= ((__popc(nn & pix) & 1) ? 0xFFu : 0x01u) << (8 * i);
                }
                B[s][t][r] = v;
            }
}

// Decides blocks p .. p+nblk-1 (nblk <= IDEC_NB) whose inputs are in `in`, then reloads `in` with the blocks at pNext (if any) while the modes run.
// On return (after a CTA barrier) sm.scost[b][0..34] holds the 35 costs of block b.
__device__ __forceinline__ void decide_blocks(DecideSmem& sm, const uint32_t (&B)[2][8][2], DecideIn& in, int nblk,
                                              const uint8_t* __restrict__ cur, const uint8_t* __restrict__ refs, size_t pNext, size_t n, int tid)
{
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
    const int cZero[4] = { 0, 0, 0, 0 };
    // this lane's rows: sub-blocks g and g+8 (same column band sx), sub-block rows q and 4+q
    const int sx = (g & 3) * 8;
    const int syA = (g >> 2) * 8, syB = syA + 16;

    {
        const int row = tid >> 3, c0 = (tid & 7) * 4;
#pragma unroll
        for (int b = 0; b < IDEC_NB; b++) {
            if (b < nblk) {
                const uint32_t w = in.w[b];
                reinterpret_cast<uint32_t*>(sm.scur[b][0])[tid] = w;
#pragma unroll
                for (int i = 0; i < 4; i++) sm.scur[b][1][(c0 + i) * 32 + row] = (uint8_t)(w >> (8 * i));
                if (tid < 129) sm.sraw[b][tid] = (uint8_t)in.r[b];
            }
        }
    }
    __syncthreads();
    decide_load(in, cur, refs, pNext, n, tid);            // the next pass's inputs travel while this pass computes
    if (warp < 2 * nblk) {                                // T(cur) and T(cur^T) of block warp >> 1
        const int b = warp >> 1, tr = warp & 1;
        uint32_t A[2][4];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const uint2 ra = *reinterpret_cast<const uint2*>(&sm.scur[b][tr][(syA + 4 * s + q) * 32 + sx]);
            const uint2 rb = *reinterpret_cast<const uint2*>(&sm.scur[b][tr][(syB + 4 * s + q) * 32 + sx]);
            A[s][0] = ra.x; A[s][2] = ra.y; A[s][1] = rb.x; A[s][3] = rb.y;
        }
#pragma unroll
        for (int t = 0; t < 8; t++) {
            int d[4];
            mma_u8s8(d, A[0], B[0][t][0], B[0][t][1], cZero);
            mma_u8s8(d, A[1], B[1][t][0], B[1][t][1], d);
            *reinterpret_cast<int4*>(&sm.stc[b][tr][(t * 32 + lane) * 4]) = make_int4(d[0], d[1], d[2], d[3]);
        }
    }
    // both working lines of this warp for every block of the pass, staged ONCE per block (they used to be rebuilt for every mode); a
    // negative-angle mode only rewrites the projected part ref[-|angle|..-1], which is all it reads below ref[0]
    for (int b = 0; b < nblk; b++) {
        const uint8_t* left = sm.sraw[b];
        const uint8_t* top = sm.sraw[b] + 64;
        for (int i = lane; i <= 71; i += 32) {
            sm.strip[warp][b][1][36 + i] = i > 64 ? (uint8_t)0 : top[i];
            sm.strip[warp][b][0][36 + i] = i > 64 ? (uint8_t)0 : (i == 0 ? top[0] : left[i - 1]);
        }
    }
    __syncthreads();

    // the (block, mode) items of the pass, round-robin over the warps
    for (int item = warp; item < 35 * nblk; item += IDEC_WARPS) {
        const int b = item >= 35 ? 1 : 0;
        const int mode = item - 35 * b;
        const uint8_t* left = sm.sraw[b];
        const uint8_t* top = sm.sraw[b] + 64;
        const bool isVer = mode >= 18;
        const int ang = c_intraAngle[mode];
        uint8_t* sref = sm.strip[warp][b][isVer ? 1 : 0] + 32 + 4;     // sref[i] = ref[i] at byte 36 of a strip (+4 keeps sref-32 word aligned)
        const uint32_t* strip32 = reinterpret_cast<const uint32_t*>(sm.strip[warp][b][isVer ? 1 : 0]);
        uint32_t A[2][4];
        if (mode >= 2) {
            if (ang < 0) {
                const int inv = c_intraInvMode[mode];
                __syncwarp();
                const int k = lane + 1;
                if (-k >= ang) {
                    const int sidx = (k * inv + 128) >> 8;
                    sref[-k] = isVer ? left[sidx - 1] : top[sidx];
                }
            }
            __syncwarp();
#pragma unroll
            for (int s = 0; s < 2; s++)
#pragma unroll
                for (int sel = 0; sel < 2; sel++) {
                    const int row = (sel ? syB : syA) + 4 * s + q;             // distance from the main reference
                    const int t = (row + 1) * ang, idx = t >> 5, f = t & 31;
                    const int o = 32 + 4 + sx + idx + 1;                        // byte offset of ref[sx+idx+1] in the strip
                    intra_row8(strip32 + (o >> 2), (o & 3) * 8, f, A[s][sel], A[s][2 + sel]);
                }
        } else if (mode == 1) {
            int sum = left[lane] + top[1 + lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const uint32_t dc = (uint32_t)((sum + 32) >> 6) * 0x01010101u;
#pragma unroll
            for (int s = 0; s < 2; s++)
#pragma unroll
                for (int r = 0; r < 4; r++) A[s][r] = dc;
        } else {
            const int tr = top[33], bl = left[32];
#pragma unroll
            for (int s = 0; s < 2; s++)
#pragma unroll
                for (int sel = 0; sel < 2; sel++) {
                    const int row = (sel ? syB : syA) + 4 * s + q;
                    uint32_t wlo = 0, whi = 0;
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const int col = sx + c;
                        const uint32_t v = (uint32_t)(((31 - col) * left[row] + (col + 1) * tr + (31 - row) * top[1 + col] + (row + 1) * bl + 32) >> 6);
                        if (c < 4) wlo |= v << (8 * c); else whi |= v << (8 * (c - 4));
                    }
                    A[s][sel] = wlo;
                    A[s][2 + sel] = whi;
                }
        }
        // T(pred) for the 16 sub-blocks, |T(cur) - T(pred)| against the parked transform (transposed one for 2..17)
        const int* tc = sm.stc[b][(mode >= 2 && !isVer) ? 1 : 0];
        unsigned sa0 = 0, sa1 = 0, sb0 = 0, sb1 = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            int d[4];
            mma_u8s8(d, A[0], B[0][t][0], B[0][t][1], cZero);
            mma_u8s8(d, A[1], B[1][t][0], B[1][t][1], d);
            const int4 c4v = *reinterpret_cast<const int4*>(&tc[(t * 32 + lane) * 4]);
            sa0 = __sad(d[0], c4v.x, sa0);
            sa1 = __sad(d[1], c4v.y, sa1);
            sb0 = __sad(d[2], c4v.z, sb0);
            sb1 = __sad(d[3], c4v.w, sb1);
        }
        unsigned sadA = sa0 + sa1, sadB = sb0 + sb1;                // sub-blocks g and g+8, partial over this lane's columns
        sadA += __shfl_xor_sync(0xffffffffu, sadA, 1); sadB += __shfl_xor_sync(0xffffffffu, sadB, 1);
        sadA += __shfl_xor_sync(0xffffffffu, sadA, 2); sadB += __shfl_xor_sync(0xffffffffu, sadB, 2);
        unsigned c4 = q == 0 ? ((sadA + 2) >> 2) + ((sadB + 2) >> 2) : 0u;
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) c4 += __shfl_xor_sync(0xffffffffu, c4, o);
        if (lane == 0) sm.scost[b][mode] = c4;
    }
    __syncthreads();
}

// After decide_blocks: warp b (b < nblk) writes block b's 35 costs (if wanted) and returns its best mode in every lane -- the smallest cost, ties to the
// smallest mode (key = cost * 64 + mode; a cost is < 2^20), folded with shuffles instead of a 35-step scan by one thread.
__device__ __forceinline__ int decide_output(const DecideSmem& sm, int b, uint32_t* __restrict__ costOut, int lane)
{
    const uint32_t c0 = sm.scost[b][lane], c1 = lane < 3 ? sm.scost[b][32 + lane] : 0u;
    if (costOut) {
        costOut[lane] = c0;
        if (lane < 3) costOut[32 + lane] = c1;
    }
    uint32_t key = c0 * 64u + (uint32_t)lane;
    if (lane < 3) key = min(key, c1 * 64u + 32u + (uint32_t)lane);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
    return (int)(key & 63u);
}

} // namespace x266
