// encode.cu -- the closed intra block loop in ONE kernel ("next" rows N1 + N3 of SURVEY 8(f)):
//
//   decide (35-mode SATD, tensor cores)  ->  best-mode prediction  ->  residual  ->  DCT32 (4/11)  ->  quantise / de-quantise
//   ->  IDCT32 (7/12)  ->  reconstruction = clip8(prediction + residual')
//
// i.e. both channels of the RTL's intra unit, `Decide` and `Recon` (src/mkIntra32-wip.bsv:39-48), around the transform of
// src_tb/dct32.c:197-198 and its inverse with the same matrix (dct32.c:30-64).  Per block 1 KiB of pixels + 129 reference bytes
// come in and 2 KiB of levels + 1 KiB of reconstruction + the mode go out; the prediction, the residual, the coefficients and
// the de-quantised coefficients never leave the SM.
//
// A CTA of 8 warps takes 8 consecutive blocks: the decision of each pair of blocks is made by the whole CTA (decide_blocks, intra_dev.cuh:
// one warp per mode, 16 IMMA per mode), then every warp reconstructs one of the 8 blocks on its own (encode_block_warp below: the
// register-resident byte-plane IMMA choreography of dct_imma.cu, forward with the residual formed by linearity
// G*(c - p) = G*c + (-G)*p on raw pixels, inverse from a per-warp shared-memory tile).  intra32_recon_kernel is the `Recon`
// channel alone: the mode is given.
//
// The quantiser is a STUB in the sense of SURVEY 8(f) N3 -- the reference has none -- with the arithmetic of the HEVC/VVC test
// models' flat (no scaling list) intra quantiser for 8-bit video and a 32x32 transform (transformShift = 2):
//     level = sign(c) * min(32767, (|c| * qScale[qp % 6] + (171 << (qBits - 9))) >> qBits),   qBits = 16 + qp / 6
//     c'    = clip16((level * (iqScale[qp % 6] << (qp / 6)) + 8) >> 4)
// PARITY: decision = pinned SATD on the BSV-pinned predictor; transform = pinned; quantiser + inverse: ours (no reference artefact exists for them).
#include "common.cuh"
#include "kernels.h"
#include "intra_dev.cuh"
#include "dct_frag.cuh"

namespace x266 {

struct QuantParams {
    int qScale, qBits, qAdd;          // forward
    int iqScale, iqAdd, iqShift;      // inverse (iqScale already shifted by qp / 6)
};

static QuantParams make_quant(int qp)
{
    static const int q[6] = { 26214, 23302, 20560, 18396, 16384, 14564 };
    static const int iq[6] = { 40, 45, 51, 57, 64, 72 };
    QuantParams p;
    p.qScale = q[qp % 6];
    p.qBits = 16 + qp / 6;
    p.qAdd = 171 << (p.qBits - 9);
    p.iqScale = iq[qp % 6] << (qp / 6);
    p.iqShift = 4;
    p.iqAdd = 8;
    return p;
}

__device__ __forceinline__ int quant1(int c, const QuantParams& qp)
{
    const int a = min((int)(((unsigned)abs(c) * (unsigned)qp.qScale + (unsigned)qp.qAdd) >> qp.qBits), 32767);
    return c < 0 ? -a : a;
}

__device__ __forceinline__ int dequant1(int level, const QuantParams& qp)
{
    return (level * qp.iqScale + qp.iqAdd) >> qp.iqShift;            // clip16 happens in the saturating pack
}

constexpr int ENC_WARPS = IDEC_WARPS;

struct EncodeWarpSmem {
    __align__(16) uint8_t pred[32 * 32];         // best-mode prediction, row-major
    __align__(16) int16_t coef[32 * 32];         // de-quantised coefficients, row-major (input of the inverse transform)
    __align__(16) uint8_t raw[2][144];           // left[64] | top[65] of this block and of the warp's next one, each as the aligned words that cover it
    __align__(16) uint8_t strip[INTRA_STRIP + 16];
};

// 32x32 prediction of `mode` into ws.pred (one warp).  Angular rows come from the SWAR generator of the predictor kernel
// (intra_angular_rows: the vertical-family prediction P_v of the main reference); horizontal modes are P_v^T.
// The 129 reference bytes of a block travel into ws.raw[buf] ahead of their use, with no register staging: when the refs array is 4-byte aligned
// the 33 aligned words that cover [refsBlock, refsBlock + 129) are copied asynchronously (the bytes then start at offset (address & 3) of the buffer);
// the last block of the array, whose covering words may end past it, and unaligned arrays are copied byte by byte.  Always commits one group.
__device__ __forceinline__ int stage_refs(EncodeWarpSmem& ws, int buf, const uint8_t* __restrict__ refsBlock, bool wordsOk, int lane)
{
    const int a = wordsOk ? (int)(reinterpret_cast<uintptr_t>(refsBlock) & 3) : 0;
    if (wordsOk) {
        const uint8_t* src = refsBlock - a;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ws.raw[buf]);
        cp_async4(dst + 4u * lane, src + 4 * lane);
        if (lane == 0) cp_async4(dst + 128u, src + 128);
    } else {
        for (int i = lane; i < 129; i += 32) ws.raw[buf][i] = refsBlock[i];
    }
    cp_async_commit();
    return a;
}

__device__ __forceinline__ void predict_to_tile(EncodeWarpSmem& ws, const uint8_t* __restrict__ left, int mode, int lane)
{
    const uint8_t* top = left + 64;                               // left = the block's 129 reference bytes in shared memory (complete and visible)
    uint32_t* pred32 = reinterpret_cast<uint32_t*>(ws.pred);
    if (mode >= 2) {
        const bool isVer = mode >= 18;
        const int ang = c_intraAngle[mode];
        uint8_t* sref = ws.strip + 36;                            // sref[i] = ref[i], ref[0] word aligned
        for (int i = lane; i <= 75; i += 32) sref[i] = i > 64 ? (uint8_t)0 : (isVer ? top[i] : (i == 0 ? top[0] : left[i - 1]));
        if (ang < 0) {
            const int inv = c_intraInvMode[mode];
            const int k = lane + 1;                               // projects ref[-k], k = 1..32
            if (-k >= ang) {
                const int s = (k * inv + 128) >> 8;
                sref[-k] = isVer ? left[s - 1] : top[s];
            }
        }
        __syncwarp();
        uint32_t w[2][4];
        intra_angular_rows(reinterpret_cast<const uint32_t*>(ws.strip), 36, ang, lane, w);
#pragma unroll
        for (int it = 0; it < 2; it++) {
            const int row = 16 * it + (lane >> 1), c0 = 16 * (lane & 1);
            if (isVer) {
                *reinterpret_cast<uint4*>(&ws.pred[row * 32 + c0]) = make_uint4(w[it][0], w[it][1], w[it][2], w[it][3]);
            } else {
#pragma unroll
                for (int j = 0; j < 16; j++) ws.pred[(c0 + j) * 32 + row] = (uint8_t)(w[it][j >> 2] >> (8 * (j & 3)));
            }
        }
    } else if (mode == 1) {
        int s = left[lane] + top[1 + lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const uint32_t dc = (uint32_t)((s + 32) >> 6) * 0x01010101u;
#pragma unroll
        for (int it = 0; it < 8; it++) pred32[it * 32 + lane] = dc;
    } else {
        const int rsub = lane >> 3, c0 = (lane & 7) * 4;
        const int tr = top[33], bl = left[32];
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int row = 4 * it + rsub;
            uint32_t packed = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int col = c0 + j;
                const int v = ((31 - col) * left[row] + (col + 1) * tr + (31 - row) * top[1 + col] + (row + 1) * bl + 32) >> 6;
                packed |= (uint32_t)v << (8 * j);
            }
            pred32[it * 32 + lane] = packed;
        }
    }
    __syncwarp();
}

// One warp: prediction of `mode` -> residual -> DCT32 (4/11) -> quant -> level out; dequant -> IDCT32 (7/12) -> recon out.
__device__ __forceinline__ void encode_block_warp(EncodeWarpSmem& ws, const uint8_t* __restrict__ curBlock, const uint8_t* __restrict__ refsSmem,
                                                  int mode, const QuantParams qp, int16_t* __restrict__ level, uint8_t* __restrict__ recon, int lane)
{
    const int g = lane >> 2, q = lane & 3;
    const int cZero[4] = { 0, 0, 0, 0 };
    // this lane's 8 pixels of the rows j = 8t + g, columns 8q..8q+7 (the B fragments of pass 1 are raw pixels)
    uint2 bc[4];
#pragma unroll
    for (int t = 0; t < 4; t++) bc[t] = *reinterpret_cast<const uint2*>(curBlock + (8 * t + g) * 32 + 8 * q);
    predict_to_tile(ws, refsSmem, mode, lane);

    // ---- forward transform, src_tb/dct32.c:197-198 with shifts 4 / 11 on (cur - pred): the choreography of frame_resi_dct32_kernel
    {
        uint32_t A1[2][4], A1n[2][4], A2[2][4];
#pragma unroll
        for (int m = 0; m < 2; m++)
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
        const int cAdd1[4] = { 8, 8, 8, 8 };
        const int cAdd2[4] = { 1024, 1024, 1024, 1024 };
        uint32_t B2L[4][2], B2H[4][2];
#pragma unroll
        for (int m = 0; m < 2; m++) {
            int r[4][4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const uint2 bp = *reinterpret_cast<const uint2*>(&ws.pred[(8 * t + g) * 32 + 8 * q]);
                int acc[4];
                mma_s8u8(acc, A1[m], bc[t].x, bc[t].y, cAdd1);
                mma_s8u8(acc, A1n[m], bp.x, bp.y, acc);
#pragma unroll
                for (int c = 0; c < 4; c++) r[t][c] = acc[c] >> 4;
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
#pragma unroll
        for (int m2 = 0; m2 < 2; m2++) {
            int r[4][4];
#pragma unroll
            for (int t2 = 0; t2 < 4; t2++) {
                int dl[4], dh[4];
                mma_s8u8(dl, A2[m2], B2L[t2][0], B2L[t2][1], cAdd2);
                mma_s8s8(dh, A2[m2], B2H[t2][0], B2H[t2][1], cZero);
#pragma unroll
                for (int c = 0; c < 4; c++) r[t2][c] = (int)(short)((dl[c] + dh[c] * 256) >> 11);      // the truncating int16 store of dct32.c:128-151
            }
            // row 16 m2 + 8 h + g, columns 8q + 2 t2 + e  <->  r[t2][2h + e]: quantise, keep the level, de-quantise into the tile
#pragma unroll
            for (int h = 0; h < 2; h++) {
                int lv[4][2], dq[4][2];
#pragma unroll
                for (int t2 = 0; t2 < 4; t2++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        lv[t2][e] = quant1(r[t2][2 * h + e], qp);
                        dq[t2][e] = dequant1(lv[t2][e], qp);
                    }
                const int row = 16 * m2 + 8 * h + g;
                st_global_stream(level + row * 32 + q * 8,
                                 make_uint4(prmt(lv[0][0], lv[0][1], 0x5410), prmt(lv[1][0], lv[1][1], 0x5410),
                                            prmt(lv[2][0], lv[2][1], 0x5410), prmt(lv[3][0], lv[3][1], 0x5410)));
                *reinterpret_cast<uint4*>(&ws.coef[row * 32 + q * 8]) =
                    make_uint4(pack_sat16(dq[0][0], dq[0][1]), pack_sat16(dq[1][0], dq[1][1]), pack_sat16(dq[2][0], dq[2][1]), pack_sat16(dq[3][0], dq[3][1]));
            }
        }
    }
    __syncwarp();

    // ---- inverse transform (vertical pass first, shifts 7 / 12, saturating): the choreography of idct32_imma_kernel
    {
        uint32_t A1[2][4];
#pragma unroll
        for (int m = 0; m < 2; m++)
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int y = 16 * m + g + 8 * (r & 1);
                const int kb = 16 * (r >> 1) + 4 * q;
                A1[m][r] = pack4(c_g8.v[kb + 0][y], c_g8.v[kb + 1][y], c_g8.v[kb + 2][y], c_g8.v[kb + 3][y]);
            }
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
        const int cAdd1[4] = { 64, 64, 64, 64 };
        const int cAdd2[4] = { 2048, 2048, 2048, 2048 };
        // columns 4g..4g+3 of rows 4q+i and 16+4q+i, 4x4 byte transposes -> B1 fragments
        uint32_t BL[4][2], BH[4][2];
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
            uint2 w[4];
#pragma unroll
            for (int i = 0; i < 4; i++) w[i] = *reinterpret_cast<const uint2*>(&ws.coef[(16 * hh + 4 * q + i) * 32 + g * 4]);
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const uint32_t r0 = half ? w[0].y : w[0].x, r1 = half ? w[1].y : w[1].x;
                const uint32_t r2 = half ? w[2].y : w[2].x, r3 = half ? w[3].y : w[3].x;
                const uint32_t t0 = prmt(r0, r1, 0x5140), t1 = prmt(r2, r3, 0x5140);
                const uint32_t t2 = prmt(r0, r1, 0x7362), t3 = prmt(r2, r3, 0x7362);
                BL[2 * half][hh] = prmt(t0, t1, 0x5410);     BH[2 * half][hh] = prmt(t0, t1, 0x7632);
                BL[2 * half + 1][hh] = prmt(t2, t3, 0x5410); BH[2 * half + 1][hh] = prmt(t2, t3, 0x7632);
            }
        }
#pragma unroll
        for (int m = 0; m < 2; m++) {
            int r[4][4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                int dl[4], dh[4];
                mma_s8u8(dl, A1[m], BL[t][0], BL[t][1], cAdd1);
                mma_s8s8(dh, A1[m], BH[t][0], BH[t][1], cZero);
#pragma unroll
                for (int c = 0; c < 4; c++) r[t][c] = (dl[c] + dh[c] * 256) >> 7;
            }
            uint32_t AL[4], AH[4];
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int hi = 0; hi < 2; hi++) {
                    const uint32_t p0 = pack_sat16(r[2 * hi][2 * h], r[2 * hi][2 * h + 1]);
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
                for (int c = 0; c < 4; c++) r2[tx][c] = max(-32768, min(32767, (dl[c] + dh[c] * 256) >> 12));
            }
            // row 16 m + 8 h + g, columns 8q + 2 tx + e  <->  r2[tx][2h + e]: add the prediction, clip to 8 bits
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int row = 16 * m + 8 * h + g;
                const uint2 pp = *reinterpret_cast<const uint2*>(&ws.pred[row * 32 + 8 * q]);
                uint32_t o[2] = { 0, 0 };
#pragma unroll
                for (int tx = 0; tx < 4; tx++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int k = 2 * tx + e;
                        const int pix = (int)(((k < 4 ? pp.x : pp.y) >> (8 * (k & 3))) & 0xFFu);
                        const int v = max(0, min(255, pix + r2[tx][2 * h + e]));
                        o[k >> 2] |= (uint32_t)v << (8 * (k & 3));
                    }
                st_global_stream_v2(recon + row * 32 + 8 * q, make_uint2(o[0], o[1]));
            }
        }
    }
    __syncwarp();
}

struct EncodeSmem {
    DecideSmem dec;
    EncodeWarpSmem w[ENC_WARPS];
    int best[ENC_WARPS];
};

__global__ void __launch_bounds__(ENC_WARPS * 32, 2)
intra32_encode_kernel(const uint8_t* __restrict__ cur, const uint8_t* __restrict__ refs, size_t n, const QuantParams qp,
                      int16_t* __restrict__ level, uint8_t* __restrict__ recon, int32_t* __restrict__ bestMode, uint32_t* __restrict__ cost)
{
    extern __shared__ __align__(16) uint8_t smemRaw[];
    EncodeSmem& sm = *reinterpret_cast<EncodeSmem*>(smemRaw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t nGroups = (n + ENC_WARPS - 1) / ENC_WARPS;
    static_assert(ENC_WARPS % IDEC_NB == 0, "a group of blocks is a whole number of decision passes");
    DecideIn in;                                                  // inputs of the next decision pass (registers), loaded one pass ahead
    decide_load(in, cur, refs, (size_t)blockIdx.x * ENC_WARPS, n, tid);
    const bool refsWords = (reinterpret_cast<uintptr_t>(refs) & 3) == 0;
    for (size_t grp = blockIdx.x; grp < nGroups; grp += gridDim.x) {
        // this warp's own block: its reference bytes travel into shared memory while the CTA decides
        const size_t pw = grp * ENC_WARPS + warp;
        const int ra = pw < n ? stage_refs(sm.w[warp], 0, refs + pw * 129, refsWords && pw + 1 < n, lane) : 0;
        {
            uint32_t B[2][8][2];                                  // +-1 fragments: live only during the decision phase
            decide_hadamard_fragments(B, lane >> 2, lane & 3);
            for (int i = 0; i < ENC_WARPS; i += IDEC_NB) {
                const size_t p = grp * ENC_WARPS + i;
                // the pass after this one: the next blocks of this group, or the first blocks of this CTA's next group (past the end: loads nothing)
                const size_t pNext = i + IDEC_NB < ENC_WARPS ? p + IDEC_NB : (grp + gridDim.x) * ENC_WARPS;
                if (p >= n) break;
                const int nblk = (n - p) < (size_t)IDEC_NB ? (int)(n - p) : IDEC_NB;
                decide_blocks(sm.dec, B, in, nblk, cur, refs, pNext, n, tid);
                if (warp < nblk) {
                    const int bm = decide_output(sm.dec, warp, cost ? cost + (p + warp) * 35 : nullptr, lane);
                    if (lane == 0) { sm.best[i + warp] = bm; bestMode[p + warp] = bm; }
                }
            }
        }
        cp_async_wait<0>();
        __syncthreads();
        if (pw < n) encode_block_warp(sm.w[warp], cur + pw * 1024, sm.w[warp].raw[0] + ra, sm.best[warp], qp, level + pw * 1024, recon + pw * 1024, lane);
        __syncthreads();
    }
}

// the `Recon` channel alone: the mode of every block is given
__global__ void __launch_bounds__(ENC_WARPS * 32, 2)
intra32_recon_kernel(const uint8_t* __restrict__ cur, const uint8_t* __restrict__ refs, const uint8_t* __restrict__ mode, size_t n,
                     const QuantParams qp, int16_t* __restrict__ level, uint8_t* __restrict__ recon)
{
    __shared__ EncodeWarpSmem ws[ENC_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the reference bytes and the mode of the warp's NEXT block are in flight while this one is reconstructed (ncu on the first form: 30 % of
    // the stall samples were long_scoreboard on the reference bytes at the top of a block)
    const bool refsWords = (reinterpret_cast<uintptr_t>(refs) & 3) == 0;
    const size_t stride = (size_t)gridDim.x * ENC_WARPS;
    size_t p = (size_t)blockIdx.x * ENC_WARPS + warp;
    if (p >= n) return;
    int buf = 0;
    int a = stage_refs(ws[warp], 0, refs + p * 129, refsWords && p + 1 < n, lane);
    int m = mode[p];
    for (; p < n; p += stride) {
        const size_t pn = p + stride;
        int an = 0, mn = 1;
        if (pn < n) { an = stage_refs(ws[warp], buf ^ 1, refs + pn * 129, refsWords && pn + 1 < n, lane); mn = mode[pn]; }
        else cp_async_commit();
        cp_async_wait<1>();                      // everything but the group just committed: this block's bytes have landed
        __syncwarp();
        encode_block_warp(ws[warp], cur + p * 1024, ws[warp].raw[buf] + a, m > 34 ? 1 : m, qp, level + p * 1024, recon + p * 1024, lane);
        __syncwarp();                            // all lanes are done with raw[buf] before the iteration after next stages into it
        buf ^= 1; a = an; m = mn;
    }
}

// quantise and/or de-quantise n coefficients (the stub on its own): level and dq may each be null
__global__ void __launch_bounds__(256)
quant_dequant_kernel(const int16_t* __restrict__ coef, int16_t* __restrict__ level, int16_t* __restrict__ dq, size_t nVec, const QuantParams qp)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nVec; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = ld_global_stream(coef + i * 8);
        const uint32_t w[4] = { v.x, v.y, v.z, v.w };
        uint32_t lo[4], dqo[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int l0 = quant1((int)(short)(w[k] & 0xFFFFu), qp), l1 = quant1((int)(short)(w[k] >> 16), qp);
            lo[k] = prmt(l0, l1, 0x5410);
            dqo[k] = pack_sat16(dequant1(l0, qp), dequant1(l1, qp));
        }
        if (level) st_global_stream(level + i * 8, make_uint4(lo[0], lo[1], lo[2], lo[3]));
        if (dq) st_global_stream(dq + i * 8, make_uint4(dqo[0], dqo[1], dqo[2], dqo[3]));
    }
}

cudaError_t launch_intra32_encode(const uint8_t* cur, const uint8_t* refs, size_t n, int qp, int16_t* level, uint8_t* recon,
                                  int32_t* bestMode, uint32_t* cost, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    if (qp < 0 || qp > 51) return cudaErrorInvalidValue;
    static std::atomic<bool> attrSet[64];
    int dev = 0;
    cudaGetDevice(&dev);
    const int smem = (int)sizeof(EncodeSmem);
    if (dev < 0 || dev >= 64 || !attrSet[dev]) {
        cudaError_t e = cudaFuncSetAttribute(intra32_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attrSet[dev] = true;
    }
    const size_t groups = (n + ENC_WARPS - 1) / ENC_WARPS;
    const size_t cap = (size_t)sm_count() * resident_ctas_per_sm((const void*)intra32_encode_kernel, ENC_WARPS * 32, smem);
    intra32_encode_kernel<<<(unsigned)(groups < cap ? groups : cap), ENC_WARPS * 32, smem, st>>>(cur, refs, n, make_quant(qp), level, recon, bestMode, cost);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_intra32_recon(const uint8_t* cur, const uint8_t* refs, const uint8_t* mode, size_t n, int qp, int16_t* level,
                                 uint8_t* recon, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    if (qp < 0 || qp > 51) return cudaErrorInvalidValue;
    const size_t want = (n + ENC_WARPS - 1) / ENC_WARPS;
    const size_t cap = (size_t)sm_count() * resident_ctas_per_sm((const void*)intra32_recon_kernel, ENC_WARPS * 32, 0);
    intra32_recon_kernel<<<(unsigned)(want < cap ? want : cap), ENC_WARPS * 32, 0, st>>>(cur, refs, mode, n, make_quant(qp), level, recon);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_quant_dequant(const int16_t* coef, int16_t* level, int16_t* dq, size_t nCoef, int qp, cudaStream_t st)
{
    if (nCoef == 0) return cudaSuccess;
    if (qp < 0 || qp > 51 || (nCoef & 7)) return cudaErrorInvalidValue;
    const size_t nVec = nCoef / 8, want = (nVec + 255) / 256, cap = (size_t)sm_count() * 8;
    quant_dequant_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(coef, level, dq, nVec, make_quant(qp));
    count_launch();
    return cudaGetLastError();
}

} // namespace x266
