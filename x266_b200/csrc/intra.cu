// intra.cu -- 32x32 intra prediction, 35 modes (0 planar, 1 DC, 2..34 angular).
//
// Spec source: src/mkIntra32-wip.bsv (work in progress in the reference, no C model -> PARITY UNPINNED,
// see DESIGN.md):  reference samples left[64] / top[65] (:34-37), per-row index and fraction tables
// = ((k+1)*angle)>>5 and &31 (:75-112), inverse-angle projection of the side reference for negative
// angles (:151-220), 2-tap interpolation (:352-368), DC (:388-392).
// One warp per prediction, 4 output pixels per lane per step, 128-byte coalesced stores.
#include "common.cuh"
#include "kernels.h"

namespace x266 {

__constant__ int c_intraAngle[35] = { 0, 0, 32, 26, 21, 17, 13, 9, 5, 2, 0, -2, -5, -9, -13, -17, -21, -26,
                                      -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32 };
// 8192/|angle| rounded, for angle = -2,-5,-9,-13,-17,-21,-26,-32
__constant__ int c_intraInv[8] = { 4096, 1638, 910, 630, 482, 390, 315, 256 };

// One warp per prediction.  Lane l produces, for it = 0..7, the 4 pixels (row 4*it + (l>>3), columns
// 4*(l&7) .. +3) and stores them as one 32-bit word: every warp store is 128 contiguous bytes (4 rows).
// The reference line ref[-32..65] lives in a per-warp shared-memory strip; VER selects which of
// (row, column) is the distance from the main reference so that the loop-invariant index/fraction
// computations are hoisted (per row for vertical modes, per lane for horizontal modes).
constexpr int INTRA_WARPS = 8;
constexpr int INTRA_STRIP = 112;            // 32 (negative part) + 66 + padding, multiple of 16

template <bool VER>
__device__ __forceinline__ void intra_angular(const uint8_t* __restrict__ ref /* -> ref[0] */, int ang, int lane, uint32_t* out)
{
    const int rsub = lane >> 3, c0 = (lane & 7) * 4;
#pragma unroll
    for (int it = 0; it < 8; it++) {
        const int row = 4 * it + rsub;
        uint32_t packed = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int col = c0 + j;
            const int xr = VER ? col : row;
            const int yd = VER ? row : col;
            const int t = (yd + 1) * ang;
            const int idx = t >> 5, f = t & 31;
            const int a = ref[xr + idx + 1], b = ref[xr + idx + 2];
            const int v = ((32 - f) * a + f * b + 16) >> 5;          // f == 0 gives exactly a
            packed |= (uint32_t)v << (8 * j);
        }
        out[it * 32 + lane] = packed;        // word index (4*it + rsub)*8 + (lane&7) == it*32 + lane
    }
}

__global__ void __launch_bounds__(INTRA_WARPS * 32)
intra32_kernel(const uint8_t* __restrict__ refs, const uint8_t* __restrict__ modes, uint8_t* __restrict__ pred, size_t n)
{
    __shared__ __align__(16) uint8_t strip[INTRA_WARPS][INTRA_STRIP];
    __shared__ __align__(16) uint8_t raw[INTRA_WARPS][144];      // left[64] | top[65]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sref = strip[warp] + 32;                            // sref[i] = ref[i], i in -32..65
    uint8_t* sraw = raw[warp];

    for (size_t p = (size_t)blockIdx.x * INTRA_WARPS + warp; p < n; p += (size_t)gridDim.x * INTRA_WARPS) {
        const int mode = modes[p] > 34 ? 1 : modes[p];           // host API rejects > 34; keep device reads in range
        const uint8_t* src = refs + p * 129;
#pragma unroll
        for (int i = lane; i < 129; i += 32) sraw[i] = src[i];
        __syncwarp();
        const uint8_t* left = sraw;          // left[i] = pixel (-1, i)
        const uint8_t* top = sraw + 64;      // top[0] = corner, top[1+i] = pixel (i, -1)
        uint32_t* out = reinterpret_cast<uint32_t*>(pred + p * 1024);
        const bool isVer = mode >= 18;
        const int ang = c_intraAngle[mode];

        if (mode >= 2) {
#pragma unroll
            for (int i = lane; i <= 65; i += 32) sref[i] = i > 64 ? (uint8_t)0 : (isVer ? top[i] : (i == 0 ? top[0] : left[i - 1]));
            if (ang < 0) {
                int inv = 0;
#pragma unroll
                for (int a = 0; a < 8; a++) if (c_intraAngle[11 + a] == ang) inv = c_intraInv[a];
                const int k = lane + 1;                           // projects ref[-k], k = 1..32
                if (-k >= ang) {
                    const int s = (k * inv + 128) >> 8;
                    sref[-k] = isVer ? left[s - 1] : top[s];
                }
            }
            __syncwarp();
            if (isVer) intra_angular<true>(sref, ang, lane, out);
            else intra_angular<false>(sref, ang, lane, out);
        } else if (mode == 1) {
            int s = left[lane] + top[1 + lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const uint32_t dc = (uint32_t)((s + 32) >> 6) * 0x01010101u;
#pragma unroll
            for (int it = 0; it < 8; it++) out[it * 32 + lane] = dc;
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
                out[it * 32 + lane] = packed;
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// "Next" row N1: fused intra prediction -> residual -> SATD mode decision (the RTL's `Decide`
// channel, src/mkIntra32-wip.bsv:39-48).  For one 32x32 block and each of the 35 modes the prediction is
// generated in registers, subtracted from the current block and costed with the reference SATD
// (src_tb/satd.c:31-118) on its 16 8x8 sub-blocks:  cost[mode] = sum_sb ((sum|H d H^T| + 2) >> 2).
// Neither the prediction nor the residual ever reaches HBM (1 KiB + 129 B in, 35 x 4 B out per block).
// One CTA per block; a half-warp owns one mode at a time, lane <-> 8x8 sub-block.
// ------------------------------------------------------------------------------------------------
constexpr int IDEC_WARPS = 8;

template <int STRIDE>
__device__ __forceinline__ void had8i(int* v)
{
#pragma unroll
    for (int dist = 4; dist >= 1; dist >>= 1)
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (!(i & dist)) {
                const int a = v[i * STRIDE], b = v[(i + dist) * STRIDE];
                v[i * STRIDE] = a + b;
                v[(i + dist) * STRIDE] = a - b;
            }
}

__global__ void __launch_bounds__(IDEC_WARPS * 32)
intra32_decide_kernel(const uint8_t* __restrict__ cur, const uint8_t* __restrict__ refs, uint32_t* __restrict__ cost,
                      int32_t* __restrict__ bestMode, size_t n)
{
    __shared__ __align__(16) uint8_t scur[1024];
    __shared__ uint8_t sraw[144];
    __shared__ __align__(16) uint8_t strip[IDEC_WARPS * 2][INTRA_STRIP];
    __shared__ uint32_t scost[35];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int half = lane >> 4, sb = lane & 15;
    const int sx = (sb & 3) * 8, sy = (sb >> 2) * 8;
    uint8_t* sref = strip[warp * 2 + half] + 32;

    for (size_t p = blockIdx.x; p < n; p += gridDim.x) {
        reinterpret_cast<uint32_t*>(scur)[tid] = reinterpret_cast<const uint32_t*>(cur + p * 1024)[tid];
        if (tid < 129) sraw[tid] = refs[p * 129 + tid];
        __syncthreads();
        const uint8_t* left = sraw;
        const uint8_t* top = sraw + 64;

        for (int m0 = warp * 2; m0 < 35; m0 += IDEC_WARPS * 2) {
            const int mode = m0 + half;
            const bool active = mode < 35;
            const int md = active ? mode : 1;
            const bool isVer = md >= 18;
            const int ang = c_intraAngle[md];
            // ---- per-mode reference strip, built by the 16 lanes of this half-warp
            if (md >= 2) {
                for (int i = sb; i <= 65; i += 16) sref[i] = i > 64 ? (uint8_t)0 : (isVer ? top[i] : (i == 0 ? top[0] : left[i - 1]));
                if (ang < 0) {
                    int inv = 0;
#pragma unroll
                    for (int a = 0; a < 8; a++) if (c_intraAngle[11 + a] == ang) inv = c_intraInv[a];
                    for (int k = sb + 1; k <= 32; k += 16)
                        if (-k >= ang) {
                            const int s = (k * inv + 128) >> 8;
                            sref[-k] = isVer ? left[s - 1] : top[s];
                        }
                }
            }
            int dc = 0;
            if (md == 1) {
                for (int i = 0; i < 32; i++) dc += left[i] + top[1 + i];
                dc = (dc + 32) >> 6;
            }
            __syncwarp();
            // ---- residual of this lane's 8x8 sub-block
            int d[64];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int row = sy + r;
                const uint2 cw = *reinterpret_cast<const uint2*>(scur + row * 32 + sx);
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const int col = sx + c;
                    int v;
                    if (md >= 2) {
                        const int xr = isVer ? col : row, yd = isVer ? row : col;
                        const int t = (yd + 1) * ang, idx = t >> 5, f = t & 31;
                        v = ((32 - f) * sref[xr + idx + 1] + f * sref[xr + idx + 2] + 16) >> 5;
                    } else if (md == 1) {
                        v = dc;
                    } else {
                        v = ((31 - col) * left[row] + (col + 1) * top[33] + (31 - row) * top[1 + col] + (row + 1) * left[32] + 32) >> 6;
                    }
                    const int px = (int)(((c < 4 ? cw.x : cw.y) >> (8 * (c & 3))) & 0xFF);
                    d[r * 8 + c] = px - v;
                }
                had8i<1>(&d[r * 8]);
            }
            unsigned sad = 0;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                had8i<8>(&d[c]);
#pragma unroll
                for (int r = 0; r < 8; r++) sad = __sad(d[r * 8 + c], 0, sad);   // 9-bit residuals: no int16 wrap reachable
            }
            unsigned c4 = (sad + 2) >> 2;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) c4 += __shfl_xor_sync(0xffffffffu, c4, o);      // sum over the 16 sub-blocks
            if (active && sb == 0) scost[mode] = c4;
            __syncwarp();
        }
        __syncthreads();
        if (tid < 35) cost[p * 35 + tid] = scost[tid];
        if (tid == 0) {
            unsigned bc = scost[0];
            int bm = 0;
            for (int m = 1; m < 35; m++) if (scost[m] < bc) { bc = scost[m]; bm = m; }
            bestMode[p] = bm;
        }
        __syncthreads();
    }
}

cudaError_t launch_intra32_decide(const uint8_t* cur, const uint8_t* refs, uint32_t* cost, int32_t* bestMode, size_t n, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    const size_t cap = (size_t)sm_count() * 4;
    intra32_decide_kernel<<<(unsigned)(n < cap ? n : cap), IDEC_WARPS * 32, 0, st>>>(cur, refs, cost, bestMode, n);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_intra32(const uint8_t* refs, const uint8_t* mode, uint8_t* pred, size_t n, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    const size_t want = (n + INTRA_WARPS - 1) / INTRA_WARPS;
    const size_t cap = (size_t)sm_count() * 8;
    intra32_kernel<<<(unsigned)(want < cap ? want : cap), INTRA_WARPS * 32, 0, st>>>(refs, mode, pred, n);
    count_launch();
    return cudaGetLastError();
}

} // namespace x266
