// frame.cu -- the encoder's frame <-> tile format on the device ("next" row N2).
// Reference behaviour: src/x266.cpp:415-453 (xConvInputFmt: planar YUV 4:2:0 -> raster of 512-byte
// ref_block_t tiles, m_Y[16*16] | m_C[8 rows x (U,V) x 8] | m_I[128]) and src/x266.cpp:455-492
// (xConvOutput420, the inverse).  One thread per 16 output bytes, fully coalesced on the tile side.
#include "common.cuh"
#include "kernels.h"

namespace x266 {

__global__ void __launch_bounds__(256)
conv_input_fmt_kernel(uint8_t* __restrict__ tiles, const uint8_t* __restrict__ Y, const uint8_t* __restrict__ U,
                      const uint8_t* __restrict__ V, intptr_t strdY, int tilesPerRow, size_t nTiles)
{
    // 24 16-byte pieces per tile: 16 luma rows + 8 interleaved chroma rows (m_I is left untouched)
    const intptr_t strdC = strdY >> 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nTiles * 24; i += (size_t)gridDim.x * blockDim.x) {
        const size_t tile = i / 24;
        const int piece = (int)(i - tile * 24);
        const size_t ty = tile / tilesPerRow, tx = tile - ty * tilesPerRow;
        uint8_t* dst = tiles + tile * 512 + piece * 16;
        uint8_t v[16];
        if (piece < 16) {
            const uint8_t* s = Y + (ty * 16 + piece) * strdY + tx * 16;
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = s[k];
        } else {
            const int r = piece - 16;
            const uint8_t* su = U + (ty * 8 + r) * strdC + tx * 8;
            const uint8_t* sv = V + (ty * 8 + r) * strdC + tx * 8;
#pragma unroll
            for (int k = 0; k < 8; k++) { v[2 * k] = su[k]; v[2 * k + 1] = sv[k]; }
        }
        uint4 o;
        o.x = v[0] | (v[1] << 8) | (v[2] << 16) | ((uint32_t)v[3] << 24);
        o.y = v[4] | (v[5] << 8) | (v[6] << 16) | ((uint32_t)v[7] << 24);
        o.z = v[8] | (v[9] << 8) | (v[10] << 16) | ((uint32_t)v[11] << 24);
        o.w = v[12] | (v[13] << 8) | (v[14] << 16) | ((uint32_t)v[15] << 24);
        *reinterpret_cast<uint4*>(dst) = o;
    }
}

__global__ void __launch_bounds__(256)
conv_output420_kernel(const uint8_t* __restrict__ tiles, uint8_t* __restrict__ Y, intptr_t strdY, uint8_t* __restrict__ U,
                      uint8_t* __restrict__ V, intptr_t strdC, int tilesPerRow, size_t nTiles)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nTiles * 24; i += (size_t)gridDim.x * blockDim.x) {
        const size_t tile = i / 24;
        const int piece = (int)(i - tile * 24);
        const size_t ty = tile / tilesPerRow, tx = tile - ty * tilesPerRow;
        const uint4 w = *reinterpret_cast<const uint4*>(tiles + tile * 512 + piece * 16);
        const uint32_t ww[4] = { w.x, w.y, w.z, w.w };
        if (piece < 16) {
            uint8_t* d = Y + (ty * 16 + piece) * strdY + tx * 16;
#pragma unroll
            for (int k = 0; k < 16; k++) d[k] = (uint8_t)(ww[k >> 2] >> (8 * (k & 3)));
        } else {
            const int r = piece - 16;
            uint8_t* du = U + (ty * 8 + r) * strdC + tx * 8;
            uint8_t* dv = V + (ty * 8 + r) * strdC + tx * 8;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                du[k] = (uint8_t)(ww[k >> 1] >> (16 * (k & 1)));
                dv[k] = (uint8_t)(ww[k >> 1] >> (16 * (k & 1) + 8));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Stand-alone transpose stage: src/mkTranspose.bsv:95-99 (mkTranspose32x32 on Bit#(8)) is a
// streaming 32x32 byte corner-turn.  One warp per 1 KiB tile: coalesced 128-bit loads, a 32x33-word-free
// byte transpose through a padded shared-memory tile, coalesced 128-bit stores.  (Inside the DCT kernels the
// transpose does not exist as a separate step: see dct_imma.cu.)
// ------------------------------------------------------------------------------------------------
constexpr int TR_WARPS = 8;

__global__ void __launch_bounds__(TR_WARPS * 32)
transpose32_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t nTiles)
{
    __shared__ uint8_t tile[TR_WARPS][32][36];          // 36-byte rows: word aligned, 9-word stride spreads the banks
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (size_t t = (size_t)blockIdx.x * TR_WARPS + warp; t < nTiles; t += (size_t)gridDim.x * TR_WARPS) {
        const uint4* s = reinterpret_cast<const uint4*>(src + t * 1024);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int gi = i * 32 + lane, row = gi >> 1, half = gi & 1;      // 16-byte piece `half` of row `row`
            const uint4 v = s[gi];
            uint32_t* d = reinterpret_cast<uint32_t*>(&tile[warp][row][half * 16]);
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
        __syncwarp();
        uint4* o = reinterpret_cast<uint4*>(dst + t * 1024);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int gi = i * 32 + lane, row = gi >> 1, half = gi & 1;      // output row `row` = input column `row`
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int c = half * 16 + 4 * k;
                w[k] = tile[warp][c][row] | (tile[warp][c + 1][row] << 8) | (tile[warp][c + 2][row] << 16) | ((uint32_t)tile[warp][c + 3][row] << 24);
            }
            o[gi] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Luma of a ref_block_t frame (src/x266.cpp:56-63, tiles of 16x16) as a plane, edge-replicated by `pad` pixels on every side: what the
// full-search kernels read.  out has stride w + 2 pad and h + 2 pad rows; pad = 0 gives the plain plane.  One thread per 4 output pixels.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tiles_to_luma_kernel(const uint8_t* __restrict__ tiles, int tilesPerRow, int w, int h, int pad, uint8_t* __restrict__ out)
{
    const int ow = w + 2 * pad, oh = h + 2 * pad;
    const int wordsPerRow = (ow + 3) >> 2;
    const size_t total = (size_t)wordsPerRow * oh;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int oy = (int)(i / wordsPerRow), ox0 = (int)(i - (size_t)oy * wordsPerRow) * 4;
        const int y = min(max(oy - pad, 0), h - 1);
        uint8_t v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int x = min(max(ox0 + k - pad, 0), w - 1);
            v[k] = tiles[((size_t)(y >> 4) * tilesPerRow + (x >> 4)) * 512 + (y & 15) * 16 + (x & 15)];
        }
        uint8_t* d = out + (size_t)oy * ow + ox0;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (ox0 + k < ow) d[k] = v[k];
    }
}

cudaError_t launch_tiles_to_luma(const uint8_t* tiles, int width, int height, int pad, uint8_t* out, cudaStream_t st)
{
    if (width <= 0 || height <= 0 || (width & 15) || (height & 15) || pad < 0) return cudaErrorInvalidValue;
    const size_t total = (size_t)((width + 2 * pad + 3) / 4) * (height + 2 * pad);
    const size_t want = (total + 255) / 256, cap = (size_t)sm_count() * 8;
    tiles_to_luma_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(tiles, width / 16, width, height, pad, out);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_transpose32(const uint8_t* src, uint8_t* dst, size_t nTiles, cudaStream_t st)
{
    if (nTiles == 0) return cudaSuccess;
    const size_t want = (nTiles + TR_WARPS - 1) / TR_WARPS, cap = (size_t)sm_count() * 8;
    transpose32_kernel<<<(unsigned)(want < cap ? want : cap), TR_WARPS * 32, 0, st>>>(src, dst, nTiles);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_conv_input_fmt(uint8_t* tiles, const uint8_t* Y, const uint8_t* U, const uint8_t* V, intptr_t strdY,
                                  int width, int height, cudaStream_t st)
{
    if (width <= 0 || height <= 0 || (width & 15) || (height & 15)) return cudaErrorInvalidValue;
    const size_t nTiles = (size_t)(width / 16) * (height / 16);
    const size_t want = (nTiles * 24 + 255) / 256, cap = (size_t)sm_count() * 8;
    conv_input_fmt_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(tiles, Y, U, V, strdY, width / 16, nTiles);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_conv_output420(const uint8_t* tiles, uint8_t* Y, intptr_t strdY, uint8_t* U, uint8_t* V, intptr_t strdC,
                                  int width, int height, cudaStream_t st)
{
    if (width <= 0 || height <= 0 || (width & 15) || (height & 15)) return cudaErrorInvalidValue;
    const size_t nTiles = (size_t)(width / 16) * (height / 16);
    const size_t want = (nTiles * 24 + 255) / 256, cap = (size_t)sm_count() * 8;
    conv_output420_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(tiles, Y, strdY, U, V, strdC, width / 16, nTiles);
    count_launch();
    return cudaGetLastError();
}

} // namespace x266
