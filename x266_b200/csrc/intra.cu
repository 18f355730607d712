// intra.cu -- 32x32 intra prediction, 35 modes (0 planar, 1 DC, 2..34 angular).
//
// Spec source: src/mkIntra32-wip.bsv (work in progress in the reference, no C model -> PARITY UNPINNED,
// see DESIGN.md):  reference samples left[64] / top[65] (:34-37), per-row index and fraction tables
// = ((k+1)*angle)>>5 and &31 (:75-112), inverse-angle projection of the side reference for negative
// angles (:151-220), 2-tap interpolation (:352-368), DC (:388-392).
// One CTA (256 threads) per prediction, 4 output pixels per thread, 1 KiB coalesced store per CTA.
#include "common.cuh"
#include "kernels.h"

namespace x266 {

__constant__ int c_intraAngle[35] = { 0, 0, 32, 26, 21, 17, 13, 9, 5, 2, 0, -2, -5, -9, -13, -17, -21, -26,
                                      -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32 };
// 8192/|angle| rounded, for angle = -2,-5,-9,-13,-17,-21,-26,-32
__constant__ int c_intraInv[8] = { 4096, 1638, 910, 630, 482, 390, 315, 256 };

__global__ void __launch_bounds__(256)
intra32_kernel(const uint8_t* __restrict__ refs, const uint8_t* __restrict__ modes, uint8_t* __restrict__ pred, size_t n)
{
    __shared__ int sref[32 + 65 + 3];       // main reference, index -32..64 at sref[32 + i]
    __shared__ uint8_t sraw[132];           // left[64] | top[65]
    __shared__ int sdc;
    const int tid = threadIdx.x;

    for (size_t p = blockIdx.x; p < n; p += gridDim.x) {
        const int mode = modes[p] > 34 ? 1 : modes[p];   // host API rejects > 34; keep device reads in range
        if (tid < 129) sraw[tid] = refs[p * 129 + tid];
        __syncthreads();
        const uint8_t* left = sraw;          // left[i] = pixel (-1, i)
        const uint8_t* top = sraw + 64;      // top[0] = corner, top[1+i] = pixel (i, -1)
        const bool isVer = mode >= 18;
        const int ang = c_intraAngle[mode];

        if (mode >= 2) {
            if (tid <= 64) sref[32 + tid] = isVer ? top[tid] : (tid == 0 ? top[0] : left[tid - 1]);
            if (ang < 0 && tid >= 1 && tid <= 32 && -tid >= ((32 * ang) >> 5)) {
                int inv = 0;
#pragma unroll
                for (int a = 0; a < 8; a++) if (c_intraAngle[11 + a] == ang) inv = c_intraInv[a];
                const int s = (tid * inv + 128) >> 8;
                sref[32 - tid] = isVer ? left[s - 1] : top[s];
            }
        } else if (mode == 1 && tid < 32) {
            int s = left[tid] + top[1 + tid];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (tid == 0) sdc = (s + 32) >> 6;
        }
        __syncthreads();

        const int row = tid >> 3, col0 = (tid & 7) * 4;
        uint32_t packed = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int col = col0 + i;
            int v;
            if (mode == 0) {
                v = ((31 - col) * left[row] + (col + 1) * top[33] + (31 - row) * top[1 + col] + (row + 1) * left[32] + 32) >> 6;
            } else if (mode == 1) {
                v = sdc;
            } else {
                const int xr = isVer ? col : row;      // position along the main reference
                const int yd = isVer ? row : col;      // distance from the main reference
                const int t = (yd + 1) * ang;
                const int idx = t >> 5, f = t & 31;
                const int a = sref[32 + xr + idx + 1];
                const int b = sref[32 + xr + idx + 2];  // in range: index <= 65 only when f == 0 (slot padded)
                v = f ? (((32 - f) * a + f * b + 16) >> 5) : a;
            }
            packed |= (uint32_t)(v & 0xFF) << (8 * i);
        }
        reinterpret_cast<uint32_t*>(pred + p * 1024)[tid] = packed;
        __syncthreads();
    }
}

cudaError_t launch_intra32(const uint8_t* refs, const uint8_t* mode, uint8_t* pred, size_t n, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    size_t cap = (size_t)sm_count() * 8;
    intra32_kernel<<<(unsigned)(n < cap ? n : cap), 256, 0, st>>>(refs, mode, pred, n);
    count_launch();
    return cudaGetLastError();
}

} // namespace x266
