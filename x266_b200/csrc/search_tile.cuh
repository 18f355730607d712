// search_tile.cuh -- scaffolding shared by the two position-tile full-search kernels (satd_search3.cu, sad.cu):
// the unit decomposition, the argmin key fold, the key decode kernel and the host-side launch wrapper.
//
// A unit = (row of 8x8 blocks, tile of 64 horizontal window positions, chunk of the 2R+1 vertical offsets) = one warp = one CTA.
// Lane (g, e) = (lane>>3, lane&7) owns positions P0+16g+e ("A") and P0+16g+8+e ("B"); with q = P0/8+2g the block of slot s is
// i = q-R/4+s for both positions, mx = e+2R-8s (A) and e+2R+8-8s (B).  Slot s serves position A for s <= R/4 (s = 0 only for
// e == 0) and position B for s >= 1 (s = 1 only for e == 0).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace x266 {

constexpr int SRCH_TILE = 64;                // window positions per unit
template <int R> __host__ __device__ constexpr int srch_chunks() { return R >= 32 ? 3 : R >= 16 ? 2 : 1; }   // CTAs per (tile, row): ~22 vertical offsets each

// Argmin of one unit.  keyA/keyB[s] = (cost << 7 | rank(my)) of the best vertical offset seen for (slot s, position A/B); mx is fixed per
// entry.  Order: cost, mvx^2+mvy^2, my, mx.  The 8 lanes of a group hold the 8 values of e for the same block: fold them with shuffles, then
// one 64-bit atomicMin per (group, slot, position) into the frame-wide key array.
template <int R, int NSLOT>
__device__ __forceinline__ void srch_flush_keys(const unsigned (&keyA)[NSLOT], const unsigned (&keyB)[NSLOT], int e, ptrdiff_t blkSlot0,
                                                unsigned long long* __restrict__ keys)
{
#pragma unroll
    for (int s = 0; s < NSLOT; s++) {
#pragma unroll
        for (int ab = 0; ab < 2; ab++) {
            if ((ab == 0 && s > R / 4) || (ab == 1 && s < 1)) continue;
            const unsigned k32 = ab ? keyB[s] : keyA[s];
            unsigned long long key = ~0ull;
            if (k32 != 0xFFFFFFFFu) {
                const unsigned rank = k32 & 127u;
                const int dy = (rank & 1) ? -(int)((rank + 1) >> 1) : (int)(rank >> 1);
                const int mx = e + 2 * R - 8 * s + 8 * ab, dx = mx - R;
                key = ((unsigned long long)(k32 >> 7) << 40) | ((unsigned long long)(dx * dx + dy * dy) << 24) |
                      ((unsigned long long)(dy + R) << 12) | (unsigned long long)mx;
            }
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                key = other < key ? other : key;
            }
            if (e == 0 && key != ~0ull) atomicMin(&keys[blkSlot0 + s], key);
        }
    }
}

// rank of a vertical offset for the tie-break: |dy| first, negative before positive (= smaller my first at equal |dy|)
__device__ __forceinline__ unsigned srch_rank(int dy) { return dy < 0 ? (unsigned)(-2 * dy - 1) : (unsigned)(2 * dy); }

__global__ void srch_keys_decode_kernel(const unsigned long long* __restrict__ keys, int32_t* __restrict__ best, size_t n, int R);

// Launches `kern` (signature of both search kernels) over the units that cover blocks [blk0, blk1), with the stream-ordered key scratch
// and the decode kernel when the argmin is wanted.
template <int R, typename Kern, typename CT>
static cudaError_t srch_launch(Kern kern, std::atomic<bool>& attrSet, const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, size_t blk0, size_t blk1,
                               CT* cost, int32_t* best, cudaStream_t st)
{
    const int bw = w / 8;
    const int y0 = (int)(blk0 / bw), y1 = (int)((blk1 - 1) / bw);
    const int nPos = 8 * (bw - 1) + 2 * R + 1;
    const dim3 grid((nPos + SRCH_TILE - 1) / SRCH_TILE, y1 - y0 + 1, srch_chunks<R>());
    cudaError_t e;
    if (!attrSet) {                                  // 16 single-warp CTAs per SM need the large shared-memory carve-out
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
        attrSet = true;
    }
    const size_t nb = blk1 - blk0;
    unsigned long long* keys = nullptr;
    if (best && (e = scratch_alloc((void**)&keys, nb * sizeof(unsigned long long), st)) != cudaSuccess) return e;
    // from here on every exit path hands the key scratch back to the pool
    if (best) e = cudaMemsetAsync(keys, 0xFF, nb * sizeof(unsigned long long), st);
    if (e == cudaSuccess) {
        kern<<<grid, 32, 0, st>>>(cur, refPad, strd, w, y0, blk0, blk1, cost, keys);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && best) {
        srch_keys_decode_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(keys, best, nb, R);
        count_launch();
        e = cudaGetLastError();
    }
    if (keys) {
        const cudaError_t ef = scratch_free(keys, st);
        if (e == cudaSuccess) e = ef;
    }
    return e;
}

// per-device "attribute already set" flag of one kernel instantiation
template <typename Tag>
static std::atomic<bool>& srch_attr_flag()
{
    static std::atomic<bool> flags[65];              // [64] = devices beyond the table: always re-set
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) { flags[64] = false; return flags[64]; }
    return flags[dev];
}

} // namespace x266
