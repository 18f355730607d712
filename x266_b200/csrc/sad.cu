// sad.cu -- plain SAD ("next" row N4): the integer-pel pre-filter that precedes SATD refinement.
// Reference behaviour: riscv/programs/benchmarks/sad/sad.c:27-38 -- sum over an n x n region of
// abs((int)a[i*n+j] - (int)b[i*n+j]); golden value 344807 for the shipped 64x64 dataset (dataset1.h:423-426).
// Both kernels use the native packed-byte VABSDIFF4.U8.ACC (4 |a-b| + accumulate per instruction).
#include "common.cuh"
#include "kernels.h"

namespace x266 {

// ---- sad(a, b, n): one n*n-byte region pair -> one int --------------------------------------------------
__global__ void __launch_bounds__(256)
sad_region_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t bytes, unsigned* __restrict__ out)
{
    unsigned s = 0;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 3) == 0;
    const size_t words = aligned ? bytes / 4 : 0;
    for (size_t i = tid; i < words; i += nth)
        s = __vsadu4(reinterpret_cast<const unsigned*>(a)[i], reinterpret_cast<const unsigned*>(b)[i]) + s;
    for (size_t i = words * 4 + tid; i < bytes; i += nth) s += (unsigned)abs((int)a[i] - (int)b[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// ---- full search: cost[b][my][mx] = SAD(cur block b, ref block displaced by mv), same conventions as
//      xSatd8x8Search (padded reference, argmin rule).  One CTA per 8x8 block; (8+2R)^2 window in smem as
//      words; thread <-> candidate (flattened my,mx); rows are re-aligned with funnel shifts. ------------
constexpr int SADS_THREADS = 256;

__global__ void __launch_bounds__(SADS_THREADS)
sad8x8_search_kernel(const uint8_t* __restrict__ cur, const uint8_t* __restrict__ refPad, intptr_t strd, int w, int range,
                     size_t blk0, uint32_t* __restrict__ cost, int32_t* __restrict__ best)
{
    extern __shared__ __align__(16) uint32_t swin[];           // [ws][wsw + 1] words
    __shared__ unsigned long long sBest[SADS_THREADS / 32];
    const int side = 2 * range + 1, ws = 2 * range + 8;
    const int wsw = (ws + 3) / 4 + 1;                           // words per row incl. one spill word
    const int tid = threadIdx.x;
    const size_t blk = blk0 + blockIdx.x;
    const int bw = w >> 3;
    const int bx = (int)(blk % bw) * 8, by = (int)(blk / bw) * 8;
    const uint8_t* wsrc = refPad + (intptr_t)by * strd + bx;
    uint8_t* wbytes = reinterpret_cast<uint8_t*>(swin);
    for (int i = tid; i < ws * wsw * 4; i += SADS_THREADS) {
        const int yy = i / (wsw * 4), xx = i - yy * (wsw * 4);
        wbytes[i] = xx < ws ? wsrc[(intptr_t)yy * strd + xx] : (uint8_t)0;
    }
    uint32_t c[16];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint8_t* p = cur + (size_t)(by + r) * w + bx;
        c[2 * r] = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
        c[2 * r + 1] = p[4] | (p[5] << 8) | (p[6] << 16) | ((uint32_t)p[7] << 24);
    }
    __syncthreads();

    unsigned long long bestKey = ~0ull;
    uint32_t* costBlk = cost ? cost + (size_t)blockIdx.x * side * side : nullptr;
    for (int cand = tid; cand < side * side; cand += SADS_THREADS) {
        const int my = cand / side, mx = cand - my * side;
        const int w0 = mx >> 2, sh = (mx & 3) * 8;
        unsigned sa = 0, sb = 0;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint32_t* row = swin + (my + r) * wsw + w0;
            const uint32_t x0 = row[0], x1 = row[1], x2 = row[2];
            sa = __vsadu4(__funnelshift_r(x0, x1, sh), c[2 * r]) + sa;
            sb = __vsadu4(__funnelshift_r(x1, x2, sh), c[2 * r + 1]) + sb;
        }
        const unsigned s = sa + sb;
        if (costBlk) costBlk[cand] = s;
        const int dx = mx - range, dy = my - range;
        const unsigned long long key = ((unsigned long long)s << 40) | ((unsigned long long)(dx * dx + dy * dy) << 24) |
                                       ((unsigned long long)my << 12) | (unsigned long long)mx;
        bestKey = key < bestKey ? key : bestKey;
    }
    if (best) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, bestKey, o);
            bestKey = other < bestKey ? other : bestKey;
        }
        if ((tid & 31) == 0) sBest[tid >> 5] = bestKey;
        __syncthreads();
        if (tid == 0) {
            unsigned long long k = sBest[0];
#pragma unroll
            for (int i = 1; i < SADS_THREADS / 32; i++) k = sBest[i] < k ? sBest[i] : k;
            int32_t* o = best + (size_t)blockIdx.x * 3;
            o[0] = (int32_t)(k >> 40);
            o[1] = (int)(k & 0xFFF) - range;
            o[2] = (int)((k >> 12) & 0xFFF) - range;
        }
    }
}

cudaError_t launch_sad_region(const uint8_t* a, const uint8_t* b, size_t bytes, unsigned* out, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    if (bytes == 0) return cudaSuccess;
    const size_t want = (bytes / 4 + 255) / 256 + 1, cap = (size_t)sm_count() * 8;
    sad_region_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(a, b, bytes, out);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_sad8x8_search(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, int h, int range,
                                 size_t blk0, size_t blk1, uint32_t* cost, int32_t* best, cudaStream_t st)
{
    if (blk1 <= blk0) return cudaSuccess;
    if (range < 0 || range > 2047 || (w & 7) || (h & 7) || blk1 > (size_t)(w / 8) * (h / 8)) return cudaErrorInvalidValue;
    const int ws = 2 * range + 8, wsw = (ws + 3) / 4 + 1;
    const size_t smem = (size_t)ws * wsw * 4;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(sad8x8_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    sad8x8_search_kernel<<<(unsigned)(blk1 - blk0), SADS_THREADS, smem, st>>>(cur, refPad, strd, w, range, blk0, cost, best);
    count_launch();
    return cudaGetLastError();
}

} // namespace x266
