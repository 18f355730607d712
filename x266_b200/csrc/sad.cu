// sad.cu -- plain SAD ("next" row N4): the integer-pel pre-filter that precedes SATD refinement.
// Reference behaviour: riscv/programs/benchmarks/sad/sad.c:27-38 -- sum over an n x n region of
// abs((int)a[i*n+j] - (int)b[i*n+j]); golden value 344807 for the shipped 64x64 dataset (dataset1.h:423-426).
// Both kernels use the native packed-byte VABSDIFF4.U8.ACC (4 |a-b| + accumulate per instruction).
#include "search_tile.cuh"

namespace x266 {

// acc + sum of the four byte |a-b|: ONE VABSDIFF4.U8.ACC.  Written as PTX because `__vsadu4(a, b) + acc` reaches ptxas as an
// un-accumulated VABSDIFF4 plus a separate add, which it then gathers into 3-input IADD3s: +8 integer-ALU instructions per candidate
// on the pipe that bounds the search (SASS of the round-2 kernel: 288 VABSDIFF4 + 128 IADD3 per vertical offset; now 288 + 2).
__device__ __forceinline__ unsigned sad4_acc(uint32_t a, uint32_t b, unsigned acc)
{
    unsigned d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}

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

template <typename CT>
__global__ void __launch_bounds__(SADS_THREADS)
sad8x8_search_kernel(const uint8_t* __restrict__ cur, const uint8_t* __restrict__ refPad, intptr_t strd, int w, int range,
                     size_t blk0, CT* __restrict__ cost, int32_t* __restrict__ best)
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
    CT* costBlk = cost ? cost + (size_t)blockIdx.x * side * side : nullptr;
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
        if (costBlk) costBlk[cand] = (CT)s;
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

// ---- full search v2 (R in {8,16,32}): the decomposition of the SATD search v3 (satd_search3.cu) without the
//      transform.  A unit = (row of blocks, tile of 64 window positions, third of the vertical offsets) = one warp = one
//      CTA.  Lane (g, e) keeps the 8x8 reference windows of positions P0+16g+e and P0+16g+8+e in registers (16 + 16 words,
//      re-aligned once per vertical offset with funnel shifts) and walks the <= R/4+2 blocks those positions can serve;
//      the current block (64 B) is read from shared memory once for both positions: 4 LDS.128 + 32 VABSDIFF4.U8.ACC per
//      two candidates (v1: 24 LDS.32 + 16 SHF + 16 VABSDIFF4 per candidate, LSU bound).  Argmin as in v3: per (slot,
//      position) running key in registers, 8-lane shuffle fold, 64-bit atomicMin into a stream-ordered scratch buffer. ----
constexpr int SAD2_TILE = SRCH_TILE;
constexpr int SAD2_WW = 18;                  // window pitch in words (72 bytes)
constexpr int SAD2_CS = 20;                  // words per current block in smem: 16 + 4 pad (four blocks of a warp on distinct banks)

template <int R, typename CT>
__global__ void __launch_bounds__(32, 16)
sad8x8_search_v2_kernel(const uint8_t* __restrict__ cur, const uint8_t* __restrict__ refPad, intptr_t strd, int w, int by0,
                        size_t blk0, size_t blk1, CT* __restrict__ cost, unsigned long long* __restrict__ keys)
{
    constexpr int SIDE = 2 * R + 1;
    constexpr int NSLOT = R / 4 + 2;
    constexpr int NBLK = R / 4 + 8;
    constexpr int CH = (SIDE + srch_chunks<R>() - 1) / srch_chunks<R>();
    constexpr int WR = CH + 7;
    __shared__ __align__(16) uint32_t win[WR * SAD2_WW];
    __shared__ __align__(16) uint32_t curs[NBLK * SAD2_CS];

    const int lane = threadIdx.x;
    const int g = lane >> 3, e = lane & 7;
    const int bw = w >> 3;
    const int padW = w + 2 * R;
    const int xa = 16 * g + e;
    const int by8 = by0 + blockIdx.y;
    const int P0 = blockIdx.x * SAD2_TILE;
    const int my0 = blockIdx.z * CH;
    const int my1 = (my0 + CH) < SIDE ? (my0 + CH) : SIDE;
    const int iBase = P0 / 8 - R / 4;
    const int iq = iBase + 2 * g;
    const size_t bRow = (size_t)by8 * bw;
    {
        const int lo = iBase < 0 ? 0 : iBase;
        const int hi = (iBase + NBLK - 1) < (bw - 1) ? (iBase + NBLK - 1) : (bw - 1);
        if (my0 >= SIDE || hi < lo || bRow + hi < blk0 || bRow + lo >= blk1) return;
    }
    unsigned keyA[NSLOT], keyB[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; s++) keyA[s] = keyB[s] = 0xFFFFFFFFu;

    {   // stage the window rows of this chunk and the current blocks.  Aligned planes go through asynchronous copies: a single-warp CTA
        // has nothing to overlap its own prologue with, and with load -> STS loops every group of loads is one exposed round trip to L2 /
        // HBM -- under the cost surface's write traffic those were 22 % of the warps' time (ncu: long_scoreboard on the prologue's STS).
        const uint8_t* wsrc = refPad + ((intptr_t)by8 * 8 + my0) * strd + P0;
        const int rows = my1 - my0 + 7;
        if ((((uintptr_t)wsrc | (uintptr_t)strd) & 3) == 0 && P0 + 4 * SAD2_WW <= padW) {
            const uint32_t winS = (uint32_t)__cvta_generic_to_shared(win);
            for (int idx = lane; idx < rows * SAD2_WW; idx += 32) {
                const int yy = idx / SAD2_WW, xx = idx - yy * SAD2_WW;
                cp_async4(winS + 4u * idx, reinterpret_cast<const uint32_t*>(wsrc + (intptr_t)yy * strd) + xx);
            }
        } else {
            uint8_t* wb = reinterpret_cast<uint8_t*>(win);
            for (int idx = lane; idx < rows * SAD2_WW * 4; idx += 32) {
                const int yy = idx / (SAD2_WW * 4), xx = idx - yy * (SAD2_WW * 4);
                wb[idx] = (P0 + xx < padW) ? wsrc[(intptr_t)yy * strd + xx] : (uint8_t)0;
            }
        }
        if ((((uintptr_t)cur | (uintptr_t)w) & 7) == 0) {
            const uint32_t curS = (uint32_t)__cvta_generic_to_shared(curs);
            for (int idx = lane; idx < 8 * NBLK; idx += 32) {
                const int r = idx / NBLK, t = idx - r * NBLK;
                const int i = iBase + t;
                const bool in = i >= 0 && i < bw;
                cp_async8_zfill(curS + 4u * (t * SAD2_CS + 2 * r), cur + (size_t)(by8 * 8 + r) * w + (in ? i : 0) * 8, in ? 8 : 0);
            }
        } else {
            uint8_t* cb = reinterpret_cast<uint8_t*>(curs);
            for (int idx = lane; idx < 64 * NBLK; idx += 32) {
                const int t = idx >> 6, r = (idx >> 3) & 7, c = idx & 7;
                const int i = iBase + t;
                cb[t * SAD2_CS * 4 + r * 8 + c] = (i >= 0 && i < bw) ? cur[(size_t)(by8 * 8 + r) * w + i * 8 + c] : (uint8_t)0;
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
    }

    const int sh = (xa & 3) * 8;
    uint32_t slotMask = 0;                       // this lane's slots that are real blocks inside [blk0, blk1): independent of the vertical offset
#pragma unroll
    for (int s = 0; s < NSLOT; s++) {
        const int i = iq + s;
        const size_t b = bRow + i;
        if (i >= 0 && i < bw && b >= blk0 && b < blk1) slotMask |= 1u << s;
    }
    for (int my = my0; my < my1; my++) {
        uint32_t A[16], B[16];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint32_t* rw = win + (my - my0 + r) * SAD2_WW + (xa >> 2);
            const uint32_t x0 = rw[0], x1 = rw[1], x2 = rw[2], x3 = rw[3], x4 = rw[4];
            A[2 * r] = __funnelshift_r(x0, x1, sh); A[2 * r + 1] = __funnelshift_r(x1, x2, sh);
            B[2 * r] = __funnelshift_r(x2, x3, sh); B[2 * r + 1] = __funnelshift_r(x3, x4, sh);
        }
        const int dy = my - R;
        const unsigned rank = srch_rank(dy);
        CT* cbase = cost ? cost + (((ptrdiff_t)bRow + iq - (ptrdiff_t)blk0) * SIDE + my) * SIDE + e + 2 * R : nullptr;
#pragma unroll
        for (int s = 0; s < NSLOT; s++) {
            if ((slotMask >> s) & 1u) {
                const bool doA = (s <= R / 4) && (s >= 1 || e == 0);
                const bool doB = (s >= 1) && (s >= 2 || e == 0);
                const uint4* cp = reinterpret_cast<const uint4*>(&curs[(2 * g + s) * SAD2_CS]);
                unsigned a0 = 0, a1 = 0, b0 = 0, b1 = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint4 c = cp[k];
                    if (s <= R / 4) {
                        a0 = sad4_acc(A[4 * k], c.x, a0); a1 = sad4_acc(A[4 * k + 1], c.y, a1);
                        a0 = sad4_acc(A[4 * k + 2], c.z, a0); a1 = sad4_acc(A[4 * k + 3], c.w, a1);
                    }
                    if (s >= 1) {
                        b0 = sad4_acc(B[4 * k], c.x, b0); b1 = sad4_acc(B[4 * k + 1], c.y, b1);
                        b0 = sad4_acc(B[4 * k + 2], c.z, b0); b1 = sad4_acc(B[4 * k + 3], c.w, b1);
                    }
                }
                if (doA) {
                    const unsigned v = a0 + a1;
                    if (cost) cbase[s * (SIDE * SIDE - 8)] = (CT)v;
                    keyA[s] = min(v * 128u + rank, keyA[s]);          // key as a multiply-add (FMA pipe); the ALU pipe is this kernel's limiter
                }
                if (doB) {
                    const unsigned v = b0 + b1;
                    if (cost) cbase[s * (SIDE * SIDE - 8) + 8] = (CT)v;
                    keyB[s] = min(v * 128u + rank, keyB[s]);
                }
            }
        }
    }

    if (keys) srch_flush_keys<R, NSLOT>(keyA, keyB, e, (ptrdiff_t)bRow + iq - (ptrdiff_t)blk0, keys);
}

template <int R, typename CT>
static cudaError_t launch_sad_v2(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, size_t blk0, size_t blk1,
                                 CT* cost, int32_t* best, cudaStream_t st)
{
    struct Tag {};
    return srch_launch<R>(sad8x8_search_v2_kernel<R, CT>, srch_attr_flag<Tag>(), cur, refPad, strd, w, blk0, blk1, cost, best, st);
}

static std::atomic<int> g_sadSearchV1{0};
void set_sad_search_v1(int on) { g_sadSearchV1 = on; }

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

template <typename CT>
static cudaError_t launch_sad8x8_search_as(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, int h, int range,
                                           size_t blk0, size_t blk1, CT* cost, int32_t* best, cudaStream_t st)
{
    if (blk1 <= blk0) return cudaSuccess;
    if (range < 0 || range > 2047 || (w & 7) || (h & 7) || blk1 > (size_t)(w / 8) * (h / 8)) return cudaErrorInvalidValue;
    if (!g_sadSearchV1) {
        if (range == 32) return launch_sad_v2<32>(cur, refPad, strd, w, blk0, blk1, cost, best, st);
        if (range == 16) return launch_sad_v2<16>(cur, refPad, strd, w, blk0, blk1, cost, best, st);
        if (range == 8) return launch_sad_v2<8>(cur, refPad, strd, w, blk0, blk1, cost, best, st);
    }
    const int ws = 2 * range + 8, wsw = (ws + 3) / 4 + 1;
    const size_t smem = (size_t)ws * wsw * 4;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(sad8x8_search_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    sad8x8_search_kernel<CT><<<(unsigned)(blk1 - blk0), SADS_THREADS, smem, st>>>(cur, refPad, strd, w, range, blk0, cost, best);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_sad8x8_search(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, int h, int range,
                                 size_t blk0, size_t blk1, uint32_t* cost, int32_t* best, cudaStream_t st)
{
    return launch_sad8x8_search_as(cur, refPad, strd, w, h, range, blk0, blk1, cost, best, st);
}

cudaError_t launch_sad8x8_search(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, int h, int range,
                                 size_t blk0, size_t blk1, uint16_t* cost, int32_t* best, cudaStream_t st)
{
    return launch_sad8x8_search_as(cur, refPad, strd, w, h, range, blk0, blk1, cost, best, st);
}

} // namespace x266
