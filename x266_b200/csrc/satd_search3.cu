// satd_search3.cu -- SATD full search v3 (R in {8,16,32}): packed 16-bit transform domain, two positions per thread.
//
// Reference behaviour: cost of one candidate = satd8x8(cur - ref(mv)), src_tb/satd.c:31-118.  The search loop,
// window convention and argmin rule are ours (the reference has no search loop).
//
// Why a third kernel: v2 (satd.cu) is bound by the shared-memory return path -- every candidate re-reads the 64
// 32-bit coefficients of T(cur) (256 B, 16 LDS.128 per warp = 64 SM cycles per 32 candidates) -- and not by HBM or
// the ALUs (ncu: issue slots 54 % busy).  v3 removes three quarters of that traffic and a third of the instructions:
//   * coefficients are packed two per word with a bias that keeps every half non-negative (satd_packed.h), so
//     T(cur) is 128 B and every butterfly of the transform is ONE plain 32-bit add for two coefficients;
//   * a thread owns TWO window positions (p, p+8) that share 8 of their 9 blocks, so one T(cur) word fetched
//     from shared memory serves two candidates: 64 B of shared-memory traffic per candidate instead of 256;
//   * |a-b| = 2 max(a,b) - a - b with sum_k T[k] = 64 x[0][0]: per pair of coefficients one VIMNMX.S16x2 (integer ALU
//     pipe) and one IDP.2A (FMA pipe) instead of two VABSDIFF on the ALU pipe.
// Work decomposition: a unit is (row of 8x8 blocks, tile of 64 horizontal window positions, third of the 2R+1 vertical
// offsets); ONE WARP = one CTA owns a unit: the window rows it touches (22+7 rows x 72 bytes at R=32) and the T(cur) of
// the R/4+8 blocks it can serve live in its 6.6 KB of shared memory, 16 such CTAs are resident per SM, and nothing in
// the kernel waits on a CTA barrier (a first version with one five-warp CTA per tile spent 24 % of its warp time in
// the prologue, the epilogue and its barrier).  Measured on 1080p +-32: 1/2/3/4/5/7 CTAs per (tile, row) = 0.613 /
// 0.575 / 0.577 / 0.63 / 0.606 / 0.66 ms -- fewer, longer CTAs amortise the prologue, more of them fill the last wave.
// Lane (g, e) = (lane>>3, lane&7) owns positions P0+16g+e and P0+16g+8+e; with q = P0/8+2g the blocks of "slot" s are
// i = q-R/4+s for both positions (mx = e+2R-8s and e+2R+8-8s), so the eight lanes of a group read the same T(cur) and
// write 32 contiguous bytes of the cost surface.  Tiles follow window positions, not blocks, so no lane is wasted on
// a strip edge (v2: 38 %).  A block's candidates span up to three tiles and three chunks: the argmin is combined with
// 64-bit atomicMin keys (cost, mvx^2+mvy^2, my, mx) in a stream-ordered scratch buffer and decoded by a second tiny
// kernel.
// Pipes (ncu): the integer ALU pipe of sm_100 issues one warp instruction every two cycles, so the 32 VIMNMX.S16x2 per
// candidate are the floor; the sums therefore go to the FMA pipe (one IDP.2A per word, satd_packed.h maxsum4<1>).
#include "search_tile.cuh"
#include "satd_packed.h"

namespace x266 {

constexpr int S3_TILE = SRCH_TILE;           // window positions per unit
constexpr int S3_WP = 72;                    // window pitch in bytes: 64 positions + 7 halo columns (+1)
constexpr int S3_TCS = 36;                   // T(cur) row: 32 words + 64*cur[0][0] + pad (16-byte rows, distinct banks for 4 blocks)
constexpr int S3_CTAS_PER_SM = 16;

// CT = element type of the cost surface (uint32_t, or uint16_t for the ...U16 entry points: a SATD of 8-bit pixels is <= 32640).
// Measured and not kept (profiles/r02_search_variants.md, source in tools/experiments/satd_search3_rolled_variants.cu.txt): slots 2..R/4
// as one rolled loop body with the running keys in shared memory (12 KB of SASS instead of 27 KB, 121 registers: +0.7 %), the same at
// 20 / 22 resident warps per SM (96 / 80 registers: +0.1 % / -10 %), and the two mx = 2R candidates of a group computed by its 8 lanes
// together (-3 %) -- the loop is bound by instruction issue (1752 warp instructions per vertical offset, 1152 of them the VIMNMX / IDP.2A
// pairs), not by the instruction cache, latency or occupancy.
template <int R, int ACCF, int NCH, typename CT>
__global__ void __launch_bounds__(32, S3_CTAS_PER_SM)
satd8x8_search_v3_kernel(const uint8_t* __restrict__ cur, const uint8_t* __restrict__ refPad, intptr_t strd, int w, int by0,
                         size_t blk0, size_t blk1, CT* __restrict__ cost, unsigned long long* __restrict__ keys)
{
    constexpr int SIDE = 2 * R + 1;
    constexpr int WS = 2 * R + 8;
    constexpr int NSLOT = R / 4 + 2;
    constexpr int NBLK = R / 4 + 8;
    constexpr int CH = (SIDE + NCH - 1) / NCH;      // vertical offsets per CTA
    constexpr int WR = CH + 7;                                   // window rows a chunk touches
    static_assert(WR <= WS, "chunk window");
    __shared__ __align__(16) uint8_t win[WR * S3_WP];
    __shared__ __align__(16) uint8_t curw[8][128];
    __shared__ __align__(16) uint32_t V[S3_WP][4];
    __shared__ __align__(16) uint32_t tcur[NBLK][S3_TCS];

    const int lane = threadIdx.x;
    const int g = lane >> 3, e = lane & 7;
    const int bw = w >> 3;
    const int padW = w + 2 * R;
    const int xa = 16 * g + e;                   // tile-local column of position A; B = A + 8

    const int by8 = by0 + blockIdx.y;
    const int P0 = blockIdx.x * S3_TILE;
    const int my0 = blockIdx.z * CH;
    const int my1 = (my0 + CH) < SIDE ? (my0 + CH) : SIDE;
    const int iBase = P0 / 8 - R / 4;
    const int iq = iBase + 2 * g;                // block of slot 0
    const size_t bRow = (size_t)by8 * bw;
    {
        const int lo = iBase < 0 ? 0 : iBase;
        const int hi = (iBase + NBLK - 1) < (bw - 1) ? (iBase + NBLK - 1) : (bw - 1);
        if (my0 >= SIDE || hi < lo || bRow + hi < blk0 || bRow + lo >= blk1) return;
    }
    unsigned keyA[NSLOT], keyB[NSLOT];           // per (slot, position): cost << 7 | rank(my); mx is fixed per entry
#pragma unroll
    for (int s = 0; s < NSLOT; s++) keyA[s] = keyB[s] = 0xFFFFFFFFu;

    // ---- stage the window rows of this chunk, the current blocks and T(cur)
    {
        const uint8_t* wsrc = refPad + ((intptr_t)by8 * 8 + my0) * strd + P0;
        const int rows = my1 - my0 + 7;
        // aligned planes: asynchronous copies, ONE exposed round trip for the whole prologue instead of one per group of loads (see sad.cu)
        if ((((uintptr_t)wsrc | (uintptr_t)strd) & 3) == 0 && P0 + S3_WP <= padW) {
            const uint32_t winS = (uint32_t)__cvta_generic_to_shared(win);
            for (int idx = lane; idx < rows * (S3_WP / 4); idx += 32) {
                const int yy = idx / (S3_WP / 4), xx = idx - yy * (S3_WP / 4);
                cp_async4(winS + 4u * idx, reinterpret_cast<const uint32_t*>(wsrc + (intptr_t)yy * strd) + xx);
            }
        } else {
            for (int idx = lane; idx < rows * S3_WP; idx += 32) {
                const int yy = idx / S3_WP, xx = idx - yy * S3_WP;
                win[idx] = (P0 + xx < padW) ? wsrc[(intptr_t)yy * strd + xx] : (uint8_t)0;
            }
        }
        if ((((uintptr_t)cur | (uintptr_t)w) & 7) == 0) {        // 8 pixels of a block row per load
            const uint32_t curS = (uint32_t)__cvta_generic_to_shared(&curw[0][0]);
            for (int idx = lane; idx < 8 * 16; idx += 32) {
                const int r = idx >> 4, x8 = idx & 15;
                const int i = iBase + x8;
                const bool in = x8 < NBLK && i >= 0 && i < bw;
                cp_async8_zfill(curS + 128u * r + 8u * x8, cur + (size_t)(by8 * 8 + r) * w + (in ? i : 0) * 8, in ? 8 : 0);
            }
        } else {
            for (int idx = lane; idx < 8 * 128; idx += 32) {
                const int r = idx >> 7, x = idx & 127;
                const int i = iBase + (x >> 3);
                curw[r][x] = (x < NBLK * 8 && i >= 0 && i < bw) ? cur[(size_t)(by8 * 8 + r) * w + i * 8 + (x & 7)] : (uint8_t)0;
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
#pragma unroll 1
        for (int k8 = 0; k8 * 8 < NBLK; k8++) {      // blocks 8*k8 .. 8*k8+7, with the code path of the window
            if (lane < 16) {
                uint32_t px[8], o[4][4];
#pragma unroll
                for (int i = 0; i < 8; i++) px[i] = *reinterpret_cast<const uint32_t*>(&curw[i][64 * k8 + 4 * lane]);
                s3::vertical4(px, o);
#pragma unroll
                for (int c = 0; c < 4; c++) *reinterpret_cast<uint4*>(V[4 * lane + c]) = make_uint4(o[c][0], o[c][1], o[c][2], o[c][3]);
            }
            __syncwarp();
            if (lane < 8 && k8 * 8 + lane < NBLK) {
                uint32_t T[32];
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const uint4 v = *reinterpret_cast<const uint4*>(V[8 * lane + c]);
                    T[c] = v.x; T[8 + c] = v.y; T[16 + c] = v.z; T[24 + c] = v.w;
                }
                s3::horizontal8(T);
                uint32_t* d = tcur[k8 * 8 + lane];
#pragma unroll
                for (int k = 0; k < 8; k++) *reinterpret_cast<uint4*>(d + 4 * k) = make_uint4(T[4 * k], T[4 * k + 1], T[4 * k + 2], T[4 * k + 3]);
                d[32] = 64u * curw[0][64 * k8 + 8 * lane];
            }
            __syncwarp();
        }
    }

    // which of this lane's slots are real blocks inside [blk0, blk1): does not depend on the vertical offset
    uint32_t slotMask = 0;
#pragma unroll
    for (int s = 0; s < NSLOT; s++) {
        const int i = iq + s;
        const size_t b = bRow + i;
        if (i >= 0 && i < bw && b >= blk0 && b < blk1) slotMask |= 1u << s;
    }
    for (int my = my0; my < my1; my++) {
        const uint8_t* wrow = win + (my - my0) * S3_WP;      // window row of this vertical offset
        // ---- one vertical offset: vertical pass of the 72 window columns, then the two positions of this lane
        if (lane < S3_WP / 4) {
            uint32_t px[8], o[4][4];
#pragma unroll
            for (int i = 0; i < 8; i++) px[i] = *reinterpret_cast<const uint32_t*>(wrow + i * S3_WP + 4 * lane);
            s3::vertical4(px, o);
#pragma unroll
            for (int c = 0; c < 4; c++) *reinterpret_cast<uint4*>(V[4 * lane + c]) = make_uint4(o[c][0], o[c][1], o[c][2], o[c][3]);
        }
        __syncwarp();
        uint32_t TA[32], TB[32];
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const uint4 a = *reinterpret_cast<const uint4*>(V[xa + c]);
            const uint4 b = *reinterpret_cast<const uint4*>(V[xa + 8 + c]);
            TA[c] = a.x; TA[8 + c] = a.y; TA[16 + c] = a.z; TA[24 + c] = a.w;
            TB[c] = b.x; TB[8 + c] = b.y; TB[16 + c] = b.z; TB[24 + c] = b.w;
        }
        __syncwarp();                            // V may be overwritten by the next iteration from here on
        s3::horizontal8(TA);
        s3::horizontal8(TB);
        // cost = (2*acc - 64*ref[0][0] - 64*cur[0][0] - 2*BIAS_SUM + 2) >> 2
        const uint32_t subA = 64u * wrow[xa] + 2u * s3::BIAS_SUM - 2u;
        const uint32_t subB = 64u * wrow[xa + 8] + 2u * s3::BIAS_SUM - 2u;
        const int dy = my - R;
        const unsigned rank = srch_rank(dy);
        // cost index of (slot s, position A) = cbase + s * (SIDE*SIDE - 8); position B: + 8
        CT* cbase = cost ? cost + (((ptrdiff_t)bRow + iq - (ptrdiff_t)blk0) * SIDE + my) * SIDE + e + 2 * R : nullptr;

#pragma unroll
        for (int s = 0; s < NSLOT; s++) {
            if ((slotMask >> s) & 1u) {
                const bool doA = (s <= R / 4) && (s >= 1 || e == 0);
                const bool doB = (s >= 1) && (s >= 2 || e == 0);
                const uint32_t* tc = tcur[2 * g + s];
                uint32_t accA = 0, accB = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint4 c = *reinterpret_cast<const uint4*>(tc + 4 * k);
                    if (s <= R / 4) accA = s3::maxsum4<ACCF>(&TA[4 * k], c.x, c.y, c.z, c.w, accA);
                    if (s >= 1) accB = s3::maxsum4<ACCF>(&TB[4 * k], c.x, c.y, c.z, c.w, accB);
                }
                const uint32_t c64 = tc[32];
                // the tail of a candidate on the FMA pipe: x >> 2 as the high word of x * 2^30, key = cost * 128 + rank as a multiply-add;
                // only the running minimum needs the integer ALU pipe (the limiter of this kernel)
                if (doA) {
                    const uint32_t c4 = __umulhi(2u * accA - subA - c64, 1u << 30);
                    if (cost) cbase[s * (SIDE * SIDE - 8)] = (CT)c4;
                    keyA[s] = min(c4 * 128u + rank, keyA[s]);
                }
                if (doB) {
                    const uint32_t c4 = __umulhi(2u * accB - subB - c64, 1u << 30);
                    if (cost) cbase[s * (SIDE * SIDE - 8) + 8] = (CT)c4;
                    keyB[s] = min(c4 * 128u + rank, keyB[s]);
                }
            }
        }
    }
    if (keys) srch_flush_keys<R, NSLOT>(keyA, keyB, e, (ptrdiff_t)bRow + iq - (ptrdiff_t)blk0, keys);
}

__global__ void srch_keys_decode_kernel(const unsigned long long* __restrict__ keys, int32_t* __restrict__ best, size_t n, int R)
{
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const unsigned long long k = keys[b];
    best[3 * b + 0] = (int32_t)(k >> 40);
    best[3 * b + 1] = (int)(k & 0xFFF) - R;
    best[3 * b + 2] = (int)((k >> 12) & 0xFFF) - R;
}

static std::atomic<int> g_accForm{1};
void set_search_acc_form(int f) { g_accForm = f; }

template <int R, int ACCF, typename CT>
static cudaError_t launch_v3f(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, size_t blk0, size_t blk1,
                              CT* cost, int32_t* best, cudaStream_t st)
{
    struct Tag {};
    return srch_launch<R>(satd8x8_search_v3_kernel<R, ACCF, srch_chunks<R>(), CT>, srch_attr_flag<Tag>(), cur, refPad, strd, w, blk0, blk1, cost, best, st);
}

template <int R, typename CT>
static cudaError_t launch_v3(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, size_t blk0, size_t blk1,
                             CT* cost, int32_t* best, cudaStream_t st)
{
    if (sizeof(CT) == 4) {                           // the alternative accumulate forms exist for the u32 surface only (diagnostic)
        if (g_accForm == 0) return launch_v3f<R, 0>(cur, refPad, strd, w, blk0, blk1, (uint32_t*)cost, best, st);
        if (g_accForm == 2) return launch_v3f<R, 2>(cur, refPad, strd, w, blk0, blk1, (uint32_t*)cost, best, st);
    }
    return launch_v3f<R, 1>(cur, refPad, strd, w, blk0, blk1, cost, best, st);
}

template <typename CT>
static cudaError_t launch_v3_any(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, int range,
                                 size_t blk0, size_t blk1, CT* cost, int32_t* best, cudaStream_t st)
{
    if (range == 32) return launch_v3<32>(cur, refPad, strd, w, blk0, blk1, cost, best, st);
    if (range == 16) return launch_v3<16>(cur, refPad, strd, w, blk0, blk1, cost, best, st);
    if (range == 8) return launch_v3<8>(cur, refPad, strd, w, blk0, blk1, cost, best, st);
    return cudaErrorInvalidValue;
}

cudaError_t launch_satd8x8_search_v3(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, int h, int range,
                                     size_t blk0, size_t blk1, uint32_t* cost, int32_t* best, cudaStream_t st)
{
    (void)h;
    return launch_v3_any(cur, refPad, strd, w, range, blk0, blk1, cost, best, st);
}

cudaError_t launch_satd8x8_search_v3(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, int h, int range,
                                     size_t blk0, size_t blk1, uint16_t* cost, int32_t* best, cudaStream_t st)
{
    (void)h;
    return launch_v3_any(cur, refPad, strd, w, range, blk0, blk1, cost, best, st);
}

} // namespace x266
