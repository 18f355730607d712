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
#include <type_traits>

namespace x266 {

constexpr int S3_TILE = SRCH_TILE;           // window positions per unit
constexpr int S3_WP = 72;                    // window pitch in bytes: 64 positions + 7 halo columns (+1)
constexpr int S3_TCS = 36;                   // T(cur) row: 32 words + 64*cur[0][0] + pad (16-byte rows, distinct banks for 4 blocks)
constexpr int S3_CTAS_PER_SM = 16;

// VAR 0: every slot unrolled, running keys in registers (the round-1/2 body).
// VAR 1: slots 2..R/4 (both positions live in all lanes) as ONE rolled loop body with the running keys in shared memory (per-lane words,
//        shared-memory atomic min = LSU pipe) -- the unrolled body is 27 KB of SASS for 16 warps that sit in 16 different places of it
//        (ncu: 8 % of the stall samples are no_inst); the three edge slots stay peeled.  The candidate's constant terms ride in the
//        accumulator's initial value: acc0 = -(64 ref00 + 2 BIAS - 2)/2 - 32 cur00, cost = acc >> 1.
// VAR 2: VAR 1 + the two candidates per group that only lane e == 0 owns (mx = 2R: slot 0 of position A, slot 1 of position B) are
//        computed by the 8 lanes of the group together, 4 words each, instead of by full 32-word passes with 7 of 8 lanes discarded.
// VAR 3 / 4: VAR 1 compiled for 20 / 22 resident warps per SM (96 / 88 registers: the rolled body no longer holds the 20 key registers).
template <int R, int ACCF, int NCH, typename CT, int VAR>
__global__ void __launch_bounds__(32, VAR == 3 ? 20 : VAR == 4 ? 22 : S3_CTAS_PER_SM)
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
    __shared__ unsigned skey[VAR ? 2 * NSLOT : 1][32];           // VAR >= 1: running keys, [slot][lane] (A) and [NSLOT + slot][lane] (B)
    __shared__ __align__(16) uint32_t sedge[VAR == 2 ? 4 : 1][2][36];   // VAR 2: T(ref) of positions A and B of the lanes e == 0

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
    if (VAR) {
#pragma unroll
        for (int s = 0; s < 2 * NSLOT; s++) skey[s][lane] = 0xFFFFFFFFu;
    }

    // ---- stage the window rows of this chunk, the current blocks and T(cur)
    {
        const uint8_t* wsrc = refPad + ((intptr_t)by8 * 8 + my0) * strd + P0;
        const int rows = my1 - my0 + 7;
        if ((((uintptr_t)wsrc | (uintptr_t)strd) & 3) == 0 && P0 + S3_WP <= padW) {
            for (int idx = lane; idx < rows * (S3_WP / 4); idx += 32) {
                const int yy = idx / (S3_WP / 4), xx = idx - yy * (S3_WP / 4);
                reinterpret_cast<uint32_t*>(win)[idx] = __ldg(reinterpret_cast<const uint32_t*>(wsrc + (intptr_t)yy * strd) + xx);
            }
        } else {
            for (int idx = lane; idx < rows * S3_WP; idx += 32) {
                const int yy = idx / S3_WP, xx = idx - yy * S3_WP;
                win[idx] = (P0 + xx < padW) ? wsrc[(intptr_t)yy * strd + xx] : (uint8_t)0;
            }
        }
        if ((((uintptr_t)cur | (uintptr_t)w) & 7) == 0) {        // 8 pixels of a block row per load
            for (int idx = lane; idx < 8 * 16; idx += 32) {
                const int r = idx >> 4, x8 = idx & 15;
                const int i = iBase + x8;
                uint2 v = make_uint2(0u, 0u);
                if (x8 < NBLK && i >= 0 && i < bw) v = __ldg(reinterpret_cast<const uint2*>(cur + (size_t)(by8 * 8 + r) * w + i * 8));
                *reinterpret_cast<uint2*>(&curw[r][8 * x8]) = v;
            }
        } else {
            for (int idx = lane; idx < 8 * 128; idx += 32) {
                const int r = idx >> 7, x = idx & 127;
                const int i = iBase + (x >> 3);
                curw[r][x] = (x < NBLK * 8 && i >= 0 && i < bw) ? cur[(size_t)(by8 * 8 + r) * w + i * 8 + (x & 7)] : (uint8_t)0;
            }
        }
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
                d[32] = VAR ? 0u - 32u * curw[0][64 * k8 + 8 * lane] : 64u * curw[0][64 * k8 + 8 * lane];
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

        if (VAR == 0) {
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
        } else {
            // constant terms of a candidate in the accumulator's initial value (tcur[.][32] = -32 cur00): cost = acc >> 1
            const uint32_t hA = 0u - (32u * wrow[xa] + (uint32_t)s3::BIAS_SUM - 1u);
            const uint32_t hB = 0u - (32u * wrow[xa + 8] + (uint32_t)s3::BIAS_SUM - 1u);
            const uint32_t keyRank = rank;
            // one (slot, position pair): DOA / DOB = which positions every lane of the warp computes
            auto slot = [&](int s, auto doAc, auto doBc, bool wrA, bool wrB) {
                constexpr bool DOA = decltype(doAc)::value, DOB = decltype(doBc)::value;
                const uint32_t* tc = tcur[2 * g + s];
                const uint32_t c32 = tc[32];
                uint32_t accA = hA + c32, accB = hB + c32;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint4 c = *reinterpret_cast<const uint4*>(tc + 4 * k);
                    if (DOA) accA = s3::maxsum4<ACCF>(&TA[4 * k], c.x, c.y, c.z, c.w, accA);
                    if (DOB) accB = s3::maxsum4<ACCF>(&TB[4 * k], c.x, c.y, c.z, c.w, accB);
                }
                if (DOA && wrA) {
                    const uint32_t c4 = __umulhi(accA, 1u << 31);
                    if (cost) cbase[s * (SIDE * SIDE - 8)] = (CT)c4;
                    atomicMin(&skey[s][lane], c4 * 128u + keyRank);
                }
                if (DOB && wrB) {
                    const uint32_t c4 = __umulhi(accB, 1u << 31);
                    if (cost) cbase[s * (SIDE * SIDE - 8) + 8] = (CT)c4;
                    atomicMin(&skey[NSLOT + s][lane], c4 * 128u + keyRank);
                }
            };
            using T_ = std::true_type;
            using F_ = std::false_type;
            if (VAR == 2) {
                // the candidates mx = 2R: (position A of lane (g,0), block 2g) and (position B of lane (g,0), block 2g+1), 4 words per lane
                if (e == 0) {
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        *reinterpret_cast<uint4*>(&sedge[g][0][4 * k]) = make_uint4(TA[4 * k], TA[4 * k + 1], TA[4 * k + 2], TA[4 * k + 3]);
                        *reinterpret_cast<uint4*>(&sedge[g][1][4 * k]) = make_uint4(TB[4 * k], TB[4 * k + 1], TB[4 * k + 2], TB[4 * k + 3]);
                    }
                }
                __syncwarp();
                const uint4 ta = *reinterpret_cast<const uint4*>(&sedge[g][0][4 * e]);
                const uint4 tb = *reinterpret_cast<const uint4*>(&sedge[g][1][4 * e]);
                const uint4 ca = *reinterpret_cast<const uint4*>(&tcur[2 * g][4 * e]);
                const uint4 cb = *reinterpret_cast<const uint4*>(&tcur[2 * g + 1][4 * e]);
                const uint32_t tav[4] = { ta.x, ta.y, ta.z, ta.w }, tbv[4] = { tb.x, tb.y, tb.z, tb.w };
                uint32_t pa = s3::maxsum4<ACCF>(tav, ca.x, ca.y, ca.z, ca.w, 0u);
                uint32_t pb = s3::maxsum4<ACCF>(tbv, cb.x, cb.y, cb.z, cb.w, 0u);
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
                    pa += __shfl_xor_sync(0xffffffffu, pa, o);
                    pb += __shfl_xor_sync(0xffffffffu, pb, o);
                }
                __syncwarp();                        // sedge may be rewritten by the next vertical offset
                if (e == 0) {
                    if (slotMask & 1u) {
                        const uint32_t c4 = __umulhi(pa + hA + tcur[2 * g][32], 1u << 31);
                        if (cost) cbase[0] = (CT)c4;
                        atomicMin(&skey[0][lane], c4 * 128u + keyRank);
                    }
                    if (slotMask & 2u) {
                        const uint32_t c4 = __umulhi(pb + hB + tcur[2 * g + 1][32], 1u << 31);
                        if (cost) cbase[(SIDE * SIDE - 8) + 8] = (CT)c4;
                        atomicMin(&skey[NSLOT + 1][lane], c4 * 128u + keyRank);
                    }
                }
                if (slotMask & 2u) slot(1, T_{}, F_{}, true, false);
            } else {
                if (slotMask & 1u) slot(0, T_{}, F_{}, e == 0, false);
                if (slotMask & 2u) slot(1, T_{}, T_{}, true, e == 0);
            }
#pragma unroll 1
            for (int s = 2; s <= R / 4; s++)
                if ((slotMask >> s) & 1u) slot(s, T_{}, T_{}, true, true);
            if ((slotMask >> (R / 4 + 1)) & 1u) slot(R / 4 + 1, F_{}, T_{}, false, true);
        }
    }
    if (VAR) {
        __syncwarp();
#pragma unroll
        for (int s = 0; s < NSLOT; s++) { keyA[s] = skey[s][lane]; keyB[s] = skey[NSLOT + s][lane]; }
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
static std::atomic<int> g_searchVar{0};
void set_search_variant(int v) { g_searchVar = v; }

template <int R, int ACCF, typename CT>
static cudaError_t launch_v3f(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, size_t blk0, size_t blk1,
                              CT* cost, int32_t* best, cudaStream_t st)
{
    struct Tag0 {};
    struct Tag1 {};
    struct Tag2 {};
    struct Tag3 {};
    struct Tag4 {};
    if (ACCF == 1 && g_searchVar == 1)
        return srch_launch<R>(satd8x8_search_v3_kernel<R, 1, srch_chunks<R>(), CT, 1>, srch_attr_flag<Tag1>(), cur, refPad, strd, w, blk0, blk1, cost, best, st);
    if (ACCF == 1 && g_searchVar == 2)
        return srch_launch<R>(satd8x8_search_v3_kernel<R, 1, srch_chunks<R>(), CT, 2>, srch_attr_flag<Tag2>(), cur, refPad, strd, w, blk0, blk1, cost, best, st);
    if (ACCF == 1 && g_searchVar == 3)
        return srch_launch<R>(satd8x8_search_v3_kernel<R, 1, srch_chunks<R>(), CT, 3>, srch_attr_flag<Tag3>(), cur, refPad, strd, w, blk0, blk1, cost, best, st);
    if (ACCF == 1 && g_searchVar == 4)
        return srch_launch<R>(satd8x8_search_v3_kernel<R, 1, srch_chunks<R>(), CT, 4>, srch_attr_flag<Tag4>(), cur, refPad, strd, w, blk0, blk1, cost, best, st);
    return srch_launch<R>(satd8x8_search_v3_kernel<R, ACCF, srch_chunks<R>(), CT, 0>, srch_attr_flag<Tag0>(), cur, refPad, strd, w, blk0, blk1, cost, best, st);
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
