// satd.cu -- 8x8 Hadamard SATD: batched cost of precomputed differences, and full-search cost surface.
//
// Reference behaviour: src_tb/satd.c:31-118 (satd8x8): rows then columns, 3 butterfly stages each with
// partner distance 4,2,1 (satd.c:41-66, :73-100), every stage stored to int16 (satd.c:35), cost =
// (sum |coef| + 2) >> 2 (satd.c:105-113).  Hardware schedule: src/mkSatd.bsv:44-176.
//
// int16 wrap: add/sub commute with reduction mod 2^16, so truncating after every stage (reference) is
// the same as computing in 32 bits and truncating once before the abs -- which is what we do.  The
// abs itself is taken on the sign-extended value in 32 bits (abs(-32768) = 32768, as in C).
#include "common.cuh"
#include "kernels.h"

namespace x266 {

// in-place 8-point Hadamard, partner distances 4,2,1 (the output order is a permutation of the
// reference's sums-then-differences order; only the multiset matters for sum|.|, and the search
// kernel uses the same order on both operands).
template <int STRIDE>
__device__ __forceinline__ void had8(int* v)
{
#pragma unroll
    for (int dist = 4; dist >= 1; dist >>= 1) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (!(i & dist)) {
                const int a = v[i * STRIDE], b = v[(i + dist) * STRIDE];
                v[i * STRIDE] = a + b;
                v[(i + dist) * STRIDE] = a - b;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Batch: n candidates of 64 int16 (128 B) -> n int32.  One candidate per thread.  Each warp owns a ring
// of STAGES x 4 KiB shared-memory slots (32 candidates) filled by 1-D TMA bulk copies, so the HBM
// stream is decoupled from the ~580 integer ops a candidate costs.
// Bank conflicts: lane L owns bytes [128L, 128L+128) of the slot, so a plain "row r" read would be an
// 8-way conflict.  Instead lane L reads row (r ^ (L&7)) as its r-th row.  XOR-permuting the rows of
// the 8x8 block by c only flips the sign of Hadamard output row k by (-1)^popcount(k&c), which the
// abs() removes (and -x wraps to the same |x| in int16), so the cost is unchanged and the reads are
// conflict free with a linear (un-swizzled) TMA destination.
// ------------------------------------------------------------------------------------------------
constexpr int SATD_WARPS = 4;
constexpr int SATD_STAGES = 3;
constexpr int SATD_SMEM = SATD_WARPS * SATD_STAGES * 4096 + SATD_WARPS * SATD_STAGES * 8;

__global__ void __launch_bounds__(SATD_WARPS * 32, 4)
satd8x8_batch_kernel(const int16_t* __restrict__ diff, int32_t* __restrict__ out, size_t n)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ring = smem_u32(smem) + warp * (SATD_STAGES * 4096);
    const uint32_t bars = smem_u32(smem) + SATD_WARPS * SATD_STAGES * 4096 + warp * (SATD_STAGES * 8);
    const size_t nGroups = (n + 31) / 32;
    const size_t first = (size_t)blockIdx.x * SATD_WARPS + warp;
    const size_t stride = (size_t)gridDim.x * SATD_WARPS;
    const uint64_t policy = policy_evict_first();

    auto issue = [&](size_t grp, int s) {
        const size_t c0 = grp * 32;
        const uint32_t bytes = (uint32_t)(((n - c0) < 32 ? (n - c0) : 32) * 128);
        mbar_arrive_expect_tx(bars + 8 * s, bytes);
        bulk_g2s(ring + s * 4096, diff + c0 * 64, bytes, bars + 8 * s, policy);
    };

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < SATD_STAGES; s++) mbar_init(bars + 8 * s, 1);
        fence_mbar_init();
        fence_proxy_async();
#pragma unroll
        for (int s = 0; s < SATD_STAGES; s++)
            if (first + (size_t)s * stride < nGroups) issue(first + (size_t)s * stride, s);
    }
    __syncwarp();

    int stage = 0;
    uint32_t parity = 0;
    for (size_t grp = first; grp < nGroups; grp += stride) {
        const size_t c0 = grp * 32;
        const int valid = (int)((n - c0) < 32 ? (n - c0) : 32);
        mbar_wait(bars + 8 * stage, parity);

        int d[64];
        const uint32_t mine = ring + stage * 4096 + lane * 128;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint4 v = ld_shared_v4(mine + ((r ^ (lane & 7)) << 4));
            d[8 * r + 0] = (int)(short)(v.x & 0xFFFF); d[8 * r + 1] = (int)v.x >> 16;
            d[8 * r + 2] = (int)(short)(v.y & 0xFFFF); d[8 * r + 3] = (int)v.y >> 16;
            d[8 * r + 4] = (int)(short)(v.z & 0xFFFF); d[8 * r + 5] = (int)v.z >> 16;
            d[8 * r + 6] = (int)(short)(v.w & 0xFFFF); d[8 * r + 7] = (int)v.w >> 16;
        }
        __syncwarp();
        if (lane == 0) {
            const size_t ng = grp + (size_t)SATD_STAGES * stride;
            if (ng < nGroups) {
                fence_proxy_async();
                issue(ng, stage);
            }
        }
#pragma unroll
        for (int r = 0; r < 8; r++) had8<1>(&d[8 * r]);       // horizontal
        int sad = 0;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            had8<8>(&d[c]);                                   // vertical (rows XOR-permuted, see above)
#pragma unroll
            for (int r = 0; r < 8; r++) sad += abs((int)(short)d[8 * r + c]);
        }
        if (lane < valid) out[c0 + lane] = (sad + 2) >> 2;
        if (++stage == SATD_STAGES) { stage = 0; parity ^= 1; }
    }
}

// ------------------------------------------------------------------------------------------------
// Batch on the int8 tensor cores.  The 2-D Hadamard of an 8x8 block is the 64x64
// Sylvester matrix H64[n][p] = (-1)^popcount(n&p) applied to the 64 samples, so for 16 candidates at a
// time   D[cand][n] = sum_p diff[cand][p] * H64[n][p]   is an m16 x n64 x k64 integer product.  As in the
// DCT kernel the int16 samples are split into byte planes (lo u8, hi s8; +-1 fits s8):
//     D = D_lo + 256 * D_hi     (exact in s32; only D mod 2^16 matters = the reference's int16 wrap)
// That full product is 32 x mma.sync.m16n8k32 per 16 candidates (2 k-steps x 8 n-tiles x 2 planes; round 1 measured it at 0.81 of the
// roofline); the shipped kernel below halves it with H64 = H2 (x) H32.  The candidate rows are the A operand (row-major A == the
// 128-byte candidate); the K order is permuted so that each lane's A registers come from one 16-byte chunk (the +-1 matrix columns are
// permuted to match).  Epilogue per lane: coefficients -> IMAD (lo + 256 hi), pack to int16 (= the reference's wrap), packed |x|,
// IDP.2A accumulate; a quad shuffle reduction gives the two costs of rows g and g+8.
// ------------------------------------------------------------------------------------------------
constexpr int SATDI_WARPS = 8;

// ------------------------------------------------------------------------------------------------
// Full search.  One CTA per 8x8 current block; the (8+2R)^2 reference window lives in shared memory.
// Linearity (exact here: 8-bit pixels give |coef| <= 16320, no int16 wrap is reachable) lets the
// transform of the current block be computed once, and the vertical half of each reference-window
// transform be shared by the 8 horizontally overlapping candidates of a row:
//     T(cur - ref) = T(cur) - H_h( V[my][mx..mx+7] ),   V[my][x] = H_v(window[my..my+7][x]).
// Per candidate: 8 x (8 LDS + 24 add/sub + 8 sad) instead of the direct 64 sub + 384 + 64 abs + 64 add.
// ------------------------------------------------------------------------------------------------
constexpr int SRCH_WARPS = 8;

template <typename CT>
__global__ void __launch_bounds__(SRCH_WARPS * 32)
satd8x8_search_kernel(const uint8_t* __restrict__ cur, const uint8_t* __restrict__ refPad, intptr_t strd,
                      int w, int range, size_t blk0, CT* __restrict__ cost, int32_t* __restrict__ best)
{
    extern __shared__ __align__(16) uint8_t dynsm[];
    const int side = 2 * range + 1;
    const int ws = 2 * range + 8;                  // window width/height (multiple of 4 when range even; padded below)
    const int wsp = (ws + 3) & ~3;
    int* tcur = reinterpret_cast<int*>(dynsm);                         // [64] transform of the current block
    int* tmp = tcur + 64;                                              // [64]
    int* vAll = tmp + 64;                                              // [SRCH_WARPS][8][wsp]
    uint8_t* win = reinterpret_cast<uint8_t*>(vAll + SRCH_WARPS * 8 * wsp);   // [ws][wsp]
    __shared__ unsigned long long sBest[SRCH_WARPS];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t blk = blk0 + blockIdx.x;
    const int bw = w >> 3;
    const int bx = (int)(blk % bw) * 8, by = (int)(blk / bw) * 8;

    // window origin: unpadded pixel (bx-R, by-R) == padded (bx, by)
    const uint8_t* wsrc = refPad + (intptr_t)by * strd + bx;
    for (int i = tid; i < ws * ws; i += SRCH_WARPS * 32) {
        const int yy = i / ws, xx = i - yy * ws;
        win[yy * wsp + xx] = wsrc[(intptr_t)yy * strd + xx];
    }
    if (tid < 8) {                                  // horizontal transform of current row tid
        int v[8];
#pragma unroll
        for (int c = 0; c < 8; c++) v[c] = cur[(size_t)(by + tid) * w + bx + c];
        had8<1>(v);
#pragma unroll
        for (int c = 0; c < 8; c++) tmp[tid * 8 + c] = v[c];
    }
    __syncthreads();
    if (tid < 8) {                                  // vertical transform of column tid
        int v[8];
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = tmp[r * 8 + tid];
        had8<1>(v);
#pragma unroll
        for (int r = 0; r < 8; r++) tcur[r * 8 + tid] = v[r];
    }
    __syncthreads();
    int tc[64];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int4 t = reinterpret_cast<const int4*>(tcur)[i];
        tc[4 * i + 0] = t.x; tc[4 * i + 1] = t.y; tc[4 * i + 2] = t.z; tc[4 * i + 3] = t.w;
    }

    int* V = vAll + warp * 8 * wsp;
    unsigned long long bestKey = ~0ull;
    CT* costBlk = cost ? cost + (size_t)blockIdx.x * side * side : nullptr;

    for (int my = warp; my < side; my += SRCH_WARPS) {
        // vertical transforms of the 8-row band starting at window row my
        for (int x = lane; x < ws; x += 32) {
            int v[8];
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = win[(my + i) * wsp + x];
            had8<1>(v);
#pragma unroll
            for (int r = 0; r < 8; r++) V[r * wsp + x] = v[r];
        }
        __syncwarp();
        for (int mx = lane; mx < side; mx += 32) {
            unsigned sa = 0, sb = 0;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                int v[8];
#pragma unroll
                for (int c = 0; c < 8; c++) v[c] = V[r * wsp + mx + c];
                had8<1>(v);
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    sa = __sad(v[c], tc[r * 8 + c], sa);
                    sb = __sad(v[c + 1], tc[r * 8 + c + 1], sb);
                }
            }
            const unsigned c4 = (sa + sb + 2) >> 2;
            if (costBlk) costBlk[my * side + mx] = (CT)c4;
            const int dx = mx - range, dy = my - range;
            const unsigned long long key = ((unsigned long long)c4 << 40) | ((unsigned long long)(dx * dx + dy * dy) << 24) |
                                           ((unsigned long long)my << 12) | (unsigned long long)mx;
            bestKey = key < bestKey ? key : bestKey;
        }
        __syncwarp();
    }

    if (best) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, bestKey, o);
            bestKey = other < bestKey ? other : bestKey;
        }
        if (lane == 0) sBest[warp] = bestKey;
        __syncthreads();
        if (tid == 0) {
            unsigned long long k = sBest[0];
#pragma unroll
            for (int i = 1; i < SRCH_WARPS; i++) k = sBest[i] < k ? sBest[i] : k;
            int32_t* o = best + (size_t)blockIdx.x * 3;
            o[0] = (int32_t)(k >> 40);
            o[1] = (int)(k & 0xFFF) - range;
            o[2] = (int)((k >> 12) & 0xFFF) - range;
        }
    }
}

static std::atomic<int> g_searchV1{0};      // 0: v3 (satd_search3.cu) for R in {8, 16, 32}; 1: the any-range kernel (one CTA per block) for every R
void set_search_v1(int on) { g_searchV1 = on; }

// ------------------------------------------------------------------------------------------------
// Batch on the tensor cores, v2 (shipped): H64 = H2 (x) H32.  The butterfly on sample-index bit 5 pairs sample p
// with p+32 -- both sit in the same lane (its two 16-byte chunks of a candidate row) -- so it is done first with
// packed 16-bit adds (VIADD.16x2; int16 wrap is exactly the reference's arithmetic, and only the result mod 2^16
// matters):  top = x[p] + x[p+32],  bot = x[p] - x[p+32].  The two halves of the coefficient vector are then
// H32*top and H32*bot: ONE k-step each against the same 32x32 +-1 matrix -> 16 IMMA per 16 candidates instead of
// 32, and 8 matrix registers instead of 32.  The subtraction is x[p] + ~x[p+32] (= bot - 1 per sample); since
// H32 * (1,1,...,1) = (32,0,...,0) the missing +1 is a +32 on coefficient 0 of the bottom half, supplied for free
// through the accumulator input of that MMA.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t vadd16x2(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("add.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}

// HBM staging: this stream is 97 % reads, and reads must be covered by bytes in flight.  The 2 KiB units (16 candidates) arrive through a per-warp
// ring of ST stages filled by per-lane 16-byte asynchronous copies (cp.async.cg, SASS LDGSTS): every lane reads back only the chunks it copied
// itself, so completion needs nothing but cp.async.wait_group, and (ST - 1) x 2 KiB per warp are in flight at no register cost
// (register double-buffering: 0.91 of the copy roofline, this ring: 0.94-0.99; profiles/r01_time_satd.log).
template <int MINB, int ST>
__global__ void __launch_bounds__(SATDI_WARPS * 32, MINB)
satd8x8_imma3_kernel(const int16_t* __restrict__ diff, int32_t* __restrict__ out, size_t n)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;

    // H32 fragments: K position 16r+4q+i <-> sample 8q+4r+i (of the 32 folded samples), column n' = 8t+g
    uint32_t B[4][2];
#pragma unroll
    for (int t = 0; t < 4; t++)
#pragma unroll
        for (int r = 0; r < 2; r++) {
            uint32_t v = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int pix = 8 * q + 4 * r + i, nn = 8 * t + g;
                v |= ((__popc(nn & pix) & 1) ? 0xFFu : 0x01u) << (8 * i);
            }
            B[t][r] = v;
        }
    const int cZero[4] = { 0, 0, 0, 0 };
    // +32 on bottom-half coefficient 0 (tile 0, column 0 = lanes with q == 0, accumulator slots 0 and 2)
    const int cFix[4] = { q == 0 ? 32 : 0, 0, q == 0 ? 32 : 0, 0 };

    const size_t nUnits = (n + 15) / 16;
    const size_t first = (size_t)blockIdx.x * SATDI_WARPS + warp;
    const size_t stride = (size_t)gridDim.x * SATDI_WARPS;

    // ST-deep ring of 2 KiB units per warp, filled by per-lane 16-byte asynchronous copies (LDGSTS): every lane reads back only what
    // it copied itself, so completion needs no barrier beyond cp.async.wait_group, and the bytes in flight no longer cost registers
    extern __shared__ __align__(16) uint4 ring[];                 // [warp][stage][chunk 0..3][lane]
    const uint32_t ringBase = smem_u32(ring) + (uint32_t)(warp * ST * 4 * 32 + lane) * 16;
    auto issue_unit = [&](size_t u, int stg) {
        if (u < nUnits) {
            size_t c0 = u * 16 + g, c1 = c0 + 8;
            c0 = c0 < n ? c0 : n - 1;
            c1 = c1 < n ? c1 : n - 1;
#pragma unroll
            for (int s = 0; s < 2; s++) {
                cp_async16(ringBase + (uint32_t)((stg * 4 + 2 * s) * 32) * 16, diff + c0 * 64 + 32 * s + 8 * q);
                cp_async16(ringBase + (uint32_t)((stg * 4 + 2 * s + 1) * 32) * 16, diff + c1 * 64 + 32 * s + 8 * q);
            }
        }
        cp_async_commit();                                       // one group per unit slot, empty past the end
    };
#pragma unroll
    for (int k = 0; k < ST - 1; k++) issue_unit(first + (size_t)k * stride, k);
    int stg = 0;

    for (size_t u = first; u < nUnits; u += stride) {
        // fold bit 5, then byte planes.  Fragment registers: [0] row g K 4q+i, [1] row g+8, [2] row g K 16+4q+i, [3] row g+8
        uint32_t TL[4], TH[4], BLo[4], BHi[4];
#pragma unroll
        cp_async_wait<ST - 2>();                                 // the oldest group (this unit) has landed
        uint4 cur4[2][2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            cur4[s][0] = ld_shared_v4(ringBase + (uint32_t)((stg * 4 + 2 * s) * 32) * 16);
            cur4[s][1] = ld_shared_v4(ringBase + (uint32_t)((stg * 4 + 2 * s + 1) * 32) * 16);
        }
        issue_unit(u + (size_t)(ST - 1) * stride, (stg + ST - 1) % ST);
        stg = (stg + 1) % ST;
#pragma unroll
        for (int row = 0; row < 2; row++) {
            const uint4 a = cur4[0][row], b = cur4[1][row];
            const uint32_t tx = vadd16x2(a.x, b.x), ty = vadd16x2(a.y, b.y), tz = vadd16x2(a.z, b.z), tw = vadd16x2(a.w, b.w);
            const uint32_t bx = vadd16x2(a.x, ~b.x), by = vadd16x2(a.y, ~b.y), bz = vadd16x2(a.z, ~b.z), bw = vadd16x2(a.w, ~b.w);
            TL[row] = prmt(tx, ty, 0x6420);      TH[row] = prmt(tx, ty, 0x7531);
            TL[2 + row] = prmt(tz, tw, 0x6420);  TH[2 + row] = prmt(tz, tw, 0x7531);
            BLo[row] = prmt(bx, by, 0x6420);     BHi[row] = prmt(bx, by, 0x7531);
            BLo[2 + row] = prmt(bz, bw, 0x6420); BHi[2 + row] = prmt(bz, bw, 0x7531);
        }

        unsigned s0a = 0, s0b = 0, s1a = 0, s1b = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            int dl[4], dh[4], el[4], eh[4];
            mma_u8s8(dl, TL, B[t][0], B[t][1], cZero);
            mma_s8s8(dh, TH, B[t][0], B[t][1], cZero);
            if (t == 0) mma_u8s8(el, BLo, B[t][0], B[t][1], cFix);
            else mma_u8s8(el, BLo, B[t][0], B[t][1], cZero);
            mma_s8s8(eh, BHi, B[t][0], B[t][1], cZero);
            // lo + 256*hi (IMAD); the int16 coefficient is the low half (= the reference's wrap, satd.c:35).  Two coefficients are packed
            // into one word and |.| is max(x, -x) on both halves (LOP3 + VIADDMNMX.S16x2; -32768 stays 0x8000 = 32768 unsigned, as
            // abs() of the widened value gives in C); IDP.2A adds the two unsigned halves to the accumulator on the FMA pipe:
            // 3 ALU + 3 FMA instructions per coefficient pair instead of 4 + 2 -- the ALU pipe is the busier one here.
            auto abs2 = [](int lo0, int hi0, int lo1, int hi1, unsigned acc) {
                const uint32_t x = prmt((uint32_t)(lo0 + hi0 * 256), (uint32_t)(lo1 + hi1 * 256), 0x5410);
                return __dp2a_lo(__vmaxs2(x, __vneg2(x)), 0x0101u, acc);
            };
            s0a = abs2(dl[0], dh[0], dl[1], dh[1], s0a); s1a = abs2(dl[2], dh[2], dl[3], dh[3], s1a);
            s0b = abs2(el[0], eh[0], el[1], eh[1], s0b); s1b = abs2(el[2], eh[2], el[3], eh[3], s1b);
        }
        unsigned sad0 = s0a + s0b, sad1 = s1a + s1b;
        sad0 += __shfl_xor_sync(0xffffffffu, sad0, 1); sad1 += __shfl_xor_sync(0xffffffffu, sad1, 1);
        sad0 += __shfl_xor_sync(0xffffffffu, sad0, 2); sad1 += __shfl_xor_sync(0xffffffffu, sad1, 2);
        if (q == 0) {
            const size_t c0 = u * 16 + g;
            if (c0 < n) out[c0] = (int)((sad0 + 2) >> 2);
            if (c0 + 8 < n) out[c0 + 8] = (int)((sad1 + 2) >> 2);
        }
    }
}

static std::atomic<int> g_satdCuda{0};      // tuning/diagnostic: 0 = tensor cores, H2 (x) H32 fold, 3-stage cp.async ring (shipped), 1 = CUDA-core kernel, 5 = 4-stage ring
void set_satd_cuda_cores(int on) { g_satdCuda = on; }

cudaError_t launch_satd8x8_batch(const int16_t* diff, int32_t* out, size_t n, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    if (g_satdCuda != 1) {
        size_t want = ((n + 15) / 16 + SATDI_WARPS - 1) / SATDI_WARPS;
        size_t cap = (size_t)sm_count() * 2;
        const int grid = (int)(want < cap ? want : cap);
        constexpr int smem3 = SATDI_WARPS * 3 * 2048, smem4 = SATDI_WARPS * 4 * 2048;
        static std::atomic<bool> attrSet4[64];
        int dev = 0;
        cudaGetDevice(&dev);
        if (g_satdCuda == 5 && (dev < 0 || dev >= 64 || !attrSet4[dev])) {
            cudaError_t e = cudaFuncSetAttribute(satd8x8_imma3_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem4);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) attrSet4[dev] = true;
        }
        if (g_satdCuda == 5) satd8x8_imma3_kernel<2, 4><<<grid, SATDI_WARPS * 32, smem4, st>>>(diff, out, n);
        else satd8x8_imma3_kernel<2, 3><<<grid, SATDI_WARPS * 32, smem3, st>>>(diff, out, n);
        count_launch();
        return cudaGetLastError();
    }
    static std::atomic<bool> attrSet[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attrSet[dev]) {
        cudaError_t e = cudaFuncSetAttribute(satd8x8_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SATD_SMEM);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attrSet[dev] = true;
    }
    size_t want = (n + SATD_WARPS * 32 - 1) / (SATD_WARPS * 32);
    size_t cap = (size_t)sm_count() * 4;
    satd8x8_batch_kernel<<<(int)(want < cap ? want : cap), SATD_WARPS * 32, SATD_SMEM, st>>>(diff, out, n);
    count_launch();
    return cudaGetLastError();
}

template <typename CT>
static cudaError_t launch_satd8x8_search_as(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, int h, int range,
                                            size_t blk0, size_t blk1, CT* cost, int32_t* best, cudaStream_t st)
{
    if (blk1 <= blk0) return cudaSuccess;
    if (range < 0 || range > 2047 || (w & 7) || (h & 7) || blk1 > (size_t)(w / 8) * (h / 8)) return cudaErrorInvalidValue;
    if (g_searchV1 == 0 && (range == 32 || range == 16 || range == 8))
        return launch_satd8x8_search_v3(cur, refPad, strd, w, h, range, blk0, blk1, cost, best, st);
    const int ws = 2 * range + 8, wsp = (ws + 3) & ~3;
    const size_t smem = 128 * sizeof(int) + (size_t)SRCH_WARPS * 8 * wsp * sizeof(int) + (size_t)ws * wsp;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(satd8x8_search_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const size_t nb = blk1 - blk0;
    // grid.x is limited to 2^31-1; frames are far below that
    satd8x8_search_kernel<CT><<<(unsigned)nb, SRCH_WARPS * 32, smem, st>>>(cur, refPad, strd, w, range, blk0, cost, best);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_satd8x8_search(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, int h, int range,
                                  size_t blk0, size_t blk1, uint32_t* cost, int32_t* best, cudaStream_t st)
{
    return launch_satd8x8_search_as(cur, refPad, strd, w, h, range, blk0, blk1, cost, best, st);
}

// 16-bit cost surface: exact, because an 8x8 SATD of 8-bit pixels cannot exceed 32640 (sum_k |T_k| <= sqrt(64) ||T||_2 = 8 * 8 ||x||_2
// <= 64 * 8 * 255 = 130560 before the >> 2 of satd.c:113) and an 8x8 SAD cannot exceed 16320
cudaError_t launch_satd8x8_search(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, int h, int range,
                                  size_t blk0, size_t blk1, uint16_t* cost, int32_t* best, cudaStream_t st)
{
    return launch_satd8x8_search_as(cur, refPad, strd, w, h, range, blk0, blk1, cost, best, st);
}

} // namespace x266
