// intra.cu -- 32x32 intra prediction, 35 modes (0 planar, 1 DC, 2..34 angular).
//
// Spec source: src/mkIntra32-wip.bsv (work in progress in the reference, no C model -> PARITY UNPINNED,
// see DESIGN.md):  reference samples left[64] / top[65] (:34-37), per-row index and fraction tables
// = ((k+1)*angle)>>5 and &31 (:75-112), inverse-angle projection of the side reference for negative
// angles (:151-220), 2-tap interpolation (:352-368), DC (:388-392).
// One warp per prediction, 4 output pixels per lane per step, 128-byte coalesced stores.
#include "common.cuh"
#include "kernels.h"
#include "intra_dev.cuh"

#include <mutex>
#include <vector>

namespace x266 {

// ------------------------------------------------------------------------------------------------
// Tensor-core form of the angular modes with a fractional angle (|angle| < 32, modes 3..17 and 19..33).
// The vertical-family prediction is linear in the reference line:  P[y][x] = sum_k W[y][k] * H[k][x]  with the Hankel
// matrix H[k][x] = ref[base + k + x] and two weights per row, W[y][idx_y + 1 - base] = 8 (32 - f_y), W[y][.. + 1] = 8 f_y
// (base = smallest idx_y + 1 of the mode, so k <= 27 < 32: ONE k32 step).  The weights carry the factor 8, so the pixel is
// byte 1 of the 32-bit sum; the rounding constant 128 rides on tap k = 31, whose Hankel row is patched to ones (one LOP3
// per register); a row with f = 0 uses weight 255 and rounding 255, which is exact for every 8-bit a: (255 a + 255) >> 8 = a.
//   vertical modes:   D = W * H      A = W from a per-mode fragment table, B = 4-byte windows of the reference strip
//   horizontal modes: D = H^T * W^T  A = the same windows of the left reference, B = W^T from the table
// so horizontal predictions come out in output orientation and nothing is transposed.  8 IMMA.16832.U8.U8 per prediction
// instead of 512 two-pixel multiply-adds; the accumulators leave through a per-warp tile as byte pairs (one PRMT each) and
// are stored with the same two 512-byte instructions as before.
// ------------------------------------------------------------------------------------------------

static const int h_intraAngle[35] = { 0, 0, 32, 26, 21, 17, 13, 9, 5, 2, 0, -2, -5, -9, -13, -17, -21, -26,
                                      -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32 };

// Fragment table, [mode][half][lane] as uint4: vertical modes hold the A fragments of W (half = m16 tile, registers a0..a3),
// horizontal modes the B fragments of W^T (half 0 = column tiles 0,1, half 1 = tiles 2,3; registers b0, b1 of each).
// Output columns are permuted over the n8 tiles -- fragment column n of tile t is pixel x = 8(n>>1) + 2t + (n&1) -- so that the
// accumulators of the four tiles give lane (g, q) the EIGHT adjacent pixels 8q..8q+7 of rows g and g+8: the prediction is stored
// straight from the registers, 8 bytes per lane and full 32-byte sectors per row, with no shared-memory round trip.
static std::vector<uint32_t> intra_mma_table_host()
{
    std::vector<uint32_t> tab(35 * 256, 0u);
    for (int mode = 2; mode < 35; mode++) {
        const int ang = h_intraAngle[mode];
        if ((ang & 31) == 0 && ang != 0) continue;                  // |angle| = 32: pure copies, handled without multiplies
        const int base = ang >= 0 ? (ang >> 5) + 1 : ang + 1;
        uint8_t W[32][32] = {};
        for (int y = 0; y < 32; y++) {
            const int t = (y + 1) * ang, idx = t >> 5, f = t & 31, k = idx + 1 - base;
            if (f == 0) { W[y][k] = 255; W[y][31] = 255; }
            else { W[y][k] = (uint8_t)(8 * (32 - f)); W[y][k + 1] = (uint8_t)(8 * f); W[y][31] = 128; }
        }
        for (int lane = 0; lane < 32; lane++) {
            const int g = lane >> 2, q = lane & 3;
            for (int r8 = 0; r8 < 8; r8++) {
                int row, k0;
                if (mode >= 18) { const int m = r8 >> 2, r = r8 & 3; row = 16 * m + g + 8 * (r & 1); k0 = 16 * (r >> 1) + 4 * q; }
                else { const int t = r8 >> 1, r = r8 & 1; row = 8 * (g >> 1) + 2 * t + (g & 1); k0 = 16 * r + 4 * q; }
                uint32_t v = 0;
                for (int i = 0; i < 4; i++) v |= (uint32_t)W[row][k0 + i] << (8 * i);
                tab[mode * 256 + (r8 >> 2) * 128 + lane * 4 + (r8 & 3)] = v;
            }
        }
    }
    return tab;
}

// host-side copy for inspection without a device (include/x266_b200.h: xIntra32MmaTable)
void intra_mma_table_copy(uint32_t* out)
{
    const std::vector<uint32_t> h = intra_mma_table_host();
    for (size_t i = 0; i < h.size(); i++) out[i] = h[i];
}

// Per-device copy of the table.  Uploaded by intra_device_init() when the library first touches a device (ffi.cu: ctx_get), so the
// launchers below only enqueue; released by intra_device_free() (xGpuFree).
static std::atomic<const uint32_t*> g_dTab[64];
static std::mutex g_dTabMu;

cudaError_t intra_device_init()
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lk(g_dTabMu);
    if (g_dTab[dev].load(std::memory_order_relaxed)) return cudaSuccess;
    const std::vector<uint32_t> h = intra_mma_table_host();
    uint32_t* d = nullptr;
    if ((e = cudaMalloc((void**)&d, h.size() * sizeof(uint32_t))) != cudaSuccess) return e;
    if ((e = cudaMemcpy(d, h.data(), h.size() * sizeof(uint32_t), cudaMemcpyHostToDevice)) != cudaSuccess) { cudaFree(d); return e; }
    g_dTab[dev].store(d, std::memory_order_release);
    return cudaSuccess;
}

void intra_device_free()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
    std::lock_guard<std::mutex> lk(g_dTabMu);
    const uint32_t* d = g_dTab[dev].exchange(nullptr);
    if (d) cudaFree(const_cast<uint32_t*>(d));
}

__global__ void mode_range_check_kernel(const uint8_t* __restrict__ mode, size_t n, int maxMode, unsigned* __restrict__ bad)
{
    unsigned c = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) c += mode[i] > maxMode;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(bad, c);
}

cudaError_t launch_mode_range_check(const uint8_t* mode, size_t n, int maxMode, unsigned* bad, cudaStream_t st)
{
    const size_t want = (n + 255) / 256, cap = (size_t)sm_count() * 8;
    mode_range_check_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(mode, n, maxMode, bad);
    count_launch();
    return cudaGetLastError();
}

const uint32_t* intra_mma_table_dev(cudaError_t* err)
{
    int dev = 0;
    *err = cudaGetDevice(&dev);
    if (*err != cudaSuccess) return nullptr;
    if (dev < 0 || dev >= 64) { *err = cudaErrorInvalidDevice; return nullptr; }
    const uint32_t* d = g_dTab[dev].load(std::memory_order_acquire);
    if (!d) *err = cudaErrorInitializationError;       // the C ABI initialises the device before any launch (ffi.cu: dev_ready)
    return d;
}

// One warp per prediction.  Lane l produces, for it = 0..7, the 4 pixels (row 4*it + (l>>3), columns
// 4*(l&7) .. +3) and stores them as one 32-bit word: every warp store is 128 contiguous bytes (4 rows).
// The reference line ref[-32..65] lives in a per-warp shared-memory strip; VER selects which of
// (row, column) is the distance from the main reference so that the loop-invariant index/fraction
// computations are hoisted (per row for vertical modes, per lane for horizontal modes).
// Generates prediction `mode` of one block into out[0..1023] (one warp).  The main part of the reference line is already in the strip
// (ref[0] at byte 36 for the vertical family, 35 for the horizontal one -- see the callers), sraw holds left[64] | top[65] whenever the
// mode reads it (planar, DC, negative angles).  T0 / T1 = the mode's fragment-table row if the caller has requested it already.
// MMA = false compiles the CUDA-core SWAR generator for every angular mode and the scalar planar (xGpuTune(8, 1)) INSTEAD of the tensor-core paths:
// one instantiation never carries both, the loop body has to stay inside the instruction cache (35 modes are in flight at once).
template <bool EARLYTAB, bool MMA>
__device__ __forceinline__ void intra_generate(int mode, uint8_t* __restrict__ stripB, const uint8_t* __restrict__ sraw, uint8_t* __restrict__ tileW,
                                               uint32_t* __restrict__ out, const uint4* __restrict__ tp, uint4 T0, uint4 T1, int lane)
{
    uint32_t* strip32 = reinterpret_cast<uint32_t*>(stripB);
    const uint8_t* left = sraw;          // left[i] = pixel (-1, i)
    const uint8_t* top = sraw + 64;      // top[0] = corner, top[1+i] = pixel (i, -1)
    const bool isVer = mode >= 18;
    const int ang = c_intraAngle[mode];
    const bool frac = MMA && mode >= 2 && (ang & 31) != 0;
    if (mode >= 2) {
        const int ref0 = isVer ? 36 : 35;
        uint8_t* sref = stripB + ref0;
        if (ang < 0) {
            __syncwarp();                                     // sraw complete
            const int inv = c_intraInvMode[mode];
            const int k = lane + 1;                           // projects ref[-k], k = 1..32
            if (-k >= ang) {
                const int s = (k * inv + 128) >> 8;
                sref[-k] = isVer ? left[s - 1] : top[s];
            }
        }
        __syncwarp();
        if (MMA && frac) {
            // ---- tensor-core path (see the comment above mma_u8u8_16832)
            const int g4 = lane >> 2, q4 = lane & 3;
            const int base = ang >= 0 ? (ang >> 5) + 1 : ang + 1;
            const uint32_t keep = q4 == 3 ? 0x00FFFFFFu : 0xFFFFFFFFu, one = q4 == 3 ? 0x01000000u : 0u;   // Hankel row k = 31 := 1
            if (!EARLYTAB) { T0 = __ldg(tp); T1 = __ldg(tp + 32); }
            uint8_t* orow = reinterpret_cast<uint8_t*>(out) + g4 * 32 + 8 * q4;      // row g, pixels 8q..8q+7
            // pixel = byte 1 of each sum: four sums -> one word (merge on the FMA pipe, intra_dev.cuh)
            const uint32_t k16 = intra_k16();
            auto word = [k16](const int (&a)[4], const int (&b)[4], int r) { return intra_word_fma(a, b, r, k16); };
            if (isVer) {
                // B = Hankel windows with permuted columns: column n = g of tile t is pixel 8(g>>1) + 2t + (g&1)
                const int cb = ref0 + base + 4 * q4 + 8 * (g4 >> 1) + (g4 & 1);
                const uint32_t* wp = strip32 + (cb >> 2);
                const int sh = (cb & 3) * 8;
                uint32_t wl[4], wh[4];                                   // windows at byte offsets 2t and 16 + 2t
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const uint32_t xa = wp[4 * j], xb = wp[4 * j + 1], xc = wp[4 * j + 2], xd = wp[4 * j + 3];
                    const uint32_t lo = __funnelshift_r(xa, xb, sh), mid = __funnelshift_r(xb, xc, sh), hi = __funnelshift_r(xc, xd, sh);
                    uint32_t* w = j ? wh : wl;
                    w[0] = lo; w[1] = __byte_perm(lo, mid, 0x5432); w[2] = mid; w[3] = __byte_perm(mid, hi, 0x5432);
                }
#pragma unroll
                for (int t = 0; t < 4; t++) wh[t] = (wh[t] & keep) | one;
#pragma unroll
                for (int m = 0; m < 2; m++) {
                    const uint4 A = m ? T1 : T0;
                    int d[4][4];
#pragma unroll
                    for (int t = 0; t < 4; t++) mma_u8u8_16832(d[t], A.x, A.y, A.z, A.w, wl[t], wh[t]);
                    st_global_stream_v2(orow + (16 * m) * 32, make_uint2(word(d[0], d[1], 0), word(d[2], d[3], 0)));
                    st_global_stream_v2(orow + (16 * m + 8) * 32, make_uint2(word(d[0], d[1], 1), word(d[2], d[3], 1)));
                }
            } else {
                // A = Hankel windows of the left reference (rows y = 16m + g, g + 8), B = W^T from the table
                const int ca = ref0 + base + 4 * q4 + g4;
                const uint32_t* wp = strip32 + (ca >> 2);
                const int sh = (ca & 3) * 8;
                uint32_t win[6];
#pragma unroll
                for (int j = 0; j < 6; j++) win[j] = __funnelshift_r(wp[2 * j], wp[2 * j + 1], sh);
                uint32_t wpat[4];
#pragma unroll
                for (int j = 0; j < 4; j++) wpat[j] = (win[j + 2] & keep) | one;
#pragma unroll
                for (int m = 0; m < 2; m++) {
                    int d[4][4];
                    mma_u8u8_16832(d[0], win[2 * m], win[2 * m + 1], wpat[2 * m], wpat[2 * m + 1], T0.x, T0.y);
                    mma_u8u8_16832(d[1], win[2 * m], win[2 * m + 1], wpat[2 * m], wpat[2 * m + 1], T0.z, T0.w);
                    mma_u8u8_16832(d[2], win[2 * m], win[2 * m + 1], wpat[2 * m], wpat[2 * m + 1], T1.x, T1.y);
                    mma_u8u8_16832(d[3], win[2 * m], win[2 * m + 1], wpat[2 * m], wpat[2 * m + 1], T1.z, T1.w);
                    st_global_stream_v2(orow + (16 * m) * 32, make_uint2(word(d[0], d[1], 0), word(d[2], d[3], 0)));
                    st_global_stream_v2(orow + (16 * m + 8) * 32, make_uint2(word(d[0], d[1], 1), word(d[2], d[3], 1)));
                }
            }
            __syncwarp();
            return;
        }
        uint32_t w[2][4];
        if (mode == 10) {
            // pure horizontal: P[y][x] = left[y] -- a row is one byte repeated, nothing to transpose
#pragma unroll
            for (int it = 0; it < 2; it++) {
                const uint32_t v = (uint32_t)sref[1 + 16 * it + (lane >> 1)] * 0x01010101u;
                w[it][0] = v; w[it][1] = v; w[it][2] = v; w[it][3] = v;
            }
        } else if (MMA) {
            // what is left for this path when the tensor cores take the fractional angles: |angle| = 32, rows are unaligned copies
            // of the working line (modes 2, 18, 34) or the line itself (26)
#pragma unroll
            for (int it = 0; it < 2; it++) {
                const int row = 16 * it + (lane >> 1), half = lane & 1;
                const int o = ref0 + 16 * half + (((row + 1) * ang) >> 5) + 1;
                const uint32_t* p = strip32 + (o >> 2);
                const int sh = (o & 3) * 8;
                const uint32_t x0 = p[0], x1 = p[1], x2 = p[2], x3 = p[3], x4 = p[4];
                w[it][0] = __funnelshift_r(x0, x1, sh); w[it][1] = __funnelshift_r(x1, x2, sh);
                w[it][2] = __funnelshift_r(x2, x3, sh); w[it][3] = __funnelshift_r(x3, x4, sh);
            }
        } else {
            intra_angular_rows(strip32, ref0, ang, lane, w);
        }
        if (MMA || isVer || mode == 10 || mode == 2) {
            // mode 2 (angle +32 on the left reference): P_v[r][c] = ref[c + r + 2] is symmetric, so P = P_v^T = P_v
#pragma unroll
            for (int it = 0; it < 2; it++)
                st_global_stream(reinterpret_cast<uint4*>(out) + it * 32 + lane, make_uint4(w[it][0], w[it][1], w[it][2], w[it][3]));
        } else {
#pragma unroll
            for (int it = 0; it < 2; it++) {
                uint32_t* trow = reinterpret_cast<uint32_t*>(&tileW[(16 * it + (lane >> 1)) * 36 + 16 * (lane & 1)]);
                trow[0] = w[it][0]; trow[1] = w[it][1]; trow[2] = w[it][2]; trow[3] = w[it][3];
            }
            __syncwarp();
            // lane <-> output rows r0..r0+3, columns c0..c0+7: two 4x4 byte blocks of P_v^T
            const int r0 = 4 * (lane >> 2), c0 = 8 * (lane & 3);
            uint32_t W[8];
#pragma unroll
            for (int k = 0; k < 8; k++) W[k] = *reinterpret_cast<const uint32_t*>(&tileW[(c0 + k) * 36 + r0]);
            uint32_t o8[4][2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t t0 = __byte_perm(W[4 * h], W[4 * h + 1], 0x5140), t1 = __byte_perm(W[4 * h + 2], W[4 * h + 3], 0x5140);
                const uint32_t t2 = __byte_perm(W[4 * h], W[4 * h + 1], 0x7362), t3 = __byte_perm(W[4 * h + 2], W[4 * h + 3], 0x7362);
                o8[0][h] = __byte_perm(t0, t1, 0x5410); o8[1][h] = __byte_perm(t0, t1, 0x7632);
                o8[2][h] = __byte_perm(t2, t3, 0x5410); o8[3][h] = __byte_perm(t2, t3, 0x7632);
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
                *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(out) + (r0 + i) * 32 + c0) = make_uint2(o8[i][0], o8[i][1]);
        }
    } else if (mode == 1) {
        __syncwarp();
        int s = left[lane] + top[1 + lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const uint32_t dc = (uint32_t)((s + 32) >> 6) * 0x01010101u;
#pragma unroll
        for (int it = 0; it < 8; it++) out[it * 32 + lane] = dc;
    } else if (MMA) {
        // planar on the tensor cores: 4 (planar sum + 32) = sum_k A[y][k] B[k][x] with the five taps
        //   A[y] = ( left[y], TR, 4 (31 - y), 4 (y + 1), 128 ),   B[.][x] = ( 4 (31 - x), 4 (x + 1), top[x], BL, 1 )      (all u8)
        // so the pixel is byte 1 of the 32-bit sum, exactly as in the angular path: same column permutation, same epilogue.
        __syncwarp();
        const int g4 = lane >> 2, q4 = lane & 3;
        const uint32_t tr = top[33], bl = left[32];
        uint32_t Bf[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int x = 8 * (g4 >> 1) + 2 * t + (g4 & 1);
            const uint32_t v = (uint32_t)(4 * (31 - x)) | ((uint32_t)(4 * (x + 1)) << 8) | ((uint32_t)top[1 + x] << 16) | (bl << 24);
            Bf[t] = q4 == 0 ? v : q4 == 1 ? 1u : 0u;
        }
        uint8_t* orow = reinterpret_cast<uint8_t*>(out) + g4 * 32 + 8 * q4;
        const uint32_t k16 = intra_k16();
        auto word = [k16](const int (&a)[4], const int (&b)[4], int r) { return intra_word_fma(a, b, r, k16); };
#pragma unroll
        for (int m = 0; m < 2; m++) {
            uint32_t Af[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int y = 16 * m + g4 + 8 * h;
                const uint32_t v = (uint32_t)left[y] | (tr << 8) | ((uint32_t)(4 * (31 - y)) << 16) | ((uint32_t)(4 * (y + 1)) << 24);
                Af[h] = q4 == 0 ? v : q4 == 1 ? 128u : 0u;
            }
            int d[4][4];
#pragma unroll
            for (int t = 0; t < 4; t++) mma_u8u8_16832(d[t], Af[0], Af[1], 0u, 0u, Bf[t], 0u);
            st_global_stream_v2(orow + (16 * m) * 32, make_uint2(word(d[0], d[1], 0), word(d[2], d[3], 0)));
            st_global_stream_v2(orow + (16 * m + 8) * 32, make_uint2(word(d[0], d[1], 1), word(d[2], d[3], 1)));
        }
    } else {
        __syncwarp();
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
}

// The generic predictor: one warp per prediction, modes in the caller's order.  This kernel is kept as ONE monolithic body on purpose: ptxas's
// schedule of it is fragile -- the same code routed through intra_generate() above, or with the two-line change that stores modes 2 / 10
// without the transpose, lands on a different register allocation (4 CTAs per SM, or 5 with the MMA section serialised) and loses 4-5 % on
// the 30 fractional modes (A/B on one box: 0.762 of the HBM roofline as below, 0.725 refactored; profiles/r02_intra_experiments.md).
// intra_generate() carries the improved whole-sample / planar paths and serves the mode-major kernel.
template <bool ALIGNED>
__global__ void __launch_bounds__(INTRA_WARPS * 32)
intra32_kernel(const uint8_t* __restrict__ refs, const uint8_t* __restrict__ modes, uint8_t* __restrict__ pred, size_t n,
               const uint32_t* __restrict__ mmaTab, int useMma)
{
    __shared__ __align__(16) uint8_t strip[INTRA_WARPS][INTRA_STRIP + 16];
    __shared__ __align__(16) uint8_t raw[INTRA_WARPS][144];      // left[64] | top[65]
    __shared__ __align__(16) uint8_t tile[INTRA_WARPS][32 * 48]; // output tile of the tensor-core path (pitch 48) / transpose tile (pitch 36)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sraw = raw[warp];
    uint32_t* strip32 = reinterpret_cast<uint32_t*>(strip[warp]);

    // software pipeline: the 129 reference bytes and the mode of the NEXT prediction are in flight (registers) while
    // this one is generated -- otherwise every prediction pays a full DRAM latency with nothing to overlap it.
    // The bytes travel as aligned 32-bit words: lane l holds word l of the 33 words that cover [129q - a, 129q + 129),
    // a = (129 q) & 3; one shuffle and one funnel shift re-align them to raw[4l .. 4l+3].  Addresses come from a 32-bit
    // prediction index with one wide multiply-add each.
    const uint32_t pstride = gridDim.x * INTRA_WARPS;
    const uint32_t nLast = (uint32_t)n - 1;                          // n < 2^32: a prediction is 1 KiB of output
    uint32_t p = blockIdx.x * INTRA_WARPS + warp;                    // prediction being generated
    if (p > nLast) return;
    uint32_t nw0 = 0, nw1 = 0;
    int na = 0, nmode = 1;
    auto prefetch = [&](uint32_t q) {
        const uint8_t* src = refs + (size_t)q * 129;
        if (ALIGNED) {
            na = (int)(reinterpret_cast<uintptr_t>(src) & 3);
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(src - na);
            nw0 = __ldg(wp + lane);
            if (lane == 0) {
                if (q != nLast) nw1 = __ldg(wp + 32);
                else {                                           // bytes 0..na of word 32 only, never past the array
                    const uint8_t* t = src - na + 128;
                    nw1 = t[0];
                    for (int j = 1; j <= na; j++) nw1 |= (uint32_t)t[j] << (8 * j);
                }
            }
        } else {
            nw0 = src[4 * lane] | (src[4 * lane + 1] << 8) | (src[4 * lane + 2] << 16) | ((uint32_t)src[4 * lane + 3] << 24);
            if (lane == 0) nw1 = src[128];
        }
        nmode = modes[q];
    };
    prefetch(p);

    for (; p <= nLast; p += pstride) {
        const int mode = nmode > 34 ? 1 : nmode;                 // host API rejects > 34; keep device reads in range
        uint32_t R, last;                                        // R = raw[4*lane .. 4*lane+3], last = raw[128]
        {
            uint32_t up = __shfl_down_sync(0xffffffffu, nw0, 1);
            const uint32_t w32 = __shfl_sync(0xffffffffu, nw1, 0);
            if (lane == 31) up = w32;
            R = __funnelshift_r(nw0, up, 8 * na);
            last = (w32 >> (8 * na)) & 0xFFu;
        }
        if (mode < 2 || (mode >= 11 && mode <= 25)) {          // only DC / planar and the negative-angle projections read the staged bytes
            reinterpret_cast<uint32_t*>(sraw)[lane] = R;
            if (lane == 0) sraw[128] = (uint8_t)last;
        }
        uint32_t* out = reinterpret_cast<uint32_t*>(pred + (size_t)p * 1024);
        if (p + pstride <= nLast) prefetch(p + pstride);
        const uint8_t* left = sraw;          // left[i] = pixel (-1, i)
        const uint8_t* top = sraw + 64;      // top[0] = corner, top[1+i] = pixel (i, -1)
        const bool isVer = mode >= 18;
        const int ang = c_intraAngle[mode];

        if (mode >= 2) {
            // reference line ref[-32..65] in the strip, ref[0] at byte ref0: vertical modes ref[i] = top[i] = raw[64+i]
            // (word aligned at 36), horizontal modes ref[0] = corner, ref[1+j] = left[j] = raw[j] (ref[1] word aligned at 36)
            // -- so the main part is one 32-bit store per lane straight from the registers.
            const int ref0 = isVer ? 36 : 35;
            uint8_t* sref = strip[warp] + ref0;
            if (isVer) {
                if (lane >= 16) strip32[9 + lane - 16] = R;
                if (lane == 0) sref[64] = (uint8_t)last;
            } else {
                if (lane < 16) strip32[9 + lane] = R;
                if (lane == 16) sref[0] = (uint8_t)R;
            }
            if (ang < 0) {
                __syncwarp();                                     // sraw complete
                const int inv = c_intraInvMode[mode];
                const int k = lane + 1;                           // projects ref[-k], k = 1..32
                if (-k >= ang) {
                    const int s = (k * inv + 128) >> 8;
                    sref[-k] = isVer ? left[s - 1] : top[s];
                }
            }
            __syncwarp();
            if (useMma && (ang & 31) != 0) {
                // ---- tensor-core path (see the comment above mma_u8u8_16832)
                const int g4 = lane >> 2, q4 = lane & 3;
                const int base = ang >= 0 ? (ang >> 5) + 1 : ang + 1;
                const uint32_t keep = q4 == 3 ? 0x00FFFFFFu : 0xFFFFFFFFu, one = q4 == 3 ? 0x01000000u : 0u;   // Hankel row k = 31 := 1
                const uint4* tp = reinterpret_cast<const uint4*>(mmaTab) + mode * 64 + lane;
                const uint4 T0 = __ldg(tp), T1 = __ldg(tp + 32);
                uint8_t* orow = reinterpret_cast<uint8_t*>(out) + g4 * 32 + 8 * q4;      // row g, pixels 8q..8q+7
                // pixel = byte 1 of each sum: four sums -> one word
                auto word = [](const int (&a)[4], const int (&b)[4], int r) {
                    return __byte_perm(__byte_perm((uint32_t)a[2 * r], (uint32_t)a[2 * r + 1], 0x5151),
                                       __byte_perm((uint32_t)b[2 * r], (uint32_t)b[2 * r + 1], 0x5151), 0x5410);
                };
                if (isVer) {
                    // B = Hankel windows with permuted columns: column n = g of tile t is pixel 8(g>>1) + 2t + (g&1)
                    const int cb = ref0 + base + 4 * q4 + 8 * (g4 >> 1) + (g4 & 1);
                    const uint32_t* wp = strip32 + (cb >> 2);
                    const int sh = (cb & 3) * 8;
                    uint32_t wl[4], wh[4];                                   // windows at byte offsets 2t and 16 + 2t
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        const uint32_t xa = wp[4 * j], xb = wp[4 * j + 1], xc = wp[4 * j + 2], xd = wp[4 * j + 3];
                        const uint32_t lo = __funnelshift_r(xa, xb, sh), mid = __funnelshift_r(xb, xc, sh), hi = __funnelshift_r(xc, xd, sh);
                        uint32_t* w = j ? wh : wl;
                        w[0] = lo; w[1] = __byte_perm(lo, mid, 0x5432); w[2] = mid; w[3] = __byte_perm(mid, hi, 0x5432);
                    }
#pragma unroll
                    for (int t = 0; t < 4; t++) wh[t] = (wh[t] & keep) | one;
#pragma unroll
                    for (int m = 0; m < 2; m++) {
                        const uint4 A = m ? T1 : T0;
                        int d[4][4];
#pragma unroll
                        for (int t = 0; t < 4; t++) mma_u8u8_16832(d[t], A.x, A.y, A.z, A.w, wl[t], wh[t]);
                        *reinterpret_cast<uint2*>(orow + (16 * m) * 32) = make_uint2(word(d[0], d[1], 0), word(d[2], d[3], 0));
                        *reinterpret_cast<uint2*>(orow + (16 * m + 8) * 32) = make_uint2(word(d[0], d[1], 1), word(d[2], d[3], 1));
                    }
                } else {
                    // A = Hankel windows of the left reference (rows y = 16m + g, g + 8), B = W^T from the table
                    const int ca = ref0 + base + 4 * q4 + g4;
                    const uint32_t* wp = strip32 + (ca >> 2);
                    const int sh = (ca & 3) * 8;
                    uint32_t win[6];
#pragma unroll
                    for (int j = 0; j < 6; j++) win[j] = __funnelshift_r(wp[2 * j], wp[2 * j + 1], sh);
                    uint32_t wpat[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) wpat[j] = (win[j + 2] & keep) | one;
#pragma unroll
                    for (int m = 0; m < 2; m++) {
                        int d[4][4];
                        mma_u8u8_16832(d[0], win[2 * m], win[2 * m + 1], wpat[2 * m], wpat[2 * m + 1], T0.x, T0.y);
                        mma_u8u8_16832(d[1], win[2 * m], win[2 * m + 1], wpat[2 * m], wpat[2 * m + 1], T0.z, T0.w);
                        mma_u8u8_16832(d[2], win[2 * m], win[2 * m + 1], wpat[2 * m], wpat[2 * m + 1], T1.x, T1.y);
                        mma_u8u8_16832(d[3], win[2 * m], win[2 * m + 1], wpat[2 * m], wpat[2 * m + 1], T1.z, T1.w);
                        *reinterpret_cast<uint2*>(orow + (16 * m) * 32) = make_uint2(word(d[0], d[1], 0), word(d[2], d[3], 0));
                        *reinterpret_cast<uint2*>(orow + (16 * m + 8) * 32) = make_uint2(word(d[0], d[1], 1), word(d[2], d[3], 1));
                    }
                }
                __syncwarp();
                continue;
            }
            uint32_t w[2][4];
            intra_angular_rows(strip32, ref0, ang, lane, w);
            if (isVer) {
#pragma unroll
                for (int it = 0; it < 2; it++)
                    reinterpret_cast<uint4*>(out)[it * 32 + lane] = make_uint4(w[it][0], w[it][1], w[it][2], w[it][3]);
            } else {
#pragma unroll
                for (int it = 0; it < 2; it++) {
                    uint32_t* trow = reinterpret_cast<uint32_t*>(&tile[warp][(16 * it + (lane >> 1)) * 36 + 16 * (lane & 1)]);
                    trow[0] = w[it][0]; trow[1] = w[it][1]; trow[2] = w[it][2]; trow[3] = w[it][3];
                }
                __syncwarp();
                // lane <-> output rows r0..r0+3, columns c0..c0+7: two 4x4 byte blocks of P_v^T
                const int r0 = 4 * (lane >> 2), c0 = 8 * (lane & 3);
                uint32_t W[8];
#pragma unroll
                for (int k = 0; k < 8; k++) W[k] = *reinterpret_cast<const uint32_t*>(&tile[warp][(c0 + k) * 36 + r0]);
                uint32_t o8[4][2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t t0 = __byte_perm(W[4 * h], W[4 * h + 1], 0x5140), t1 = __byte_perm(W[4 * h + 2], W[4 * h + 3], 0x5140);
                    const uint32_t t2 = __byte_perm(W[4 * h], W[4 * h + 1], 0x7362), t3 = __byte_perm(W[4 * h + 2], W[4 * h + 3], 0x7362);
                    o8[0][h] = __byte_perm(t0, t1, 0x5410); o8[1][h] = __byte_perm(t0, t1, 0x7632);
                    o8[2][h] = __byte_perm(t2, t3, 0x5410); o8[3][h] = __byte_perm(t2, t3, 0x7632);
                }
#pragma unroll
                for (int i = 0; i < 4; i++)
                    *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(out) + (r0 + i) * 32 + c0) = make_uint2(o8[i][0], o8[i][1]);
            }
        } else if (mode == 1) {
            __syncwarp();
            int s = left[lane] + top[1 + lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const uint32_t dc = (uint32_t)((s + 32) >> 6) * 0x01010101u;
#pragma unroll
            for (int it = 0; it < 8; it++) out[it * 32 + lane] = dc;
        } else {
            __syncwarp();
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
// (A first CUDA-core version -- one half-warp per mode, lane <-> sub-block, 10.1 M blocks/s -- was replaced by the
// tensor-core form below, 45-50 M blocks/s.)
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// Mode decision v2 (shipped): the 16 sub-blocks of one (block, mode) are exactly one m16 tile of the
// tensor-core SATD (satd.cu): A = the 8-bit prediction pixels (a single u8 plane, no byte split), B = the +-1
// Sylvester matrix, 16 IMMA per mode.  By linearity the cost is sum |T(cur) - T(pred)| against the transform of
// the current block, computed once per block with the same MMA and parked in shared memory in accumulator
// layout.  Horizontal modes (2..17) are evaluated on the TRANSPOSED problem: pred_h = P_v^T where P_v is the
// vertical-family routine run on the left reference, and SATD(X^T) = SATD(X) for every 8x8 block (H is
// symmetric), so one SWAR row generator (8 pixels of one prediction row = one A-fragment register pair,
// 2 pixels per 32-bit multiply-add) serves all 33 angular modes.  One warp owns one mode at a time.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(IDEC_WARPS * 32, 2)
intra32_decide_v2_kernel(const uint8_t* __restrict__ cur, const uint8_t* __restrict__ refs, uint32_t* __restrict__ cost,
                         int32_t* __restrict__ bestMode, size_t n)
{
    __shared__ DecideSmem sm;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t B[2][8][2];
    decide_hadamard_fragments(B, lane >> 2, lane & 3);
    // a pass = IDEC_NB consecutive blocks: 70 (block, mode) items over the 8 warps; the inputs of the next pass are loaded during this one
    const size_t stride = (size_t)gridDim.x * IDEC_NB;
    size_t p = (size_t)blockIdx.x * IDEC_NB;
    DecideIn in;
    decide_load(in, cur, refs, p, n, tid);
    for (; p < n; p += stride) {
        const int nblk = (n - p) < (size_t)IDEC_NB ? (int)(n - p) : IDEC_NB;
        decide_blocks(sm, B, in, nblk, cur, refs, p + stride, n, tid);
        if (warp < nblk) {
            const int bm = decide_output(sm, warp, cost + (p + warp) * 35, lane);
            if (lane == 0) bestMode[p + warp] = bm;
        }
        // no barrier here: the next pass writes scost only after its own two barriers, and scur / sraw are not read after decide_blocks' last one
    }
}

void set_decide_v1(int) {}             // (xGpuTune key 5 selected the removed CUDA-core decision kernel; kept as a no-op for old scripts)

cudaError_t launch_intra32_decide(const uint8_t* cur, const uint8_t* refs, uint32_t* cost, int32_t* bestMode, size_t n, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    const size_t cap = (size_t)sm_count() * resident_ctas_per_sm((const void*)intra32_decide_v2_kernel, IDEC_WARPS * 32, 0);
    const size_t passes = (n + IDEC_NB - 1) / IDEC_NB;
    intra32_decide_v2_kernel<<<(unsigned)(passes < cap ? passes : cap), IDEC_WARPS * 32, 0, st>>>(cur, refs, cost, bestMode, n);
    count_launch();
    return cudaGetLastError();
}

static std::atomic<int> g_intraCtas{0};      // tuning/diagnostic: CTAs per SM of the persistent grid (0 = 8)
void set_intra_ctas(int v) { g_intraCtas = v; }
static std::atomic<int> g_intraSwar{0};      // tuning/diagnostic: 1 = CUDA-core SWAR interpolation for every angular mode
void set_intra_swar(int on) { g_intraSwar = on; }

// ------------------------------------------------------------------------------------------------
// Mode-major entry (xIntra32PredModes): ONE block's 129 reference bytes -> every wanted mode, as the encoder's rate-distortion search
// asks for them.  A warp owns a block: the reference bytes are loaded and re-aligned once, both working lines (top-based for the
// vertical family, left-based for the horizontal one) are staged once, and only the negative-angle projections are redone per mode --
// the 130 input bytes and their staging are paid once per 35 predictions instead of once per prediction.
// pred[b][j] = prediction of the j-th set bit of modeMask, 1 KiB each.
// ------------------------------------------------------------------------------------------------
template <bool ALIGNED, bool MMA>
__global__ void __launch_bounds__(INTRA_WARPS * 32, 5)
intra32_modes_kernel(const uint8_t* __restrict__ refs, unsigned long long modeMask, int nModes, uint8_t* __restrict__ pred, size_t nBlocks,
                     const uint32_t* __restrict__ mmaTab)
{
    __shared__ __align__(16) uint8_t stripV[INTRA_WARPS][INTRA_STRIP + 16];
    __shared__ __align__(16) uint8_t stripH[INTRA_WARPS][INTRA_STRIP + 16];
    __shared__ __align__(16) uint8_t raw[INTRA_WARPS][144];
    __shared__ __align__(16) uint8_t tile[INTRA_WARPS][MMA ? 16 : 32 * 36];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sraw = raw[warp];
    const size_t stride = (size_t)gridDim.x * INTRA_WARPS;
    size_t b = (size_t)blockIdx.x * INTRA_WARPS + warp;
    if (b >= nBlocks) return;
    uint32_t nw0 = 0, nw1 = 0;
    int na = 0;
    auto prefetch = [&](size_t q) {
        const uint8_t* src = refs + q * 129;
        if (ALIGNED) {
            na = (int)(reinterpret_cast<uintptr_t>(src) & 3);
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(src - na);
            nw0 = __ldg(wp + lane);
            if (lane == 0) {
                if (q != nBlocks - 1) nw1 = __ldg(wp + 32);
                else {                                           // bytes 0..na of word 32 only, never past the array
                    const uint8_t* t = src - na + 128;
                    nw1 = t[0];
                    for (int j = 1; j <= na; j++) nw1 |= (uint32_t)t[j] << (8 * j);
                }
            }
        } else {
            nw0 = src[4 * lane] | (src[4 * lane + 1] << 8) | (src[4 * lane + 2] << 16) | ((uint32_t)src[4 * lane + 3] << 24);
            if (lane == 0) nw1 = src[128];
        }
    };
    prefetch(b);
    for (; b < nBlocks; b += stride) {
        uint32_t R, last;
        {
            uint32_t up = __shfl_down_sync(0xffffffffu, nw0, 1);
            const uint32_t w32 = __shfl_sync(0xffffffffu, nw1, 0);
            if (lane == 31) up = w32;
            R = __funnelshift_r(nw0, up, 8 * na);
            last = (w32 >> (8 * na)) & 0xFFu;
        }
        __syncwarp();                                            // the previous block's last mode has finished with the staging areas
        reinterpret_cast<uint32_t*>(sraw)[lane] = R;
        if (lane == 0) sraw[128] = (uint8_t)last;
        // both working lines: vertical ref[i] = top[i] = raw[64+i] (ref[0] at byte 36), horizontal ref[0] = corner, ref[1+j] = left[j] (ref[0] at 35)
        if (lane >= 16) reinterpret_cast<uint32_t*>(stripV[warp])[9 + lane - 16] = R;
        else reinterpret_cast<uint32_t*>(stripH[warp])[9 + lane] = R;
        if (lane == 0) stripV[warp][36 + 64] = (uint8_t)last;
        if (lane == 16) stripH[warp][35] = (uint8_t)R;
        if (b + stride < nBlocks) prefetch(b + stride);
        __syncwarp();
        uint8_t* outB = pred + b * (size_t)nModes * 1024;
        unsigned long long left64 = modeMask;
        int j = 0;
        while (left64) {
            const int mode = __ffsll((long long)left64) - 1;
            left64 &= left64 - 1;
            const uint4* tp = reinterpret_cast<const uint4*>(mmaTab) + mode * 64 + lane;
            intra_generate<false, MMA>(mode, mode >= 18 ? stripV[warp] : stripH[warp], sraw, tile[warp], reinterpret_cast<uint32_t*>(outB + (size_t)j * 1024),
                                       tp, make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), lane);
            __syncwarp();
            j++;
        }
    }
}

template <bool ALIGNED, bool MMA>
static cudaError_t launch_intra32_modes_as(const uint8_t* refs, unsigned long long modeMask, int nModes, uint8_t* pred, size_t nBlocks,
                                           const uint32_t* tab, cudaStream_t st)
{
    const void* kern = (const void*)intra32_modes_kernel<ALIGNED, MMA>;
    const size_t want = (nBlocks + INTRA_WARPS - 1) / INTRA_WARPS;
    const int perSm = g_intraCtas > 0 ? g_intraCtas.load() : resident_ctas_per_sm(kern, INTRA_WARPS * 32, 0);
    const size_t cap = (size_t)sm_count() * perSm;
    intra32_modes_kernel<ALIGNED, MMA><<<(unsigned)(want < cap ? want : cap), INTRA_WARPS * 32, 0, st>>>(refs, modeMask, nModes, pred, nBlocks, tab);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_intra32_modes(const uint8_t* refs, unsigned long long modeMask, uint8_t* pred, size_t nBlocks, cudaStream_t st)
{
    modeMask &= (1ull << 35) - 1;
    const int nModes = __builtin_popcountll(modeMask);
    if (nBlocks == 0 || nModes == 0) return cudaSuccess;
    cudaError_t e;
    const uint32_t* tab = intra_mma_table_dev(&e);
    if (!tab) return e;
    const bool aligned4 = (reinterpret_cast<uintptr_t>(refs) & 3) == 0;
    if (g_intraSwar) return aligned4 ? launch_intra32_modes_as<true, false>(refs, modeMask, nModes, pred, nBlocks, tab, st)
                                     : launch_intra32_modes_as<false, false>(refs, modeMask, nModes, pred, nBlocks, tab, st);
    return aligned4 ? launch_intra32_modes_as<true, true>(refs, modeMask, nModes, pred, nBlocks, tab, st)
                    : launch_intra32_modes_as<false, true>(refs, modeMask, nModes, pred, nBlocks, tab, st);
}

cudaError_t launch_intra32(const uint8_t* refs, const uint8_t* mode, uint8_t* pred, size_t n, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    if (n > ((size_t)1 << 31)) return cudaErrorInvalidValue;          // 32-bit prediction index in the kernel (2 TiB of output)
    const size_t want = (n + INTRA_WARPS - 1) / INTRA_WARPS;
    const bool aligned4 = (reinterpret_cast<uintptr_t>(refs) & 3) == 0;
    const int perSm = g_intraCtas > 0 ? g_intraCtas.load()
                                      : resident_ctas_per_sm(aligned4 ? (const void*)intra32_kernel<true> : (const void*)intra32_kernel<false>, INTRA_WARPS * 32, 0);
    const size_t cap = (size_t)sm_count() * perSm;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    cudaError_t e;
    const uint32_t* tab = intra_mma_table_dev(&e);
    if (!tab) return e;
    if (aligned4) intra32_kernel<true><<<grid, INTRA_WARPS * 32, 0, st>>>(refs, mode, pred, n, tab, !g_intraSwar);
    else intra32_kernel<false><<<grid, INTRA_WARPS * 32, 0, st>>>(refs, mode, pred, n, tab, !g_intraSwar);
    count_launch();
    return cudaGetLastError();
}

} // namespace x266
