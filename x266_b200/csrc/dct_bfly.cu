// dct_bfly.cu -- CUDA-core partial-butterfly forward transforms (the north-star default mapping:
// one 32x32 transform block per warp, lane = row, transpose stage in swizzled shared memory).
//
// Reference behaviour: src_tb/dct32.c:66-170 (partialButterfly32) applied twice (dct32.c:197-198);
// the even/odd folding follows the E/O, EE/EO, EEE/EEO, EEEE/EEEO ladder of dct32.c:78-126 and the
// transpose stage is what src/mkTranspose.bsv / the BRAM bank scheme of src/mkDct32.bsv:162-210 do
// in hardware.  Arithmetic contract: 32-bit accumulate, +(1<<(shift-1)), arithmetic >>, truncating
// int16 store (dct32.c:128-151).
#include "common.cuh"
#include "kernels.h"

namespace x266 {

// ------------------------------------------------------------------------------------------------
// 1-D N-point transform on registers.  y[k*STEP] = sum_n G_N[k][n] x[n], G_N[k][n] = g32(k*32/N, n).
// Even outputs recurse on the folded sums, odd outputs are a dense N/2 dot product on the folded
// differences; every coefficient is an immediate after full unrolling.
// ------------------------------------------------------------------------------------------------
template <int N, int STEP, int NTOP>
struct Dct1D {
    static __device__ __forceinline__ void run(const int (&x)[N], int (&y)[NTOP])
    {
        int e[N / 2], o[N / 2];
#pragma unroll
        for (int m = 0; m < N / 2; m++) {
            e[m] = x[m] + x[N - 1 - m];
            o[m] = x[m] - x[N - 1 - m];
        }
        Dct1D<N / 2, STEP * 2, NTOP>::run(e, y);
#pragma unroll
        for (int k = 1; k < N; k += 2) {
            int acc = 0;
#pragma unroll
            for (int m = 0; m < N / 2; m++) acc += g32(k * (32 / N), m) * o[m];
            y[k * STEP] = acc;
        }
    }
};

template <int STEP, int NTOP>
struct Dct1D<1, STEP, NTOP> {
    static __device__ __forceinline__ void run(const int (&x)[1], int (&y)[NTOP]) { y[0] = 64 * x[0]; }
};

__device__ __forceinline__ void unpack8(const uint4& v, int* x)
{
    x[0] = (int)(short)(v.x & 0xFFFF); x[1] = (int)v.x >> 16;
    x[2] = (int)(short)(v.y & 0xFFFF); x[3] = (int)v.y >> 16;
    x[4] = (int)(short)(v.z & 0xFFFF); x[5] = (int)v.z >> 16;
    x[6] = (int)(short)(v.w & 0xFFFF); x[7] = (int)v.w >> 16;
}

// 32x32 int16 tile in shared memory, 64-byte rows, 16-byte chunks XOR-swizzled by (row>>1)&3 so that
// both "every lane reads its own row" (128-bit) and "every lane writes one column element" (16-bit)
// are bank-conflict free.
__device__ __forceinline__ uint32_t tile_chunk(uint32_t base, int row, int chunk)
{
    return base + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
}

constexpr int BFLY_WARPS = 8;

__global__ void __launch_bounds__(BFLY_WARPS * 32)
dct32_bfly_kernel(const int16_t* __restrict__ src, int16_t* __restrict__ dst, size_t nBlocks, int shift1, int shift2)
{
    __shared__ __align__(128) uint8_t sm[BFLY_WARPS][2][2048];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t buf0 = smem_u32(&sm[warp][0][0]);
    const uint32_t buf1 = smem_u32(&sm[warp][1][0]);
    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);

    for (size_t b = (size_t)blockIdx.x * BFLY_WARPS + warp; b < nBlocks; b += (size_t)gridDim.x * BFLY_WARPS) {
        const int16_t* s = src + b * 1024;
        int16_t* d = dst + b * 1024;

        // coalesced 128-bit loads of the 2 KiB block, swizzled into buf0
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int g = i * 32 + lane;
            st_shared_v4(tile_chunk(buf0, g >> 2, g & 3), ld_global_stream(s + g * 8));
        }
        __syncwarp();

        int x[32], y[32];
        // ---- pass 1: lane = source row j ----------------------------------------------------
#pragma unroll
        for (int c = 0; c < 4; c++) unpack8(ld_shared_v4(tile_chunk(buf0, lane, c)), &x[8 * c]);
        Dct1D<32, 1, 32>::run(x, y);
        // transposed store: coef[k][j] -> buf1 row k, column j = lane   (dct32.c: dst[k*line + j])
#pragma unroll
        for (int k = 0; k < 32; k++) {
            const short v = (short)((y[k] + add1) >> shift1);
            const uint32_t a = tile_chunk(buf1, k, lane >> 3) + (lane & 7) * 2;
            asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "h"(v) : "memory");
        }
        __syncwarp();

        // ---- pass 2: lane = coefficient row k ----------------------------------------------
#pragma unroll
        for (int c = 0; c < 4; c++) unpack8(ld_shared_v4(tile_chunk(buf1, lane, c)), &x[8 * c]);
        Dct1D<32, 1, 32>::run(x, y);
#pragma unroll
        for (int k = 0; k < 32; k++) {
            const short v = (short)((y[k] + add2) >> shift2);
            const uint32_t a = tile_chunk(buf0, k, lane >> 3) + (lane & 7) * 2;
            asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "h"(v) : "memory");
        }
        __syncwarp();

        // coalesced 128-bit stores
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int g = i * 32 + lane;
            st_global_stream(d + g * 8, ld_shared_v4(tile_chunk(buf0, g >> 2, g & 3)));
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// One 1-D pass over `line` rows of 32 with transposed store (Tier-2 partialButterfly32 semantics for
// any `line`).  One thread per row; stores dst[k*line + j] are coalesced over j.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
partial32_kernel(const int16_t* __restrict__ src, int16_t* __restrict__ dst, int shift, int line)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= line) return;
    const int add = 1 << (shift - 1);
    int x[32], y[32];
    const int16_t* s = src + (size_t)j * 32;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
        for (int c = 0; c < 4; c++) unpack8(*reinterpret_cast<const uint4*>(s + 8 * c), &x[8 * c]);
    } else {
#pragma unroll
        for (int n = 0; n < 32; n++) x[n] = s[n];
    }
    Dct1D<32, 1, 32>::run(x, y);
#pragma unroll
    for (int k = 0; k < 32; k++) dst[(size_t)k * line + j] = (short)((y[k] + add) >> shift);
}

// ------------------------------------------------------------------------------------------------
// N x N blocks, N in {4,8,16}: each warp stages 1024 samples (2 KiB = 64/16/4 blocks), lanes take
// rows round-robin, transpose through shared memory.  Matrix rows g32(k*32/N, .) and shifts per
// src/mkDct32.bsv:93-98 (caller supplies the shifts).
// ------------------------------------------------------------------------------------------------
constexpr int DCTN_WARPS = 8;

template <int LOG2N>
__global__ void __launch_bounds__(DCTN_WARPS * 32)
dctN_kernel(const int16_t* __restrict__ src, int16_t* __restrict__ dst, size_t nSamples, int shift1, int shift2)
{
    constexpr int N = 1 << LOG2N;
    constexpr int ROWS_PER_LANE = (1024 / N) / 32;
    __shared__ __align__(16) int16_t sm[DCTN_WARPS][2][1024];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int16_t* s0 = sm[warp][0];
    int16_t* s1 = sm[warp][1];
    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);
    const size_t nChunks = (nSamples + 1023) / 1024;

    for (size_t ch = (size_t)blockIdx.x * DCTN_WARPS + warp; ch < nChunks; ch += (size_t)gridDim.x * DCTN_WARPS) {
        const size_t base = ch * 1024;
        const int valid = (int)((nSamples - base) < 1024 ? (nSamples - base) : 1024);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int idx = (i * 32 + lane) * 8;
            if (idx < valid) *reinterpret_cast<uint4*>(s0 + idx) = ld_global_stream(src + base + idx);
        }
        __syncwarp();
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
            const int16_t* in = pass ? s1 : s0;
            int16_t* out = pass ? s0 : s1;
            const int add = pass ? add2 : add1, shift = pass ? shift2 : shift1;
#pragma unroll
            for (int rr = 0; rr < ROWS_PER_LANE; rr++) {
                const int r = rr * 32 + lane;
                if (r * N < valid) {
                    const int blk = r >> LOG2N, j = r & (N - 1);
                    int x[N], y[N];
                    const uint32_t* w = reinterpret_cast<const uint32_t*>(in + r * N);
#pragma unroll
                    for (int n = 0; n < N / 2; n++) {
                        const uint32_t v = w[n];
                        x[2 * n] = (int)(short)(v & 0xFFFF);
                        x[2 * n + 1] = (int)v >> 16;
                    }
                    Dct1D<N, 1, N>::run(x, y);
#pragma unroll
                    for (int k = 0; k < N; k++) out[blk * N * N + k * N + j] = (short)((y[k] + add) >> shift);
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int idx = (i * 32 + lane) * 8;
            if (idx < valid) st_global_stream(dst + base + idx, *reinterpret_cast<const uint4*>(s0 + idx));
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// 4x4 blocks, fully coalesced variant: a warp moves 2 KiB (64 blocks) per iteration with four 512-byte
// load instructions; lane pairs then swap 16-byte halves (8 SHFL) so that every lane owns two whole blocks,
// transforms them in registers, swaps the halves back (8 SHFL) and stores with four 512-byte instructions.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dct4_block(const uint32_t (&w)[8], uint32_t (&o)[8], int add1, int shift1, int add2, int shift2)
{
    int c[4][4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int x[4], y[4];
        x[0] = (int)(short)(w[2 * j] & 0xFFFF); x[1] = (int)w[2 * j] >> 16;
        x[2] = (int)(short)(w[2 * j + 1] & 0xFFFF); x[3] = (int)w[2 * j + 1] >> 16;
        Dct1D<4, 1, 4>::run(x, y);
#pragma unroll
        for (int k = 0; k < 4; k++) c[k][j] = (int)(short)((y[k] + add1) >> shift1);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int y[4];
        Dct1D<4, 1, 4>::run(c[k], y);
#pragma unroll
        for (int k2 = 0; k2 < 4; k2++) c[k][k2] = (y[k2] + add2) >> shift2;
    }
#pragma unroll
    for (int k2 = 0; k2 < 4; k2++) {
        o[2 * k2] = prmt((uint32_t)c[0][k2], (uint32_t)c[1][k2], 0x5410);
        o[2 * k2 + 1] = prmt((uint32_t)c[2][k2], (uint32_t)c[3][k2], 0x5410);
    }
}

__device__ __forceinline__ uint4 shfl_xor1(uint4 v)
{
    v.x = __shfl_xor_sync(0xffffffffu, v.x, 1); v.y = __shfl_xor_sync(0xffffffffu, v.y, 1);
    v.z = __shfl_xor_sync(0xffffffffu, v.z, 1); v.w = __shfl_xor_sync(0xffffffffu, v.w, 1);
    return v;
}

constexpr int DCT4X_WARPS = 8;

__global__ void __launch_bounds__(DCT4X_WARPS * 32)
dct4_xchg_kernel(const int16_t* __restrict__ src, int16_t* __restrict__ dst, size_t nBlocks, int shift1, int shift2)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool odd = lane & 1;
    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);
    const size_t nPieces = nBlocks * 2;                       // 16-byte pieces
    const size_t nUnits = (nPieces + 127) / 128;
    for (size_t u = (size_t)blockIdx.x * DCT4X_WARPS + warp; u < nUnits; u += (size_t)gridDim.x * DCT4X_WARPS) {
        const size_t base = u * 128;                          // piece index of this unit
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            size_t pc = base + j * 32 + lane;
            pc = pc < nPieces ? pc : nPieces - 1;             // ragged tail: clamp (never stored)
            v[j] = ld_global_stream(src + pc * 8);
        }
        // even lanes own the blocks of j = 0,1, odd lanes those of j = 2,3
        const uint4 r0 = shfl_xor1(odd ? v[0] : v[2]);
        const uint4 r1 = shfl_xor1(odd ? v[1] : v[3]);
        const uint4 a0 = odd ? r0 : v[0], a1 = odd ? v[2] : r0;       // block A: halves 0,1
        const uint4 b0 = odd ? r1 : v[1], b1 = odd ? v[3] : r1;       // block B
        const uint32_t wa[8] = { a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w };
        const uint32_t wb[8] = { b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w };
        uint32_t oa[8], ob[8];
        dct4_block(wa, oa, add1, shift1, add2, shift2);
        dct4_block(wb, ob, add1, shift1, add2, shift2);
        const uint4 A0 = make_uint4(oa[0], oa[1], oa[2], oa[3]), A1 = make_uint4(oa[4], oa[5], oa[6], oa[7]);
        const uint4 B0 = make_uint4(ob[0], ob[1], ob[2], ob[3]), B1 = make_uint4(ob[4], ob[5], ob[6], ob[7]);
        // even lane keeps halves 0 of its blocks and receives the odd lane's halves 0; odd keeps halves 1
        const uint4 s0 = shfl_xor1(odd ? A0 : A1);
        const uint4 s1 = shfl_xor1(odd ? B0 : B1);
        uint4 o[4];
        o[0] = odd ? s0 : A0;  o[1] = odd ? s1 : B0;  o[2] = odd ? A1 : s0;  o[3] = odd ? B1 : s1;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const size_t pc = base + j * 32 + lane;
            if (pc < nPieces) st_global_stream(dst + pc * 8, o[j]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static std::atomic<int> g_dct4Ctas{0};       // tuning/diagnostic: CTAs per SM of dct4_xchg's persistent grid (0 = 12: measured 4/6/8/10/12/16 ->
                                 // 5829/5681/6124/6065/6357/6338 GB/s on 1 Gi samples, scripts/time_small_dct_grid.py)
void set_dct4_ctas(int v) { g_dct4Ctas = v; }

static int grid_for(size_t units, int unitsPerCta, int ctasPerSm)
{
    size_t want = (units + unitsPerCta - 1) / unitsPerCta;
    size_t cap = (size_t)sm_count() * ctasPerSm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

cudaError_t launch_dct32_bfly(const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2, cudaStream_t st)
{
    if (nBlocks == 0) return cudaSuccess;
    dct32_bfly_kernel<<<grid_for(nBlocks, BFLY_WARPS, 4), BFLY_WARPS * 32, 0, st>>>(src, dst, nBlocks, s1, s2);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_partial32(const int16_t* src, int16_t* dst, int shift, int line, cudaStream_t st)
{
    if (line <= 0) return cudaSuccess;
    partial32_kernel<<<(line + 127) / 128, 128, 0, st>>>(src, dst, shift, line);
    count_launch();
    return cudaGetLastError();
}

static std::atomic<int> g_smallCuda{0};     // 0: shipped (IMMA for N=8,16; lane-exchange kernel for N=4); 1: smem-staged CUDA-core dctN kernels
void set_small_dct_cuda_cores(int on) { g_smallCuda = on; }

cudaError_t launch_dctN(int log2n, const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2, cudaStream_t st)
{
    if (nBlocks == 0) return cudaSuccess;
    if (log2n == 4 && g_smallCuda != 1) return launch_dct16_imma(src, dst, nBlocks, s1, s2, st);
    if (log2n == 3 && g_smallCuda != 1) return launch_dct8_imma(src, dst, nBlocks, s1, s2, st);
    if (log2n == 2 && g_smallCuda != 1) {
        dct4_xchg_kernel<<<grid_for((nBlocks + 63) / 64, DCT4X_WARPS, g_dct4Ctas > 0 ? g_dct4Ctas.load() : 12), DCT4X_WARPS * 32, 0, st>>>(src, dst, nBlocks, s1, s2);
        count_launch();
        return cudaGetLastError();
    }
    const size_t nSamples = nBlocks << (2 * log2n);
    const int grid = grid_for((nSamples + 1023) / 1024, DCTN_WARPS, 4);
    switch (log2n) {
    case 2: dctN_kernel<2><<<grid, DCTN_WARPS * 32, 0, st>>>(src, dst, nSamples, s1, s2); break;
    case 3: dctN_kernel<3><<<grid, DCTN_WARPS * 32, 0, st>>>(src, dst, nSamples, s1, s2); break;
    case 4: dctN_kernel<4><<<grid, DCTN_WARPS * 32, 0, st>>>(src, dst, nSamples, s1, s2); break;
    default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

} // namespace x266
