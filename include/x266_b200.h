/*
 * x266_b200.h -- C ABI of libx266_b200.so: the B200 (sm_100a) drop-in for the block-parallel encode
 * hot path of chenm001/x266 (32x32 integer DCT-II incl. transpose stage, 8x8 Hadamard SATD, 32x32
 * intra prediction).  Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * Reference citations are file:line in chenm001/x266 @ 379268c.
 *
 * Error convention (src/x266.cpp:494-513): int return, 0 = ok, -1 = failure; xGpuLastError() gives
 * the text.  There is NO CPU fallback anywhere in this library: with no usable CUDA device every
 * compute entry point fails (-1, or abort() for the void Tier-1/Tier-2 symbols).
 */
#ifndef X266_B200_H
#define X266_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ================================================================================================
 * Tier 1 -- the reference's own exported symbols, same names / signatures / semantics.
 * These are what the Bluesim BDPI runtime binds via `import "BDPI"` (src/mkDct32.bsv:409-411,
 * src/mkSatd.bsv:204-206) when the testbench is linked against src_tb/ *.c (build/Makefile:67).
 * State is module-global and not re-entrant, exactly like the reference.
 * ============================================================================================== */

/* replaces src_tb/dct32.c:30-64 -- same values, same layout */
#ifndef X266_B200_NO_GT32_DECL
extern const short g_t32[32][32];
#endif

/* replaces src_tb/dct32.c:178-202.  Draws 2048 rand() values exactly like the reference
 * ((rand()&0xFF) - (rand()&0xFF) per sample, row-major), then computes the 2-D transform with
 * shifts 4 / 11 ON THE GPU instead of the two host partialButterfly32 calls. */
void dct32_genNew(void);
/* replaces src_tb/dct32.c:205-220: rows lastDiff, lastDiff+1 as 32 little-endian u32 words. */
void dct32_getDiff(unsigned int res[/*32*/]);
/* replaces src_tb/dct32.c:223-246: 4 coefficients, column-major walk, packed into 64 bits. */
unsigned long long dct32_getDct(void);

/* replaces src_tb/satd.c:124-140 (stimulus as above; cost from the GPU kernel). */
void satd8x8_genNew(void);
/* replaces src_tb/satd.c:143-147 */
void satd8x8_getDiff(unsigned int res[/*4*/]);
/* replaces src_tb/satd.c:149-152 */
unsigned int satd8x8_getSatd(void);

/* ================================================================================================
 * Tier 2 -- the two arithmetic kernels, promoted from `static` to exported, HOST pointers,
 * synchronous, bit-exact.  These are the names a block loop in src/x266.cpp:537-546 would call.
 * ============================================================================================== */

/* replaces src_tb/dct32.c:66-170.  One 1-D pass over `line` rows of 32 with transposed store:
 * dst[k*line + j] = (int16)((sum_n g_t32[k][n]*src[j*32+n] + (1<<(shift-1))) >> shift). */
void partialButterfly32(const int16_t* src, int16_t* dst, int shift, int line);

/* replaces src_tb/satd.c:31-118 (int16-wrapping 8x8 Hadamard, (sum|.|+2)>>2). */
int satd8x8(const int16_t diff[64]);

/* ================================================================================================
 * Tier 3 -- batched entry points in the style of src/x266.cpp (x-prefixed, int 0/-1).
 * Host-pointer forms copy in/out through an internal chunked stream pipeline (csrc/ffi.cu: run_chunked).
 * Caller memory the DMA engines can address (cudaHostAlloc'd, or registered with xGpuHostRegister) is copied
 * directly.  PAGEABLE caller memory -- what the reference's caller allocates with _aligned_malloc,
 * src/x266.cpp:505,647-649 -- is staged through a library-owned pinned ring by a pool of host copy threads
 * (all batch entry points, the intra entry points and the search outputs; the few call-wide inputs -- search
 * planes, tiled frames, the Tier-2 single calls -- are handed to cudaMemcpyAsync, which stages them itself).
 * Each call leases its own streams, so host threads sharing a GPU overlap.  *Dev forms take device pointers on
 * the current device plus a cudaStream_t passed as void* (NULL = legacy default stream), enqueue only, and
 * never synchronise (the first call on a device initialises it: xGpuInit does the same ahead of time).
 * ============================================================================================== */

/* Mirrors xCodecInit/xCodecFree (src/x266.cpp:494-524).  device < 0 = current device.  Optional:
 * every entry point initialises lazily on the calling thread's current device. */
int  xGpuInit(int device);
void xGpuFree(void);
const char* xGpuLastError(void);
/* Optional, for callers that allocate their frame buffers once (as main() does, src/x266.cpp:647-649): page-lock an
 * existing allocation so every later host-pointer call DMAs it directly instead of staging it.  Unregister before free(). */
int  xGpuHostRegister(void* p, size_t bytes);
int  xGpuHostUnregister(void* p);
/* threads (incl. the caller) that copy pageable caller memory to and from the pinned ring (xGpuTune 13 / X266_HOST_COPY_THREADS) */
int  xGpuHostCopyThreads(void);
/* Number of kernels this library has launched since load (for the bench's gpu_launches claim). */
unsigned long long xGpuKernelLaunches(void);

/* DCT kernel variants */
enum {
    X266_DCT_AUTO  = 0,   /* best measured variant (IMMA) */
    X266_DCT_BFLY  = 1,   /* CUDA-core partial butterfly, one block per warp, smem transpose */
    X266_DCT_IMMA  = 2    /* int8 tensor-core (mma.sync m16n8k32) byte-plane dense product   */
};
int xGpuSetDctVariant(int variant);
/* Diagnostic/tuning hook (not part of the reference-facing surface; 0 / -1 = shipped default everywhere):
 * key 0 IMMA DCT32 instantiation (0 / 3 TMA rings, 6 direct loads = shipped, 7, 12; scripts/tune_dct.py) | 1 SATD search kernel (0 v3 packed
 * transform domain, 1 one CTA per block for any range) | 2 SATD batch (0 tensor cores + 3-stage ring, 1 CUDA cores, 5 4-stage ring) | 3 CUDA-core
 * DCT 4/8/16 | 4 blocks per chunk of the host-pointer DCT pipeline | 5 (unused) | 6 accumulate form of the v3 search | 7 first-generation SAD search |
 * 8 CUDA-core SWAR intra interpolation |
 * 9 / 10 / 11 CTAs per SM of the intra / DCT8 / DCT4 persistent grids (0 = shipped) | 12 pageable host buffers: 0 staged through the
 * pinned ring (shipped), 1 handed to the driver, 2 cudaHostRegister per call | 13 host copy threads (0 = X266_HOST_COPY_THREADS or
 * min(8, cpus/2)) | 14 non-temporal staging copies, bit 0 into the pinned slots, bit 1 into caller memory (3 = shipped) | 15 device-side mode[] range check in the *Dev intra entry points |
 * 16 fused residual + DCT32: 0 two blocks in flight per warp (shipped), 1 one. */
int xGpuTune(int key, int value);

/* 2-D forward 32x32 transform of nBlocks contiguous row-major int16 blocks:
 * dst = pass(shift2nd) o pass(shift1st), i.e. src_tb/dct32.c:197-198 on every block. */
int xDct32Batch(const int16_t* src, int16_t* dst, size_t nBlocks, int shift1st, int shift2nd);
int xDct32BatchDev(const int16_t* dSrc, int16_t* dDst, size_t nBlocks, int shift1st, int shift2nd, void* stream);

/* Multi-GPU form: blocks are independent, so GPU g of nGpus transforms the contiguous range
 * [g*N/G, (g+1)*N/G) of the host arrays; one host thread per device, no collective.  nGpus <= 0 = all visible. */
int xDct32BatchMultiGpu(const int16_t* src, int16_t* dst, size_t nBlocks, int shift1st, int shift2nd, int nGpus);

/* N x N forward transform, log2N in {2,3,4,5}; matrix rows g_t32[k*32/N][0..N) (dct32.c:109-143),
 * shifts per src/mkDct32.bsv:93-98. */
int xDctNBatch(int log2N, const int16_t* src, int16_t* dst, size_t nBlocks, int shift1st, int shift2nd);
int xDctNBatchDev(int log2N, const int16_t* dSrc, int16_t* dDst, size_t nBlocks, int shift1st, int shift2nd, void* stream);

/* Stand-alone transpose stage (src/mkTranspose.bsv:95-99 mkTranspose32x32 on Bit#(8)):
 * nTiles contiguous 32x32 byte tiles, dst[t][j][i] = src[t][i][j].  (The DCT kernels need no such step.) */
int xTranspose32x32Batch(const uint8_t* src, uint8_t* dst, size_t nTiles);
int xTranspose32x32BatchDev(const uint8_t* dSrc, uint8_t* dDst, size_t nTiles, void* stream);

/* Inverse 32x32 transform ("next" row N3; not in the reference C -- parity unpinned): vertical pass first,
 * tmp = clip16((G^T * coef + rnd) >> shift1st), out = clip16((tmp * G + rnd) >> shift2nd), clip16 = saturation,
 * i.e. the HEVC/VVC decoder order and the HM partialButterflyInverse32 arithmetic (shifts 7 / 12 for 8-bit). */
int xIdct32Batch(const int16_t* src, int16_t* dst, size_t nBlocks, int shift1st, int shift2nd);
int xIdct32BatchDev(const int16_t* dSrc, int16_t* dDst, size_t nBlocks, int shift1st, int shift2nd, void* stream);

/* one 1-D pass, device pointers (Tier-2 partialButterfly32 on resident data) */
int xPartialButterfly32Dev(const int16_t* dSrc, int16_t* dDst, int shift, int line, void* stream);

/* SATD of n contiguous 8x8 int16 difference blocks (128 B each) -> n int32 costs. */
int xSatd8x8Batch(const int16_t* diff, int32_t* satd, size_t n);
int xSatd8x8BatchDev(const int16_t* dDiff, int32_t* dSatd, size_t n, void* stream);

/* Full search.  cur: w x h u8, stride w (w,h multiples of 8).  refPadded: reference plane
 * edge-replicated by `range` pixels on every side, stride strd >= w + 2*range.  For the 8x8 blocks
 * [blk0, blk1) in raster order and every mv in [-range,range]^2:
 *   cost[(b-blk0)][(mvy+range)][(mvx+range)] = satd8x8(cur_block - ref_block(mv))       (u32)
 *   best[(b-blk0)][3] = {min cost, mvx, mvy}; ties -> smallest mvx^2+mvy^2, then raster order.
 * cost and/or best may be NULL. */
int xSatd8x8Search(const uint8_t* cur, const uint8_t* refPadded, intptr_t strd, int w, int h, int range,
                   size_t blk0, size_t blk1, uint32_t* cost, int32_t* best);
int xSatd8x8SearchDev(const uint8_t* dCur, const uint8_t* dRefPadded, intptr_t strd, int w, int h, int range,
                      size_t blk0, size_t blk1, uint32_t* dCost, int32_t* dBest, void* stream);

/* Plain SAD ("next" row N4).  sad() keeps the reference's name and signature
 * (riscv/programs/benchmarks/sad/sad.c:27-38: sum |a-b| over an n x n byte region; host pointers, synchronous).
 * xSad8x8Search is the integer-pel pre-filter with exactly the conventions of xSatd8x8Search (cost = SAD). */
int sad(unsigned char* input_data1, unsigned char* input_data2, size_t n);
int xSad8x8Search(const uint8_t* cur, const uint8_t* refPadded, intptr_t strd, int w, int h, int range,
                  size_t blk0, size_t blk1, uint32_t* cost, int32_t* best);
int xSad8x8SearchDev(const uint8_t* dCur, const uint8_t* dRefPadded, intptr_t strd, int w, int h, int range,
                     size_t blk0, size_t blk1, uint32_t* dCost, int32_t* dBest, void* stream);

/* The same searches with a 16-bit cost surface, cost[(b-blk0)][my][mx] as uint16_t -- exact, because an 8x8 SATD of 8-bit pixels cannot
 * exceed 32640 (sum |T_k| <= 8 ||T||_2 = 64 ||x||_2 <= 130560 before the >> 2 of src_tb/satd.c:113) and an 8x8 SAD cannot exceed 16320.
 * Halves the bytes of the surface (547.6 -> 273.8 MB per 1080p +-32 frame), which is all but 4 MB of the searches' HBM and host-link traffic. */
int xSatd8x8SearchU16(const uint8_t* cur, const uint8_t* refPadded, intptr_t strd, int w, int h, int range,
                      size_t blk0, size_t blk1, uint16_t* cost, int32_t* best);
int xSatd8x8SearchU16Dev(const uint8_t* dCur, const uint8_t* dRefPadded, intptr_t strd, int w, int h, int range,
                         size_t blk0, size_t blk1, uint16_t* dCost, int32_t* dBest, void* stream);
int xSad8x8SearchU16(const uint8_t* cur, const uint8_t* refPadded, intptr_t strd, int w, int h, int range,
                     size_t blk0, size_t blk1, uint16_t* cost, int32_t* best);
int xSad8x8SearchU16Dev(const uint8_t* dCur, const uint8_t* dRefPadded, intptr_t strd, int w, int h, int range,
                        size_t blk0, size_t blk1, uint16_t* dCost, int32_t* dBest, void* stream);

/* 32x32 intra prediction (src/mkIntra32-wip.bsv:34-48,61-397): n predictions; refs[i] = 64 left
 * pixels then 65 top pixels (corner first) = 129 bytes; mode[i] in 0..34 (0 planar, 1 DC, 2..34
 * angular); pred[i] = 32x32 u8 row-major. */
int xIntra32Pred(const uint8_t* refs, const uint8_t* mode, uint8_t* pred, size_t n);
int xIntra32PredDev(const uint8_t* dRefs, const uint8_t* dMode, uint8_t* dPred, size_t n, void* stream);

/* Mode-major form: for each of nBlocks blocks, the predictions of EVERY mode whose bit is set in modeMask (bit m = mode m, m in 0..34),
 * in ascending mode order: pred[b][j] = 32x32 u8, j-th set bit.  The block's 129 reference bytes are loaded and staged once for all its
 * modes (what a rate-distortion search asks for; xIntra32Decide is the fused form that never writes the predictions). */
int xIntra32PredModes(const uint8_t* refs, size_t nBlocks, uint64_t modeMask, uint8_t* pred);
int xIntra32PredModesDev(const uint8_t* dRefs, size_t nBlocks, uint64_t modeMask, uint8_t* dPred, void* stream);

/* Fused intra mode decision ("next" row N1; the RTL's Decide channel, src/mkIntra32-wip.bsv:39-48): for each of n
 * 32x32 blocks (cur[i] = 32x32 u8 row-major, refs[i] as for xIntra32Pred) and each mode m in 0..34,
 *   cost[i][m] = sum over the 16 8x8 sub-blocks of satd8x8(cur - pred_m)       (src_tb/satd.c:31-118)
 *   bestMode[i] = argmin_m cost[i][m] (ties -> lowest mode).  Prediction and residual never leave the SM. */
int xIntra32Decide(const uint8_t* cur, const uint8_t* refs, uint32_t* cost, int32_t* bestMode, size_t n);
int xIntra32DecideDev(const uint8_t* dCur, const uint8_t* dRefs, uint32_t* dCost, int32_t* dBestMode, size_t n, void* stream);
/* The closed intra block loop in ONE kernel ("next" rows N1 + N3; both channels of the RTL's intra unit, `Decide` and `Recon`,
 * src/mkIntra32-wip.bsv:39-48, around the transform of src_tb/dct32.c:197-198 and its inverse with the same matrix dct32.c:30-64).
 * For each of n 32x32 blocks (cur[i] = 32x32 u8 row-major, refs[i] as for xIntra32Pred):
 *   bestMode[i] = the xIntra32Decide decision (cost[i][35] as there; cost may be NULL)
 *   level[i]    = Q(DCT32(cur - pred_best)), forward shifts 4 / 11 (8-bit video), [32][32] int16
 *   recon[i]    = clip8(pred_best + IDCT32(Q^-1(level)))   with the inverse shifts 7 / 12, [32][32] u8
 * Prediction, residual and coefficients never leave the SM: 1 KiB + 129 B in, 2 KiB + 1 KiB + 4 B out per block.
 * Q is a STUB (the reference has no quantiser): the flat intra quantiser arithmetic of the HEVC/VVC test models for 8-bit video,
 *   level = sign(c) * min(32767, (|c| * qScale[qp%6] + (171 << (qBits-9))) >> qBits), qBits = 16 + qp/6, qScale = {26214,23302,20560,18396,16384,14564}
 *   c'    = clip16((level * (iqScale[qp%6] << (qp/6)) + 8) >> 4),                                       iqScale = {40,45,51,57,64,72};  qp in 0..51. */
int xIntra32EncodeBlock(const uint8_t* cur, const uint8_t* refs, size_t n, int qp, int16_t* level, uint8_t* recon,
                        int32_t* bestMode, uint32_t* cost);
int xIntra32EncodeBlockDev(const uint8_t* dCur, const uint8_t* dRefs, size_t n, int qp, int16_t* dLevel, uint8_t* dRecon,
                           int32_t* dBestMode, uint32_t* dCost, void* stream);
/* the `Recon` channel alone: mode[i] given (0..34) */
int xIntra32Recon(const uint8_t* cur, const uint8_t* refs, const uint8_t* mode, size_t n, int qp, int16_t* level, uint8_t* recon);
int xIntra32ReconDev(const uint8_t* dCur, const uint8_t* dRefs, const uint8_t* dMode, size_t n, int qp, int16_t* dLevel,
                     uint8_t* dRecon, void* stream);
/* the quantiser stub on its own: nCoef (multiple of 8) int16 coefficients -> levels and/or de-quantised coefficients (either may be NULL) */
int xQuantDequantDev(const int16_t* dCoef, int16_t* dLevel, int16_t* dDequant, size_t nCoef, int qp, void* stream);

/* Diagnostic (host only, needs no device): the per-mode MMA fragment table the intra kernel multiplies with, 35 x 256 words,
 * [mode][half][lane][4] -- A fragments of the weight matrix for vertical modes, B fragments of its transpose (output columns
 * permuted, x = 8(n>>1) + 2t + (n&1)) for horizontal modes.  tests/intra_mma_model.py replays the kernel's choreography with it. */
int xIntra32MmaTable(uint32_t* table /* [35 * 256] */);

/* ================================================================================================
 * "Next" rows: the encoder's tiled frame stores on the device.
 * A frame is a raster of 512-byte ref_block_t tiles (src/x266.cpp:56-63: m_Y[16*16] | m_C[2*8*8] | m_I[128]),
 * width/16 tiles per row (m_frames_strd, x266.cpp:503); passed here as void* / uint8_t*.
 * ============================================================================================== */

/* src/x266.cpp:55-63: the 512-byte tile of the encoder's frame stores (all members are bytes: no padding, PACKED or not) */
typedef struct _ref_block_t {
    uint8_t m_Y[16 * 16];         /* 256 bytes - Y  */
    uint8_t m_C[2 * 8 * 8];       /* 128 bytes - UV */
    uint8_t m_I[128];             /* 128 bytes - Info */
} ref_block_t;

/* replaces src/x266.cpp:415-453 and :455-492 with the reference's own names and signatures (HOST pointers, synchronous, void; the
 * reference asserts on width / height not being multiples of 16 -- this library prints the error and aborts).  These are the two calls
 * xEncodeFrame already makes (x266.cpp:537).  m_I is left untouched, chroma stride of xConvInputFmt is strdY >> 1 (:426). */
void xConvInputFmt(ref_block_t* pBlock, const uint8_t* inpY, const uint8_t* inpU, const uint8_t* inpV, const intptr_t strdY,
                   const int width, const int height);
void xConvOutput420(const ref_block_t* pBlock, uint8_t* outY, const intptr_t strdY, uint8_t* outU, uint8_t* outV, intptr_t strdC,
                    const int width, const int height);

/* replaces src/x266.cpp:415-453 (planar YUV 4:2:0 -> tiles; m_I untouched) on device-resident planes */
int xConvInputFmtDev(void* dTiles, const uint8_t* dY, const uint8_t* dU, const uint8_t* dV, intptr_t strdY,
                     int width, int height, void* stream);
/* replaces src/x266.cpp:455-492 (tiles -> planar) */
int xConvOutput420Dev(const void* dTiles, uint8_t* dY, intptr_t strdY, uint8_t* dU, uint8_t* dV, intptr_t strdC,
                      int width, int height, void* stream);

/* Full search straight on the encoder's frame stores: curTiles / refTiles are ref_block_t frames (m_frames[0] and a reference,
 * src/x266.cpp:99), width and height multiples of 16; the reference is edge-replicated by `range` pixels inside the call.
 * Outputs and conventions exactly as xSatd8x8Search / xSad8x8Search. */
int xSatd8x8SearchTiled(const void* curTiles, const void* refTiles, int width, int height, int range, size_t blk0, size_t blk1,
                        uint32_t* cost, int32_t* best);
int xSatd8x8SearchTiledDev(const void* dCurTiles, const void* dRefTiles, int width, int height, int range, size_t blk0, size_t blk1,
                           uint32_t* dCost, int32_t* dBest, void* stream);
int xSad8x8SearchTiledDev(const void* dCurTiles, const void* dRefTiles, int width, int height, int range, size_t blk0, size_t blk1,
                          uint32_t* dCost, int32_t* dBest, void* stream);
/* 16-bit cost surface (see xSatd8x8SearchU16) */
int xSatd8x8SearchTiledU16Dev(const void* dCurTiles, const void* dRefTiles, int width, int height, int range, size_t blk0, size_t blk1,
                              uint16_t* dCost, int32_t* dBest, void* stream);
int xSad8x8SearchTiledU16Dev(const void* dCurTiles, const void* dRefTiles, int width, int height, int range, size_t blk0, size_t blk1,
                             uint16_t* dCost, int32_t* dBest, void* stream);

/* Fused residual + transform for the block loop of xEncodeFrame (src/x266.cpp:537-546): for every 32x32 luma
 * block b (raster order, width and height multiples of 32) of the tiled frames cur and pred,
 *     coef[b] = DCT32(cur_b - pred_b)   with shifts s1/s2 (src_tb/dct32.c:197-198 on the int16 residual).
 * coef is [nBlocks][32][32] int16, the layout xDct32Batch uses.  The residual is never materialised. */
int xFrameResiDct32(const void* curTiles, const void* predTiles, int width, int height, int16_t* coef,
                    int shift1st, int shift2nd);
int xFrameResiDct32Dev(const void* dCurTiles, const void* dPredTiles, int width, int height, int16_t* dCoef,
                       int shift1st, int shift2nd, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* X266_B200_H */
