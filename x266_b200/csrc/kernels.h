// kernels.h -- internal launcher prototypes (C++ side of libx266_b200; not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace x266 {

int sm_count();                 // SMs of the current device (cached per device)
// CTAs of `kernel` that are resident per SM at this block size (cudaOccupancyMaxActiveBlocksPerMultiprocessor, cached per kernel
// and device).  Persistent grids are sized with it: a grid larger than what is resident runs a second, under-filled wave (measured on
// the intra kernel: 8 CTAs/SM requested, 5 resident -> 8 % slower than 5 or 10).
int resident_ctas_per_sm(const void* kernel, int blockThreads, size_t dynSmemBytes);
// Stream-ordered scratch memory from a library-owned pool of the current device (release threshold = keep: a
// synchronisation between two calls must not hand the memory back to the driver and re-allocate it on the next call).
cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t st);
cudaError_t scratch_free(void* p, cudaStream_t st);
void count_launch();            // bumps the library-wide kernel launch counter

cudaError_t launch_dct32_bfly(const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2, cudaStream_t st);
cudaError_t launch_dct32_imma(const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2, cudaStream_t st);
cudaError_t launch_dct16_imma(const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2, cudaStream_t st);
cudaError_t launch_idct32_imma(const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2, cudaStream_t st);
cudaError_t launch_dct8_imma(const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2, cudaStream_t st);
void set_small_dct_cuda_cores(int on);   // tuning/diagnostic: CUDA-core dctN kernels for N=8,16 instead of IMMA
void set_imma_config(int id);   // tuning/diagnostic: selects a (warps, stages, CTAs/SM, staging) instantiation
void set_satd_cuda_cores(int on);   // tuning/diagnostic: CUDA-core SATD batch kernel instead of IMMA
void set_decide_v1(int on);     // tuning/diagnostic: CUDA-core intra decision kernel instead of the tensor-core one
void set_search_v1(int on);     // tuning/diagnostic: 0 = v3 (default), 1 = v1 (one CTA per block), 2/3 = v2 (strips)
cudaError_t launch_partial32(const int16_t* src, int16_t* dst, int shift, int line, cudaStream_t st);
cudaError_t launch_dctN(int log2n, const int16_t* src, int16_t* dst, size_t nBlocks, int s1, int s2, cudaStream_t st);

void set_frame_resi_config(int v);   // tuning/diagnostic: (prefetch depth, CTAs/SM) of the fused residual + DCT32 kernel
cudaError_t launch_frame_resi_dct32(const uint8_t* cur, const uint8_t* pred, int width, int height, int16_t* dst,
                                    int s1, int s2, cudaStream_t st);
cudaError_t launch_tiles_to_luma(const uint8_t* tiles, int width, int height, int pad, uint8_t* out, cudaStream_t st);
cudaError_t launch_transpose32(const uint8_t* src, uint8_t* dst, size_t nTiles, cudaStream_t st);
cudaError_t launch_conv_input_fmt(uint8_t* tiles, const uint8_t* Y, const uint8_t* U, const uint8_t* V, intptr_t strdY,
                                  int width, int height, cudaStream_t st);
cudaError_t launch_conv_output420(const uint8_t* tiles, uint8_t* Y, intptr_t strdY, uint8_t* U, uint8_t* V, intptr_t strdC,
                                  int width, int height, cudaStream_t st);
cudaError_t launch_satd8x8_batch(const int16_t* diff, int32_t* out, size_t n, cudaStream_t st);
cudaError_t launch_satd8x8_search(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, int h, int range,
                                  size_t blk0, size_t blk1, uint32_t* cost, int32_t* best, cudaStream_t st);
cudaError_t launch_satd8x8_search_v3(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, int h, int range,
                                     size_t blk0, size_t blk1, uint32_t* cost, int32_t* best, cudaStream_t st);
// the same searches with a 16-bit cost surface (exact: SATD <= 32640, SAD <= 16320 for 8-bit pixels)
cudaError_t launch_satd8x8_search(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, int h, int range,
                                  size_t blk0, size_t blk1, uint16_t* cost, int32_t* best, cudaStream_t st);
cudaError_t launch_satd8x8_search_v3(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, int h, int range,
                                     size_t blk0, size_t blk1, uint16_t* cost, int32_t* best, cudaStream_t st);
cudaError_t launch_sad8x8_search(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, int h, int range,
                                 size_t blk0, size_t blk1, uint16_t* cost, int32_t* best, cudaStream_t st);
void intra_mma_table_copy(uint32_t* out);   // 35 x 256 words: the per-mode MMA fragment table of the intra kernel (host copy)
void set_dct8_ctas(int v);        // tuning/diagnostic: CTAs per SM of the dct8 / dct4 persistent grids
void set_dct4_ctas(int v);
void set_intra_ctas(int v);       // tuning/diagnostic: CTAs per SM of the intra kernel's persistent grid
void set_intra_swar(int on);      // tuning/diagnostic: CUDA-core SWAR interpolation instead of the tensor-core angular path
void set_sad_search_v1(int on);   // tuning/diagnostic: first-generation SAD search (one CTA per block)
void set_search_acc_form(int f);  // tuning/diagnostic: accumulate form of the v3 search (satd_packed.h maxsum4)
cudaError_t launch_intra32_decide(const uint8_t* cur, const uint8_t* refs, uint32_t* cost, int32_t* bestMode, size_t n, cudaStream_t st);
cudaError_t launch_sad_region(const uint8_t* a, const uint8_t* b, size_t bytes, unsigned* out, cudaStream_t st);
cudaError_t launch_sad8x8_search(const uint8_t* cur, const uint8_t* refPad, intptr_t strd, int w, int h, int range,
                                 size_t blk0, size_t blk1, uint32_t* cost, int32_t* best, cudaStream_t st);
cudaError_t launch_intra32(const uint8_t* refs, const uint8_t* mode, uint8_t* pred, size_t n, cudaStream_t st);
// encode.cu: the closed intra block loop (decide -> predict -> residual -> DCT32 -> quant/dequant stub -> IDCT32 -> reconstruction)
cudaError_t launch_intra32_encode(const uint8_t* cur, const uint8_t* refs, size_t n, int qp, int16_t* level, uint8_t* recon,
                                  int32_t* bestMode, uint32_t* cost, cudaStream_t st);
cudaError_t launch_intra32_recon(const uint8_t* cur, const uint8_t* refs, const uint8_t* mode, size_t n, int qp, int16_t* level,
                                 uint8_t* recon, cudaStream_t st);
cudaError_t launch_quant_dequant(const int16_t* coef, int16_t* level, int16_t* dq, size_t nCoef, int qp, cudaStream_t st);
cudaError_t launch_intra32_modes(const uint8_t* refs, unsigned long long modeMask, uint8_t* pred, size_t nBlocks, cudaStream_t st);
const uint32_t* intra_mma_table_dev(cudaError_t* err);   // device copy of the fragment table on the current device (after intra_device_init)
cudaError_t intra_device_init();   // uploads the intra fragment table to the current device (called once per device by ffi.cu: ctx_get)
void intra_device_free();          // releases it (xGpuFree)
// debug (xGpuTune 15): *bad += number of mode[i] > maxMode
cudaError_t launch_mode_range_check(const uint8_t* mode, size_t n, int maxMode, unsigned* bad, cudaStream_t st);

} // namespace x266
