"""ctypes binding of libx266_b200.so.  Host-pointer calls take numpy arrays; *Dev calls take raw device
pointers (ints, e.g. torch.Tensor.data_ptr()) and a stream handle (int, e.g.
torch.cuda.current_stream().cuda_stream)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("X266_B200_LIB", os.path.join(HERE, "libx266_b200.so"))       # override: A/B runs against another build

DCT_AUTO, DCT_BFLY, DCT_IMMA = 0, 1, 2


class X266Error(RuntimeError):
    pass


def build(verbose=False):
    """Compile every CUDA source for sm_100a into x266_b200/libx266_b200.so (nvcc cross-compiles
    without a GPU)."""
    out = subprocess.run(["make", "-C", os.path.join(HERE, "csrc")], capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode:
        raise X266Error("nvcc build of libx266_b200.so failed")
    return LIB_PATH


_lib = None


def _set(L, name, attr, value):
    """prototype of one entry point; a missing symbol is an error unless an older build was loaded on purpose (X266_B200_LIB)"""
    try:
        setattr(getattr(L, name), attr, value)
    except AttributeError:
        if "X266_B200_LIB" not in os.environ:
            raise


def lib():
    """Load the CUDA library.  Raises if it has not been built -- no fallback of any kind."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise X266Error(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(the x266_b200 hot path has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, sz, i = C.c_void_p, C.c_size_t, C.c_int
    _set(L, "xGpuInit", "argtypes", [i])
    _set(L, "xGpuLastError", "restype", C.c_char_p)
    _set(L, "xGpuKernelLaunches", "restype", C.c_ulonglong)
    _set(L, "xGpuSetDctVariant", "argtypes", [i])
    _set(L, "xGpuTune", "argtypes", [i, i])
    _set(L, "xGpuHostRegister", "argtypes", [vp, sz])
    _set(L, "xGpuHostUnregister", "argtypes", [vp])
    _set(L, "xIntra32MmaTable", "argtypes", [vp])
    _set(L, "xDct32Batch", "argtypes", [vp, vp, sz, i, i])
    _set(L, "xDct32BatchDev", "argtypes", [vp, vp, sz, i, i, vp])
    _set(L, "xDct32BatchMultiGpu", "argtypes", [vp, vp, sz, i, i, i])
    _set(L, "xIdct32Batch", "argtypes", [vp, vp, sz, i, i])
    _set(L, "xIdct32BatchDev", "argtypes", [vp, vp, sz, i, i, vp])
    _set(L, "xDctNBatch", "argtypes", [i, vp, vp, sz, i, i])
    _set(L, "xDctNBatchDev", "argtypes", [i, vp, vp, sz, i, i, vp])
    _set(L, "xPartialButterfly32Dev", "argtypes", [vp, vp, i, i, vp])
    _set(L, "xSatd8x8Batch", "argtypes", [vp, vp, sz])
    _set(L, "xSatd8x8BatchDev", "argtypes", [vp, vp, sz, vp])
    _set(L, "xSatd8x8Search", "argtypes", [vp, vp, C.c_ssize_t, i, i, i, sz, sz, vp, vp])
    _set(L, "xSatd8x8SearchDev", "argtypes", [vp, vp, C.c_ssize_t, i, i, i, sz, sz, vp, vp, vp])
    _set(L, "xIntra32Pred", "argtypes", [vp, vp, vp, sz])
    _set(L, "xIntra32PredDev", "argtypes", [vp, vp, vp, sz, vp])
    _set(L, "xConvInputFmt", "argtypes", [vp, vp, vp, vp, C.c_ssize_t, i, i])
    _set(L, "xConvInputFmt", "restype", None)
    _set(L, "xConvOutput420", "argtypes", [vp, vp, C.c_ssize_t, vp, vp, C.c_ssize_t, i, i])
    _set(L, "xConvOutput420", "restype", None)
    _set(L, "xSatd8x8SearchTiled", "argtypes", [vp, vp, i, i, i, sz, sz, vp, vp])
    _set(L, "xSatd8x8SearchTiledDev", "argtypes", [vp, vp, i, i, i, sz, sz, vp, vp, vp])
    _set(L, "xSad8x8SearchTiledDev", "argtypes", [vp, vp, i, i, i, sz, sz, vp, vp, vp])
    _set(L, "xConvInputFmtDev", "argtypes", [vp, vp, vp, vp, C.c_ssize_t, i, i, vp])
    _set(L, "xConvOutput420Dev", "argtypes", [vp, vp, C.c_ssize_t, vp, vp, C.c_ssize_t, i, i, vp])
    _set(L, "xFrameResiDct32", "argtypes", [vp, vp, i, i, vp, i, i])
    _set(L, "xFrameResiDct32Dev", "argtypes", [vp, vp, i, i, vp, i, i, vp])
    _set(L, "xIntra32Decide", "argtypes", [vp, vp, vp, vp, sz])
    _set(L, "xIntra32DecideDev", "argtypes", [vp, vp, vp, vp, sz, vp])
    _set(L, "xIntra32PredModes", "argtypes", [vp, sz, C.c_uint64, vp])
    _set(L, "xIntra32PredModesDev", "argtypes", [vp, sz, C.c_uint64, vp, vp])
    _set(L, "xIntra32EncodeBlock", "argtypes", [vp, vp, sz, i, vp, vp, vp, vp])
    _set(L, "xIntra32EncodeBlockDev", "argtypes", [vp, vp, sz, i, vp, vp, vp, vp, vp])
    _set(L, "xIntra32Recon", "argtypes", [vp, vp, vp, sz, i, vp, vp])
    _set(L, "xIntra32ReconDev", "argtypes", [vp, vp, vp, sz, i, vp, vp, vp])
    _set(L, "xQuantDequantDev", "argtypes", [vp, vp, vp, sz, i, vp])
    _set(L, "xTranspose32x32Batch", "argtypes", [vp, vp, sz])
    _set(L, "xTranspose32x32BatchDev", "argtypes", [vp, vp, sz, vp])
    _set(L, "sad", "argtypes", [vp, vp, sz])
    _set(L, "sad", "restype", i)
    _set(L, "xSad8x8Search", "argtypes", [vp, vp, C.c_ssize_t, i, i, i, sz, sz, vp, vp])
    _set(L, "xSad8x8SearchDev", "argtypes", [vp, vp, C.c_ssize_t, i, i, i, sz, sz, vp, vp, vp])
    for nm in ("xSatd8x8SearchU16", "xSad8x8SearchU16"):
        _set(L, nm, "argtypes", [vp, vp, C.c_ssize_t, i, i, i, sz, sz, vp, vp])
        _set(L, nm + "Dev", "argtypes", [vp, vp, C.c_ssize_t, i, i, i, sz, sz, vp, vp, vp])
    _set(L, "xSatd8x8SearchTiledU16Dev", "argtypes", [vp, vp, i, i, i, sz, sz, vp, vp, vp])
    _set(L, "xSad8x8SearchTiledU16Dev", "argtypes", [vp, vp, i, i, i, sz, sz, vp, vp, vp])
    _set(L, "partialButterfly32", "argtypes", [vp, vp, i, i])
    _set(L, "partialButterfly32", "restype", None)
    _set(L, "satd8x8", "argtypes", [vp])
    _set(L, "satd8x8", "restype", i)
    _set(L, "dct32_getDct", "restype", C.c_ulonglong)
    _set(L, "satd8x8_getSatd", "restype", C.c_uint)
    _lib = L
    return L


def last_error():
    return lib().xGpuLastError().decode()


def kernel_launches():
    return int(lib().xGpuKernelLaunches())


def set_dct_variant(v):
    _ck(lib().xGpuSetDctVariant(v), "xGpuSetDctVariant")


def xIntra32MmaTable():
    """host-only: the MMA fragment table of the intra kernel, [35][2][32][4] uint32"""
    t = np.zeros((35, 2, 32, 4), np.uint32)
    _ck(lib().xIntra32MmaTable(t.ctypes.data), "xIntra32MmaTable")
    return t


def host_register(a):
    """page-lock an existing numpy array (xGpuHostRegister); pair with host_unregister before it is freed"""
    _ck(lib().xGpuHostRegister(a.ctypes.data, a.nbytes), "xGpuHostRegister")


def host_copy_threads():
    return int(lib().xGpuHostCopyThreads())


def host_unregister(a):
    _ck(lib().xGpuHostUnregister(a.ctypes.data), "xGpuHostUnregister")


def tune(key, value):
    _ck(lib().xGpuTune(key, value), "xGpuTune")


def _ck(rc, what):
    if rc != 0:
        raise X266Error(f"{what} failed: {last_error()}")


def _np(a, dtype):
    a = np.ascontiguousarray(a, dtype)
    return a


def g_t32():
    return np.ctypeslib.as_array((C.c_int16 * 1024).in_dll(lib(), "g_t32")).reshape(32, 32).copy()


# ---- Tier 2 (reference names) ---------------------------------------------------------------------
def partialButterfly32(src, shift, line):
    src = _np(src, np.int16)
    assert src.size >= line * 32
    dst = np.zeros(line * 32, np.int16)
    lib().partialButterfly32(src.ctypes.data, dst.ctypes.data, shift, line)
    return dst


def satd8x8(diff):
    diff = _np(diff, np.int16)
    assert diff.size == 64
    return int(lib().satd8x8(diff.ctypes.data))


# ---- Tier 3, host pointers -----------------------------------------------------------------------
def xDct32Batch(src, shift1st=4, shift2nd=11, out=None):
    src = _np(src, np.int16)
    assert src.size % 1024 == 0
    dst = np.empty_like(src) if out is None else out
    _ck(lib().xDct32Batch(src.ctypes.data, dst.ctypes.data, src.size // 1024, shift1st, shift2nd), "xDct32Batch")
    return dst


def xDct32BatchMultiGpu(src, shift1st=4, shift2nd=11, n_gpus=0, out=None):
    src = _np(src, np.int16)
    assert src.size % 1024 == 0
    dst = np.empty_like(src) if out is None else out
    _ck(lib().xDct32BatchMultiGpu(src.ctypes.data, dst.ctypes.data, src.size // 1024, shift1st, shift2nd, n_gpus), "xDct32BatchMultiGpu")
    return dst


def xIdct32Batch(src, shift1st=7, shift2nd=12):
    src = _np(src, np.int16)
    assert src.size % 1024 == 0
    dst = np.empty_like(src)
    _ck(lib().xIdct32Batch(src.ctypes.data, dst.ctypes.data, src.size // 1024, shift1st, shift2nd), "xIdct32Batch")
    return dst


def xIdct32BatchDev(d_src, d_dst, n_blocks, shift1st, shift2nd, stream=0):
    _ck(lib().xIdct32BatchDev(d_src, d_dst, n_blocks, shift1st, shift2nd, stream), "xIdct32BatchDev")


def xDctNBatch(log2n, src, shift1st, shift2nd):
    src = _np(src, np.int16)
    bs = 1 << (2 * log2n)
    assert src.size % bs == 0
    dst = np.empty_like(src)
    _ck(lib().xDctNBatch(log2n, src.ctypes.data, dst.ctypes.data, src.size // bs, shift1st, shift2nd), "xDctNBatch")
    return dst


def xSatd8x8Batch(diff):
    diff = _np(diff, np.int16)
    assert diff.size % 64 == 0
    out = np.empty(diff.size // 64, np.int32)
    _ck(lib().xSatd8x8Batch(diff.ctypes.data, out.ctypes.data, out.size), "xSatd8x8Batch")
    return out


def xSatd8x8Search(cur, ref_padded, rng, blk0=0, blk1=None, want_cost=True, want_best=True, u16=False):
    """u16=True: xSatd8x8SearchU16 (16-bit cost surface)"""
    cur = _np(cur, np.uint8)
    ref_padded = _np(ref_padded, np.uint8)
    h, w = cur.shape
    assert ref_padded.shape == (h + 2 * rng, w + 2 * rng)
    if blk1 is None:
        blk1 = (w // 8) * (h // 8)
    side = 2 * rng + 1
    nb = blk1 - blk0
    cost = np.empty((nb, side, side), np.uint16 if u16 else np.uint32) if want_cost else None
    best = np.empty((nb, 3), np.int32) if want_best else None
    fn = lib().xSatd8x8SearchU16 if u16 else lib().xSatd8x8Search
    _ck(fn(cur.ctypes.data, ref_padded.ctypes.data, ref_padded.shape[1], w, h, rng, blk0, blk1,
           cost.ctypes.data if want_cost else None, best.ctypes.data if want_best else None),
        "xSatd8x8SearchU16" if u16 else "xSatd8x8Search")
    return cost, best


def xTranspose32x32Batch(src):
    src = _np(src, np.uint8)
    assert src.size % 1024 == 0
    dst = np.empty_like(src)
    _ck(lib().xTranspose32x32Batch(src.ctypes.data, dst.ctypes.data, src.size // 1024), "xTranspose32x32Batch")
    return dst


def xTranspose32x32BatchDev(d_src, d_dst, n_tiles, stream=0):
    _ck(lib().xTranspose32x32BatchDev(d_src, d_dst, n_tiles, stream), "xTranspose32x32BatchDev")


def xIntra32Decide(cur, refs):
    cur = _np(cur, np.uint8).reshape(-1, 1024)
    refs = _np(refs, np.uint8).reshape(-1, 129)
    assert cur.shape[0] == refs.shape[0]
    cost = np.empty((cur.shape[0], 35), np.uint32)
    best = np.empty(cur.shape[0], np.int32)
    _ck(lib().xIntra32Decide(cur.ctypes.data, refs.ctypes.data, cost.ctypes.data, best.ctypes.data, cur.shape[0]), "xIntra32Decide")
    return cost, best


def xIntra32PredModes(refs, mode_mask=(1 << 35) - 1):
    """every mode of mode_mask for every block: [nBlocks][popcount(mask)][32][32]"""
    refs = _np(refs, np.uint8).reshape(-1, 129)
    nm = bin(mode_mask).count("1")
    pred = np.empty((refs.shape[0], nm, 32, 32), np.uint8)
    _ck(lib().xIntra32PredModes(refs.ctypes.data, refs.shape[0], mode_mask, pred.ctypes.data), "xIntra32PredModes")
    return pred


def xIntra32PredModesDev(d_refs, n_blocks, mode_mask, d_pred, stream=0):
    _ck(lib().xIntra32PredModesDev(d_refs, n_blocks, mode_mask, d_pred, stream), "xIntra32PredModesDev")


def xIntra32EncodeBlock(cur, refs, qp, want_cost=True):
    """fused decide -> predict -> residual -> DCT32 -> quant/dequant -> IDCT32 -> recon: (level, recon, bestMode, cost)"""
    cur = _np(cur, np.uint8).reshape(-1, 1024)
    refs = _np(refs, np.uint8).reshape(-1, 129)
    n = cur.shape[0]
    assert refs.shape[0] == n
    level = np.empty((n, 32, 32), np.int16)
    recon = np.empty((n, 32, 32), np.uint8)
    best = np.empty(n, np.int32)
    cost = np.empty((n, 35), np.uint32) if want_cost else None
    _ck(lib().xIntra32EncodeBlock(cur.ctypes.data, refs.ctypes.data, n, qp, level.ctypes.data, recon.ctypes.data, best.ctypes.data,
                                  cost.ctypes.data if want_cost else None), "xIntra32EncodeBlock")
    return level, recon, best, cost


def xIntra32EncodeBlockDev(d_cur, d_refs, n, qp, d_level, d_recon, d_best, d_cost=0, stream=0):
    _ck(lib().xIntra32EncodeBlockDev(d_cur, d_refs, n, qp, d_level, d_recon, d_best, d_cost, stream), "xIntra32EncodeBlockDev")


def xIntra32Recon(cur, refs, modes, qp):
    cur = _np(cur, np.uint8).reshape(-1, 1024)
    refs = _np(refs, np.uint8).reshape(-1, 129)
    modes = _np(modes, np.uint8).ravel()
    n = cur.shape[0]
    assert refs.shape[0] == n == modes.size
    level = np.empty((n, 32, 32), np.int16)
    recon = np.empty((n, 32, 32), np.uint8)
    _ck(lib().xIntra32Recon(cur.ctypes.data, refs.ctypes.data, modes.ctypes.data, n, qp, level.ctypes.data, recon.ctypes.data), "xIntra32Recon")
    return level, recon


def xIntra32ReconDev(d_cur, d_refs, d_modes, n, qp, d_level, d_recon, stream=0):
    _ck(lib().xIntra32ReconDev(d_cur, d_refs, d_modes, n, qp, d_level, d_recon, stream), "xIntra32ReconDev")


def xQuantDequantDev(d_coef, d_level, d_dq, n_coef, qp, stream=0):
    _ck(lib().xQuantDequantDev(d_coef, d_level, d_dq, n_coef, qp, stream), "xQuantDequantDev")


def xIntra32DecideDev(d_cur, d_refs, d_cost, d_best, n, stream=0):
    _ck(lib().xIntra32DecideDev(d_cur, d_refs, d_cost, d_best, n, stream), "xIntra32DecideDev")


def sad(a, b):
    """reference name/signature: int sad(type* a, type* b, size_t n) over an n x n byte region"""
    a = _np(a, np.uint8); b = _np(b, np.uint8)
    n = int(round(a.size ** 0.5))
    assert n * n == a.size == b.size
    return int(lib().sad(a.ctypes.data, b.ctypes.data, n))


def xSad8x8Search(cur, ref_padded, rng, blk0=0, blk1=None, want_cost=True, want_best=True, u16=False):
    """u16=True: xSad8x8SearchU16 (16-bit cost surface)"""
    cur = _np(cur, np.uint8); ref_padded = _np(ref_padded, np.uint8)
    h, w = cur.shape
    assert ref_padded.shape == (h + 2 * rng, w + 2 * rng)
    if blk1 is None:
        blk1 = (w // 8) * (h // 8)
    side = 2 * rng + 1
    nb = blk1 - blk0
    cost = np.empty((nb, side, side), np.uint16 if u16 else np.uint32) if want_cost else None
    best = np.empty((nb, 3), np.int32) if want_best else None
    fn = lib().xSad8x8SearchU16 if u16 else lib().xSad8x8Search
    _ck(fn(cur.ctypes.data, ref_padded.ctypes.data, ref_padded.shape[1], w, h, rng, blk0, blk1,
           cost.ctypes.data if want_cost else None, best.ctypes.data if want_best else None), "xSad8x8SearchU16" if u16 else "xSad8x8Search")
    return cost, best


def xSad8x8SearchDev(d_cur, d_ref, strd, w, h, rng, blk0, blk1, d_cost, d_best, stream=0):
    _ck(lib().xSad8x8SearchDev(d_cur, d_ref, strd, w, h, rng, blk0, blk1, d_cost, d_best, stream), "xSad8x8SearchDev")


def xSad8x8SearchU16Dev(d_cur, d_ref, strd, w, h, rng, blk0, blk1, d_cost, d_best, stream=0):
    _ck(lib().xSad8x8SearchU16Dev(d_cur, d_ref, strd, w, h, rng, blk0, blk1, d_cost, d_best, stream), "xSad8x8SearchU16Dev")


def xSatd8x8SearchU16Dev(d_cur, d_ref, strd, w, h, rng, blk0, blk1, d_cost, d_best, stream=0):
    _ck(lib().xSatd8x8SearchU16Dev(d_cur, d_ref, strd, w, h, rng, blk0, blk1, d_cost, d_best, stream), "xSatd8x8SearchU16Dev")


def xSatd8x8SearchTiledU16Dev(d_cur_tiles, d_ref_tiles, width, height, rng, blk0, blk1, d_cost, d_best, stream=0):
    _ck(lib().xSatd8x8SearchTiledU16Dev(d_cur_tiles, d_ref_tiles, width, height, rng, blk0, blk1, d_cost, d_best, stream), "xSatd8x8SearchTiledU16Dev")


def xSad8x8SearchTiledU16Dev(d_cur_tiles, d_ref_tiles, width, height, rng, blk0, blk1, d_cost, d_best, stream=0):
    _ck(lib().xSad8x8SearchTiledU16Dev(d_cur_tiles, d_ref_tiles, width, height, rng, blk0, blk1, d_cost, d_best, stream), "xSad8x8SearchTiledU16Dev")


def xIntra32Pred(refs, modes):
    refs = _np(refs, np.uint8).reshape(-1, 129)
    modes = _np(modes, np.uint8).ravel()
    assert refs.shape[0] == modes.size
    pred = np.empty((modes.size, 32, 32), np.uint8)
    _ck(lib().xIntra32Pred(refs.ctypes.data, modes.ctypes.data, pred.ctypes.data, modes.size), "xIntra32Pred")
    return pred


def xFrameResiDct32(cur_tiles, pred_tiles, width, height, shift1st=4, shift2nd=11):
    cur_tiles = _np(cur_tiles, np.uint8)
    pred_tiles = _np(pred_tiles, np.uint8)
    assert cur_tiles.size == pred_tiles.size == (width // 16) * (height // 16) * 512
    coef = np.empty(((width // 32) * (height // 32), 32, 32), np.int16)
    _ck(lib().xFrameResiDct32(cur_tiles.ctypes.data, pred_tiles.ctypes.data, width, height, coef.ctypes.data, shift1st, shift2nd),
        "xFrameResiDct32")
    return coef


def xConvInputFmt(tiles, Y, U, V, strd_y, width, height):
    """reference name and signature (src/x266.cpp:415-421): host planes -> host tiles, in place (m_I untouched)"""
    lib().xConvInputFmt(tiles.ctypes.data, Y.ctypes.data, U.ctypes.data, V.ctypes.data, strd_y, width, height)
    return tiles


def xConvOutput420(tiles, Y, strd_y, U, V, strd_c, width, height):
    """reference name and signature (src/x266.cpp:455-462): host tiles -> host planes, in place"""
    lib().xConvOutput420(tiles.ctypes.data, Y.ctypes.data, strd_y, U.ctypes.data, V.ctypes.data, strd_c, width, height)


def xSatd8x8SearchTiled(cur_tiles, ref_tiles, width, height, rng, blk0=0, blk1=None, want_cost=True, want_best=True):
    cur_tiles = _np(cur_tiles, np.uint8); ref_tiles = _np(ref_tiles, np.uint8)
    if blk1 is None:
        blk1 = (width // 8) * (height // 8)
    side = 2 * rng + 1
    nb = blk1 - blk0
    cost = np.empty((nb, side, side), np.uint32) if want_cost else None
    best = np.empty((nb, 3), np.int32) if want_best else None
    _ck(lib().xSatd8x8SearchTiled(cur_tiles.ctypes.data, ref_tiles.ctypes.data, width, height, rng, blk0, blk1,
                                  cost.ctypes.data if want_cost else None, best.ctypes.data if want_best else None), "xSatd8x8SearchTiled")
    return cost, best


def xSatd8x8SearchTiledDev(d_cur_tiles, d_ref_tiles, width, height, rng, blk0, blk1, d_cost, d_best, stream=0):
    _ck(lib().xSatd8x8SearchTiledDev(d_cur_tiles, d_ref_tiles, width, height, rng, blk0, blk1, d_cost, d_best, stream), "xSatd8x8SearchTiledDev")


def xSad8x8SearchTiledDev(d_cur_tiles, d_ref_tiles, width, height, rng, blk0, blk1, d_cost, d_best, stream=0):
    _ck(lib().xSad8x8SearchTiledDev(d_cur_tiles, d_ref_tiles, width, height, rng, blk0, blk1, d_cost, d_best, stream), "xSad8x8SearchTiledDev")


def xConvInputFmtDev(d_tiles, d_y, d_u, d_v, strd_y, width, height, stream=0):
    _ck(lib().xConvInputFmtDev(d_tiles, d_y, d_u, d_v, strd_y, width, height, stream), "xConvInputFmtDev")


def xConvOutput420Dev(d_tiles, d_y, strd_y, d_u, d_v, strd_c, width, height, stream=0):
    _ck(lib().xConvOutput420Dev(d_tiles, d_y, strd_y, d_u, d_v, strd_c, width, height, stream), "xConvOutput420Dev")


def xFrameResiDct32Dev(d_cur, d_pred, width, height, d_coef, shift1st, shift2nd, stream=0):
    _ck(lib().xFrameResiDct32Dev(d_cur, d_pred, width, height, d_coef, shift1st, shift2nd, stream), "xFrameResiDct32Dev")


# ---- Tier 3, device pointers (ints) --------------------------------------------------------------
def xDct32BatchDev(d_src, d_dst, n_blocks, shift1st, shift2nd, stream=0):
    _ck(lib().xDct32BatchDev(d_src, d_dst, n_blocks, shift1st, shift2nd, stream), "xDct32BatchDev")


def xDctNBatchDev(log2n, d_src, d_dst, n_blocks, shift1st, shift2nd, stream=0):
    _ck(lib().xDctNBatchDev(log2n, d_src, d_dst, n_blocks, shift1st, shift2nd, stream), "xDctNBatchDev")


def xPartialButterfly32Dev(d_src, d_dst, shift, line, stream=0):
    _ck(lib().xPartialButterfly32Dev(d_src, d_dst, shift, line, stream), "xPartialButterfly32Dev")


def xSatd8x8BatchDev(d_diff, d_out, n, stream=0):
    _ck(lib().xSatd8x8BatchDev(d_diff, d_out, n, stream), "xSatd8x8BatchDev")


def xSatd8x8SearchDev(d_cur, d_ref, strd, w, h, rng, blk0, blk1, d_cost, d_best, stream=0):
    _ck(lib().xSatd8x8SearchDev(d_cur, d_ref, strd, w, h, rng, blk0, blk1, d_cost, d_best, stream), "xSatd8x8SearchDev")


def xIntra32PredDev(d_refs, d_modes, d_pred, n, stream=0):
    _ck(lib().xIntra32PredDev(d_refs, d_modes, d_pred, n, stream), "xIntra32PredDev")


# ---- Tier 1: drive the BDPI stream the way the Bluesim testbench does -------------------------------
def bdpi_dct_block():
    """dct32_genNew(); 16 x dct32_getDiff; 256 x dct32_getDct (src/mkDct32.bsv:430-470)."""
    L = lib()
    L.dct32_genNew()
    buf = (C.c_uint * 32)()
    diff = []
    for _ in range(16):
        L.dct32_getDiff(buf)
        diff.append(np.array(buf[:], np.uint32))
    words = np.array([L.dct32_getDct() for _ in range(256)], np.uint64)
    return np.stack(diff), words


def bdpi_satd_block():
    """satd8x8_genNew(); 8 x satd8x8_getDiff; satd8x8_getSatd (src/mkSatd.bsv:222-252)."""
    L = lib()
    L.satd8x8_genNew()
    buf = (C.c_uint * 4)()
    rows = []
    for _ in range(8):
        L.satd8x8_getDiff(buf)
        rows.append(np.array(buf[:], np.uint32))
    return np.stack(rows), int(L.satd8x8_getSatd())
