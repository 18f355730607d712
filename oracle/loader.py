"""ctypes bindings for oracle/liboracle.so (our CPU restatement) and oracle/_ref/libx266ref.so (the
unmodified reference C of src_tb/dct32.c + src_tb/satd.c, compiled in place).  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_TREE = os.environ.get("X266_REF", "/root/reference")

_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the checker libraries (gcc).  The reference build only happens when the reference
    tree is present; on the GPU box the prebuilt oracle/_ref/*.so that travelled with the snapshot
    is used as is."""
    need = force or not os.path.exists(os.path.join(HERE, "liboracle.so"))
    src_m = os.path.getmtime(os.path.join(HERE, "x266_oracle.c"))
    if not need and os.path.getmtime(os.path.join(HERE, "liboracle.so")) < src_m:
        need = True
    if need:
        subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so", f"X266_REF={REF_TREE}"])
    ref_so = os.path.join(HERE, "_ref", "libx266ref.so")
    conv_so = os.path.join(HERE, "_ref", "libx266conv.so")
    if os.path.exists(os.path.join(REF_TREE, "src_tb", "dct32.c")) and (force or not os.path.exists(ref_so) or not os.path.exists(conv_so)):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", f"X266_REF={REF_TREE}"])


def have_ref():
    return os.path.exists(os.path.join(HERE, "_ref", "libx266ref.so"))


class Oracle:
    """Our restatement (oracle/x266_oracle.c)."""

    def __init__(self):
        build()
        # X266_ORACLE_LIB: another build of the same source (scripts/sanitize_cpu.sh loads the ASan + UBSan one)
        L = self.lib = C.CDLL(os.environ.get("X266_ORACLE_LIB", os.path.join(HERE, "liboracle.so")))
        L.orc_build_g32.argtypes = [_i16p]
        L.orc_partialButterfly.argtypes = [_i16p, _i16p, C.c_int, C.c_int, C.c_int]
        L.orc_partialDense.argtypes = [_i16p, _i16p, C.c_int, C.c_int, C.c_int]
        L.orc_dct2d.argtypes = [_i16p, _i16p, C.c_int, C.c_int, C.c_int]
        L.orc_idct2d.argtypes = [_i16p, _i16p, C.c_int, C.c_int, C.c_int]
        L.orc_satd8x8.argtypes = [_i16p]
        L.orc_satd8x8.restype = C.c_int
        L.orc_satd_search_block.argtypes = [_u8p, C.c_int, _u8p, C.c_ssize_t, C.c_int, C.c_int, C.c_int, _u32p]
        L.orc_satd_search_frame.argtypes = [_u8p, C.c_int, C.c_int, _u8p, C.c_ssize_t, C.c_int, C.c_size_t,
                                            C.c_size_t, C.c_void_p, C.c_void_p]
        L.orc_sad_search_frame.argtypes = [_u8p, C.c_int, C.c_int, _u8p, C.c_ssize_t, C.c_int, C.c_size_t,
                                           C.c_size_t, C.c_void_p, C.c_void_p]
        L.orc_sad.argtypes = [_u8p, C.c_ssize_t, _u8p, C.c_ssize_t, C.c_int, C.c_int]
        L.orc_sad.restype = C.c_uint32
        L.orc_intra32.argtypes = [_u8p, _u8p, C.c_int, _u8p]
        L.orc_intra32_direct.argtypes = [_u8p, _u8p, C.c_int, _u8p]
        L.orc_intra_ref_line.argtypes = [_u8p, _u8p, C.c_int, _i32p]
        L.orc_intra_idx_frac.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_quant.argtypes = [_i16p, _i16p, C.c_size_t, C.c_int]
        L.orc_dequant.argtypes = [_i16p, _i16p, C.c_size_t, C.c_int]
        L.orc_intra32_recon.argtypes = [_u8p, _u8p, _u8p, C.c_int, C.c_int, _i16p, _u8p]
        L.orc_intra32_encode.argtypes = [_u8p, _u8p, _u8p, C.c_int, _i16p, _u8p, _u32p, C.c_void_p]
        L.orc_intra32_decide.argtypes = [_u8p, _u8p, _u8p, _u32p, C.c_void_p]
        L.orc_intra_mode_angle.argtypes = [C.c_int]
        L.orc_intra_mode_angle.restype = C.c_int
        L.orc_conv_input_fmt.argtypes = [_u8p, _u8p, _u8p, _u8p, C.c_ssize_t, C.c_int, C.c_int]
        L.orc_conv_output420.argtypes = [_u8p, _u8p, C.c_ssize_t, _u8p, _u8p, C.c_ssize_t, C.c_int, C.c_int]
        L.orc_frame_resi_dct32.argtypes = [_u8p, _u8p, C.c_int, C.c_int, _i16p, C.c_int, C.c_int]
        L.orc_fnv1a64.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_fnv1a64.restype = C.c_uint64
        L.orc_fill_residual.argtypes = [_i16p, C.c_size_t, C.c_uint64, C.c_int]
        L.orc_dct_batch.argtypes = [_i16p, _i16p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_dct_batch.restype = C.c_int
        L.orc_satd8x8_batch.argtypes = [_i16p, _i32p, C.c_size_t, C.c_int]
        L.orc_satd8x8_batch.restype = C.c_int

    # --- transform -------------------------------------------------------------------------
    def g32(self):
        g = np.zeros((32, 32), np.int16)
        self.lib.orc_build_g32(g)
        return g

    def partial(self, src, shift, line, log2n=5, dense=False):
        src = np.ascontiguousarray(src, np.int16)
        dst = np.zeros(src.size, np.int16)
        (self.lib.orc_partialDense if dense else self.lib.orc_partialButterfly)(src.ravel(), dst, shift, line, log2n)
        return dst

    def dct(self, src, log2n=5, shift1=4, shift2=11, threads=1, out=None):
        """src: [nBlocks, N, N] int16 -> same shape (out: preallocated result, so a timed call pays no first-touch faults)."""
        src = np.ascontiguousarray(src, np.int16)
        dst = np.empty_like(src) if out is None else out
        n = src.size >> (2 * log2n)
        rc = self.lib.orc_dct_batch(src.ravel(), dst.ravel(), n, log2n, shift1, shift2, threads)
        assert rc == 0
        return dst

    def idct(self, src, log2n=5, shift1=7, shift2=12):
        src = np.ascontiguousarray(src, np.int16)
        bs = 1 << (2 * log2n)
        flat = src.reshape(-1, bs)
        dst = np.empty_like(flat)
        for i in range(flat.shape[0]):
            self.lib.orc_idct2d(flat[i], dst[i], log2n, shift1, shift2)
        return dst.reshape(src.shape)

    # --- SATD ------------------------------------------------------------------------------
    def satd(self, diff, threads=1):
        diff = np.ascontiguousarray(diff, np.int16)
        n = diff.size // 64
        out = np.empty(n, np.int32)
        assert self.lib.orc_satd8x8_batch(diff.ravel(), out, n, threads) == 0
        return out

    def satd_search(self, cur, ref_pad, rng, blk0, blk1, want_cost=True, want_best=True):
        cur = np.ascontiguousarray(cur, np.uint8)
        ref_pad = np.ascontiguousarray(ref_pad, np.uint8)
        h, w = cur.shape
        side = 2 * rng + 1
        nb = blk1 - blk0
        cost = np.empty((nb, side, side), np.uint32) if want_cost else None
        best = np.empty((nb, 3), np.int32) if want_best else None
        self.lib.orc_satd_search_frame(cur, w, h, ref_pad, ref_pad.shape[1], rng, blk0, blk1,
                                       cost.ctypes.data if want_cost else None,
                                       best.ctypes.data if want_best else None)
        return cost, best

    def sad_search(self, cur, ref_pad, rng, blk0, blk1):
        cur = np.ascontiguousarray(cur, np.uint8); ref_pad = np.ascontiguousarray(ref_pad, np.uint8)
        h, w = cur.shape
        side = 2 * rng + 1
        cost = np.empty((blk1 - blk0, side, side), np.uint32); best = np.empty((blk1 - blk0, 3), np.int32)
        self.lib.orc_sad_search_frame(cur, w, h, ref_pad, ref_pad.shape[1], rng, blk0, blk1, cost.ctypes.data, best.ctypes.data)
        return cost, best

    def sad(self, a, b):
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        return int(self.lib.orc_sad(a, a.shape[1], b, b.shape[1], a.shape[1], a.shape[0]))

    # --- intra -----------------------------------------------------------------------------
    def intra32(self, left, top, mode):
        pred = np.empty((32, 32), np.uint8)
        self.lib.orc_intra32(np.ascontiguousarray(left, np.uint8), np.ascontiguousarray(top, np.uint8), mode, pred)
        return pred

    def intra32_direct(self, left, top, mode):
        """the second, table-free restatement (per pixel, sample-position form)"""
        pred = np.empty((32, 32), np.uint8)
        self.lib.orc_intra32_direct(np.ascontiguousarray(left, np.uint8), np.ascontiguousarray(top, np.uint8), mode, pred)
        return pred

    def intra_ref_line(self, left, top, mode):
        """ref[-32..64] of an angular mode as the predictor builds it; -1 = never read"""
        out = np.empty(97, np.int32)
        self.lib.orc_intra_ref_line(np.ascontiguousarray(left, np.uint8), np.ascontiguousarray(top, np.uint8), mode, out)
        return out

    def intra_idx_frac(self, mode, k):
        idx, f = C.c_int(0), C.c_int(0)
        self.lib.orc_intra_idx_frac(mode, k, C.byref(idx), C.byref(f))
        return idx.value, f.value

    def intra32_decide(self, cur, left, top):
        cost = np.zeros(35, np.uint32)
        best = C.c_int32(0)
        self.lib.orc_intra32_decide(np.ascontiguousarray(cur, np.uint8).ravel(), np.ascontiguousarray(left, np.uint8),
                                    np.ascontiguousarray(top, np.uint8), cost, C.byref(best))
        return cost, best.value

    def quant(self, coef, qp):
        coef = np.ascontiguousarray(coef, np.int16)
        out = np.empty_like(coef)
        self.lib.orc_quant(coef.ravel(), out.ravel(), coef.size, qp)
        return out

    def dequant(self, level, qp):
        level = np.ascontiguousarray(level, np.int16)
        out = np.empty_like(level)
        self.lib.orc_dequant(level.ravel(), out.ravel(), level.size, qp)
        return out

    def intra32_recon(self, cur, left, top, mode, qp):
        level = np.empty(1024, np.int16); recon = np.empty(1024, np.uint8)
        self.lib.orc_intra32_recon(np.ascontiguousarray(cur, np.uint8).ravel(), np.ascontiguousarray(left, np.uint8),
                                   np.ascontiguousarray(top, np.uint8), mode, qp, level, recon)
        return level.reshape(32, 32), recon.reshape(32, 32)

    def intra32_encode(self, cur, left, top, qp):
        level = np.empty(1024, np.int16); recon = np.empty(1024, np.uint8)
        cost = np.zeros(35, np.uint32); best = C.c_int32(0)
        self.lib.orc_intra32_encode(np.ascontiguousarray(cur, np.uint8).ravel(), np.ascontiguousarray(left, np.uint8),
                                    np.ascontiguousarray(top, np.uint8), qp, level, recon, cost, C.byref(best))
        return level.reshape(32, 32), recon.reshape(32, 32), cost, best.value

    # --- tiled frames ----------------------------------------------------------------------
    def conv_input_fmt(self, Y, U, V):
        h, w = Y.shape
        tiles = np.zeros((w // 16) * (h // 16) * 512, np.uint8)
        self.lib.orc_conv_input_fmt(tiles, np.ascontiguousarray(Y), np.ascontiguousarray(U), np.ascontiguousarray(V), w, w, h)
        return tiles

    def conv_output420(self, tiles, w, h):
        Y = np.zeros((h, w), np.uint8); U = np.zeros((h // 2, w // 2), np.uint8); V = np.zeros((h // 2, w // 2), np.uint8)
        self.lib.orc_conv_output420(np.ascontiguousarray(tiles, np.uint8), Y, w, U, V, w // 2, w, h)
        return Y, U, V

    def frame_resi_dct32(self, cur_tiles, pred_tiles, w, h, s1, s2):
        coef = np.zeros((w // 32) * (h // 32) * 1024, np.int16)
        self.lib.orc_frame_resi_dct32(np.ascontiguousarray(cur_tiles, np.uint8), np.ascontiguousarray(pred_tiles, np.uint8), w, h, coef, s1, s2)
        return coef.reshape(-1, 32, 32)

    # --- KAT helpers -----------------------------------------------------------------------
    def fnv(self, arr):
        arr = np.ascontiguousarray(arr)
        return int(self.lib.orc_fnv1a64(arr.ctypes.data, arr.nbytes))

    def residual(self, n, seed, kind):
        out = np.empty(n, np.int16)
        self.lib.orc_fill_residual(out, n, seed, kind)
        return out


def have_ref_conv():
    return os.path.exists(os.path.join(HERE, "_ref", "libx266conv.so"))


class RefConv:
    """The unmodified xConvInputFmt / xConvOutput420 of src/x266.cpp:415-492 (oracle/_ref/libx266conv.so, ref_conv_slice.sh)."""

    def __init__(self):
        build()
        L = self.lib = C.CDLL(os.path.join(HERE, "_ref", "libx266conv.so"))
        L.ref_xConvInputFmt.argtypes = [_u8p, _u8p, _u8p, _u8p, C.c_ssize_t, C.c_int, C.c_int]
        L.ref_xConvOutput420.argtypes = [_u8p, _u8p, C.c_ssize_t, _u8p, _u8p, C.c_ssize_t, C.c_int, C.c_int]
        assert L.ref_sizeof_ref_block_t() == 512

    def input_fmt(self, y_buf, u_buf, v_buf, strd_y, w, h, tiles):
        """planes as flat byte buffers with luma stride strd_y (chroma stride strd_y >> 1, x266.cpp:426); tiles is written in place"""
        self.lib.ref_xConvInputFmt(tiles, y_buf, u_buf, v_buf, strd_y, w, h)
        return tiles

    def output420(self, tiles, y_buf, strd_y, u_buf, v_buf, strd_c, w, h):
        self.lib.ref_xConvOutput420(tiles, y_buf, strd_y, u_buf, v_buf, strd_c, w, h)


class Ref:
    """The unmodified reference (oracle/_ref/libx266ref.so).  variant 'o2' (headline) or 'o3'."""

    def __init__(self, variant="o2"):
        build()
        name = "libx266ref.so" if variant == "o2" else "libx266ref_o3.so"
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.lib = C.CDLL(path)
        L.ref_partialButterfly32.argtypes = [_i16p, _i16p, C.c_int, C.c_int]
        L.ref_satd8x8.argtypes = [_i16p]
        L.ref_satd8x8.restype = C.c_int
        L.ref_dct32_2d.argtypes = [_i16p, _i16p, C.c_int, C.c_int]
        L.ref_dct32_batch.argtypes = [_i16p, _i16p, C.c_size_t, C.c_int, C.c_int, C.c_int]
        L.ref_dct32_batch.restype = C.c_int
        L.ref_satd8x8_batch.argtypes = [_i16p, _i32p, C.c_size_t, C.c_int]
        L.ref_satd8x8_batch.restype = C.c_int
        for f in ("ref_dct32_lastMat", "ref_dct32_lastDct", "ref_satd_lastMat"):
            getattr(L, f).restype = C.POINTER(C.c_int16)
        L.dct32_getDct.restype = C.c_uint64
        L.satd8x8_getSatd.restype = C.c_uint32
        self.libc = C.CDLL(None)

    def g32(self):
        return np.ctypeslib.as_array((C.c_int16 * 1024).in_dll(self.lib, "g_t32")).reshape(32, 32).copy()

    def partial32(self, src, shift, line):
        src = np.ascontiguousarray(src, np.int16)
        dst = np.zeros(src.size, np.int16)
        self.lib.ref_partialButterfly32(src.ravel(), dst, shift, line)
        return dst

    def dct32(self, src, shift1=4, shift2=11, threads=1, out=None):
        src = np.ascontiguousarray(src, np.int16)
        dst = np.empty_like(src) if out is None else out
        assert self.lib.ref_dct32_batch(src.ravel(), dst.ravel(), src.size // 1024, shift1, shift2, threads) == 0
        return dst

    def satd(self, diff, threads=1, out=None):
        diff = np.ascontiguousarray(diff, np.int16)
        out = np.empty(diff.size // 64, np.int32) if out is None else out
        assert self.lib.ref_satd8x8_batch(diff.ravel(), out, out.size, threads) == 0
        return out

    # Tier-1 BDPI stream as the Bluesim testbench drives it (mkDct32.bsv:430-470, mkSatd.bsv:222-252)
    def srand(self, seed):
        self.libc.srand(seed)

    def bdpi_dct_block(self):
        self.lib.dct32_genNew()
        buf = (C.c_uint * 32)()
        diff = []
        for _ in range(16):
            self.lib.dct32_getDiff(buf)
            diff.append(np.array(buf[:], np.uint32))
        words = np.array([self.lib.dct32_getDct() for _ in range(256)], np.uint64)
        mat = np.ctypeslib.as_array(self.lib.ref_dct32_lastMat(), (1024,)).copy()
        dct = np.ctypeslib.as_array(self.lib.ref_dct32_lastDct(), (1024,)).copy()
        return np.stack(diff), words, mat, dct

    def bdpi_satd_block(self):
        self.lib.satd8x8_genNew()
        buf = (C.c_uint * 4)()
        rows = []
        for _ in range(8):
            self.lib.satd8x8_getDiff(buf)
            rows.append(np.array(buf[:], np.uint32))
        return np.stack(rows), int(self.lib.satd8x8_getSatd())
