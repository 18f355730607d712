"""GPU suite: bit-exact parity of the CUDA path, called through the C ABI, against the oracle, the
committed reference vectors, the KAT hashes, and (when oracle/_ref travelled) the reference itself."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHIFTS = {2: (1, 8), 3: (2, 9), 4: (3, 10), 5: (4, 11)}
VARIANTS = ["bfly", "imma"]


@pytest.fixture(params=VARIANTS)
def variant(request, x266):
    x266.set_dct_variant(x266.DCT_BFLY if request.param == "bfly" else x266.DCT_IMMA)
    yield request.param
    x266.set_dct_variant(x266.DCT_AUTO)


# ---------------------------------------------------------------------------------------- DCT32
def test_dct32_reference_vectors(x266, vectors, variant):
    x = vectors["dct_in"]
    for key, sh in (("dct_out_4_11", (4, 11)), ("dct_out_6_11", (6, 11)), ("dct_out_1_1", (1, 1)), ("dct_out_9_16", (9, 16))):
        got = x266.xDct32Batch(x, *sh)
        assert np.array_equal(got, vectors[key]), (variant, key)


@pytest.mark.parametrize("name", ["KAT-A", "KAT-B", "KAT-C"])
def test_dct32_kat(x266, orc, kat, variant, name):
    """config 2: one 1080p frame (2040 blocks), memcmp + FNV of the output."""
    k = kat[name]
    x = orc.residual(k["blocks"] * 1024, k["seed"], k["kind"])
    y = x266.xDct32Batch(x, *k["shifts"])
    assert y[:4].tolist() == k["out_first4"]
    assert f"{orc.fnv(y):016x}" == k["fnv_out"]
    assert np.array_equal(y, orc.dct(x.reshape(-1, 32, 32), 5, *k["shifts"], threads=8).ravel())


def test_dct32_8k_frame_kat(x266, orc, kat):
    """config 5 flavour: one 7680x4320 frame, 11-bit residuals, shifts 6/11 -- hash minted from the reference."""
    k = kat["KAT-8K"]
    x = orc.residual(k["blocks"] * 1024, k["seed"], k["kind"])
    for v in (x266.DCT_IMMA, x266.DCT_BFLY):
        x266.set_dct_variant(v)
        y = x266.xDct32Batch(x, *k["shifts"])
        assert f"{orc.fnv(y):016x}" == k["fnv_out"]
    x266.set_dct_variant(x266.DCT_AUTO)


@pytest.mark.parametrize("n", [0, 1, 2, 7, 8, 9, 63, 295, 296, 297, 2367, 2369, 8192 + 5, 3 * 8192 + 1])
def test_dct32_ragged_batch_sizes(x266, orc, variant, n):
    """empty / ragged batches around the warps-per-CTA, resident-grid and pipeline-chunk boundaries."""
    x = orc.residual(n * 1024, 31 + n, 2)
    y = x266.xDct32Batch(x, 4, 11)
    assert y.shape == x.shape
    if n:
        assert np.array_equal(y, orc.dct(x.reshape(-1, 32, 32), 5, 4, 11, threads=8).ravel())


def test_dct32_extremes_and_wrap(x266, orc, variant):
    blocks = [np.full(1024, v, np.int16) for v in (0, 1, -1, 255, -255, 1023, -1023, 32767, -32768)]
    blocks.append(np.where(np.arange(1024) % 2 == 0, 32767, -32768).astype(np.int16))
    blocks.append(np.where(np.arange(1024) % 2 == 0, -32768, 32767).astype(np.int16))
    g = orc.g32()
    for k in (1, 15, 31):            # worst-case sign patterns: maximise |sum| of row k in both passes
        blocks.append(np.where(np.tile(g[k], 32) >= 0, 32767, -32768).astype(np.int16))
        blocks.append(np.where(np.repeat(g[k], 32) >= 0, -32768, 32767).astype(np.int16))
    x = np.stack(blocks).reshape(-1, 32, 32)
    for sh in ((4, 11), (6, 11), (1, 1), (16, 16), (1, 16)):
        assert np.array_equal(x266.xDct32Batch(x, *sh), orc.dct(x, 5, *sh)), sh
    assert int(x266.xDct32Batch(np.full(1024, 1023, np.int16), 4, 11)[0]) == -128


def test_dct32_against_live_reference(x266, ref, variant):
    rng = np.random.default_rng(11)
    x = rng.integers(-32768, 32768, (4096, 32, 32)).astype(np.int16)
    assert np.array_equal(x266.xDct32Batch(x, 4, 11), ref.dct32(x, 4, 11, threads=8))


def test_dct32_linearity_property_full_size(x266, orc):
    """size-independent property at the bench size (config 5: 64 frames of 8K = 2.07M blocks is too big
    for a CPU check): with shifts that cannot round (inputs multiples of 2^15 are not needed -- we use
    the impulse identity instead): T(e_k) reads back g_t32, and T is exactly additive on inputs whose
    pass-1 sums are multiples of 2^shift1 and pass-2 sums multiples of 2^shift2."""
    g = orc.g32().astype(np.int64)
    # impulse at (r, c) scaled by 2^10: pass 1 gives 2^6 * g[k][c] at coef[k][r]; pass 2 gives
    # (2^6 * g[k][c] * g[k2][r] + 1024) >> 11 -- closed form, checked for a full frame of impulses
    n = 32400
    rr = np.arange(n) % 32
    cc = (np.arange(n) // 32) % 32
    x = np.zeros((n, 32, 32), np.int16)
    x[np.arange(n), rr, cc] = 1024
    y = x266.xDct32Batch(x, 4, 11).reshape(n, 32, 32)
    for b in (0, 1, 33, 1023, 5000, n - 1):
        want = ((64 * np.outer(g[:, rr[b]], g[:, cc[b]]) + 1024) >> 11).astype(np.int16)
        assert np.array_equal(y[b], want)


# ------------------------------------------------------------------------------ Tier 2 / Tier 1
@pytest.mark.parametrize("line", [1, 5, 32, 33, 1000])
def test_partialButterfly32_tier2(x266, orc, line):
    x = orc.residual(line * 32, line, 2)
    for shift in (1, 4, 11, 16):
        assert np.array_equal(x266.partialButterfly32(x, shift, line), orc.partial(x, shift, line))


def test_partialButterfly32_reference_vector(x266, vectors):
    assert np.array_equal(x266.partialButterfly32(vectors["partial_src"], 4, 5), vectors["partial_out_shift4_line5"])


def test_satd8x8_tier2(x266, orc, vectors):
    for d, want in zip(vectors["satd_in"][:40], vectors["satd_out"][:40]):
        assert x266.satd8x8(d) == int(want)
    assert x266.satd8x8(np.full(64, 1023, np.int16)) == 16


def test_bdpi_stream_matches_testbench_vectors(x266, kat, vectors):
    """Tier 1: srand(1); dct32_genNew/getDiff/getDct and satd8x8_* call for call, as mkTb drives them
    (src/mkDct32.bsv:430-470, src/mkSatd.bsv:222-252), against the stream the reference produced."""
    libc = C.CDLL(None)
    libc.srand(1)
    diff, words = x266.bdpi_dct_block()
    assert np.array_equal(diff, vectors["srand1_getDiff"])
    assert np.array_equal(words, vectors["srand1_getDct"])
    assert f"{int(words[0]):016x}" == kat["srand1"]["getDct_word0"]
    libc.srand(1)
    rows, satd = x266.bdpi_satd_block()
    assert np.array_equal(rows, vectors["srand1_satd_rows"]) and satd == kat["srand1"]["satd_first"]


def test_bdpi_stream_against_live_reference(x266, ref):
    libc = C.CDLL(None)
    libc.srand(12345)
    want = [ref.bdpi_dct_block()[:2] for _ in range(11)]            # the testbench checks 11 blocks
    libc.srand(12345)
    for wd, ww in want:
        gd, gw = x266.bdpi_dct_block()
        assert np.array_equal(gd, wd) and np.array_equal(gw, ww)
    libc.srand(777)
    want = [ref.bdpi_satd_block() for _ in range(256)]              # mkSatd.bsv: 256 iterations
    libc.srand(777)
    for wr, ws in want:
        gr, gs = x266.bdpi_satd_block()
        assert np.array_equal(gr, wr) and gs == ws


# ------------------------------------------------------------------------------------ DCT N<32
@pytest.mark.parametrize("log2n", [2, 3, 4])
@pytest.mark.parametrize("nblk", [1, 2, 3, 4, 5, 64, 1000, 4097, 75777])
@pytest.mark.parametrize("cuda_core", [0, 1])
def test_dctN(x266, orc, log2n, nblk, cuda_core):
    """N<32: tensor-core kernels where they exist (tune 3 = 0) and the CUDA-core dctN kernels (tune 3 = 1)"""
    n = 1 << log2n
    x266.tune(3, cuda_core)
    try:
        for kind in (0, 2):
            x = orc.residual(nblk * n * n, 5 + nblk, kind)
            for s1, s2 in (SHIFTS[log2n], (1, 16)):
                got = x266.xDctNBatch(log2n, x, s1, s2)
                assert np.array_equal(got, orc.dct(x.reshape(-1, n, n), log2n, s1, s2, threads=4).ravel())
    finally:
        x266.tune(3, 0)


def test_config4_mixed_partition(x266, orc):
    """config 4: 3840x2176, every 32x32 region split into 1x32^2 / 4x16^2 / 16x8^2 / 64x4^2 by a
    seeded draw, blocks gathered per size class, 9-bit residuals, shifts (1,8)/(2,9)/(3,10)/(4,11)."""
    regions = (3840 // 32) * (2176 // 32)
    z = np.frombuffer(orc.residual(regions, 268, 2).tobytes(), np.uint16) & 3        # partition class per region
    for cls, log2n in enumerate((5, 4, 3, 2)):
        n = 1 << log2n
        nblk = int((z == cls).sum()) * (1024 // (n * n))
        x = orc.residual(nblk * n * n, 300 + cls, 0)
        s1, s2 = SHIFTS[log2n]
        got = x266.xDctNBatch(log2n, x, s1, s2)
        assert np.array_equal(got, orc.dct(x.reshape(-1, n, n), log2n, s1, s2, threads=8).ravel()), log2n


# ---------------------------------------------------------------------------------------- SATD
SATD_VARIANTS = {"imma-ring3": 0, "cuda-core": 1, "imma-ring4": 5}      # xGpuTune(2, id)


@pytest.fixture(params=SATD_VARIANTS)
def satd_variant(request, x266):
    x266.tune(2, SATD_VARIANTS[request.param])
    yield request.param
    x266.tune(2, 0)


@pytest.mark.parametrize("name", ["KAT-D", "KAT-E"])
def test_satd_kat(x266, orc, kat, name, satd_variant):
    k = kat[name]
    x = orc.residual(k["blocks"] * 64, k["seed"], k["kind"])
    y = x266.xSatd8x8Batch(x)
    assert y[:4].tolist() == k["out_first4"]
    assert f"{orc.fnv(y):016x}" == k["fnv_out"]


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 15, 16, 17, 31, 32, 33, 255, 257, 100003, (1 << 17) + 3])
def test_satd_ragged(x266, orc, n, satd_variant):
    x = orc.residual(n * 64, n + 1, 2)
    y = x266.xSatd8x8Batch(x)
    assert y.shape == (n,)
    if n:
        assert np.array_equal(y, orc.satd(x, threads=8))


def test_satd_reference_vectors_and_extremes(x266, vectors, satd_variant):
    assert np.array_equal(x266.xSatd8x8Batch(vectors["satd_in"]), vectors["satd_out"])
    assert int(x266.xSatd8x8Batch(np.full(64, 255, np.int16))[0]) == 4080
    assert int(x266.xSatd8x8Batch(np.full(64, -32768, np.int16))[0]) == int(vectors["satd_out"][66])


def make_frames(w, h, rng_px, seed=266):
    """config 3 synthetic pair: ref = cur shifted by (+5,-3) + small noise, edge-replicated."""
    r = np.random.default_rng(seed)
    cur = r.integers(0, 256, (h, w)).astype(np.uint8)
    ys = np.clip(np.arange(h) + 3, 0, h - 1)
    xs = np.clip(np.arange(w) - 5, 0, w - 1)
    ref = np.clip(cur[ys][:, xs].astype(int) + r.integers(-4, 5, (h, w)), 0, 255).astype(np.uint8)
    return cur, np.pad(ref, rng_px, mode="edge")


@pytest.mark.parametrize("rng_px", [0, 3, 8, 16, 32])
@pytest.mark.parametrize("v1", [0, 1])
def test_satd_search_small(x266, orc, rng_px, v1):
    """both search kernels (0: v3 packed transform domain, two positions per thread, R in {8,16,32}; 1: one CTA per block, any range --
    other ranges take it anyway)"""
    x266.tune(1, v1)
    cur, refp = make_frames(64, 48, rng_px)
    cost, best = x266.xSatd8x8Search(cur, refp, rng_px)
    x266.tune(1, 0)
    wc, wb = orc.satd_search(cur, refp, rng_px, 0, 48)
    assert np.array_equal(cost, wc)
    assert np.array_equal(best, wb)


@pytest.mark.parametrize("rng_px", [8, 16, 32])
def test_satd_search_ragged_strips_and_subranges(x266, orc, rng_px):
    """width 200 = one full strip of 16 blocks + a ragged strip of 9; block sub-ranges that start and end
    mid-strip / mid-row; extreme flat frames (all ties -> the tie-break rule decides)."""
    cur, refp = make_frames(200, 24, rng_px, seed=5)
    nblk = 25 * 3
    wc, wb = orc.satd_search(cur, refp, rng_px, 0, nblk)
    cost, best = x266.xSatd8x8Search(cur, refp, rng_px)
    assert np.array_equal(cost, wc) and np.array_equal(best, wb)
    for b0, b1 in ((3, 4), (14, 19), (20, 60), (74, 75)):
        c, b = x266.xSatd8x8Search(cur, refp, rng_px, b0, b1)
        assert np.array_equal(c, wc[b0:b1]) and np.array_equal(b, wb[b0:b1])
        _, b = x266.xSatd8x8Search(cur, refp, rng_px, b0, b1, want_cost=False)
        assert np.array_equal(b, wb[b0:b1])
    flat = np.full((24, 200), 200, np.uint8)
    flatp = np.full((24 + 2 * rng_px, 200 + 2 * rng_px), 10, np.uint8)
    c, b = x266.xSatd8x8Search(flat, flatp, rng_px)
    assert (b[:, 1] == 0).all() and (b[:, 2] == 0).all() and (c == c[0, 0, 0]).all()


@pytest.mark.parametrize("rng_px", [8, 16, 32])
def test_satd_search_extreme_patterns(x266, orc, rng_px):
    """v3 keeps two biased coefficients per 32-bit word: frames built from the 64 Hadamard basis patterns (pixels
    0/255, every coefficient driven to +-8160 / 16320) and from random 0/255 pixels must not carry between halves."""
    r = np.random.default_rng(11)
    H = np.array([[(-1) ** bin(a & b).count("1") for b in range(64)] for a in range(64)])
    w, h = 256, 32
    cur = np.zeros((h, w), np.uint8)
    ref = np.zeros((h, w), np.uint8)
    for by in range(h // 8):
        for bx in range(w // 8):
            k = int(r.integers(0, 64))
            cur[by * 8:by * 8 + 8, bx * 8:bx * 8 + 8] = np.where(H[k] > 0, 255, 0).reshape(8, 8)
            ref[by * 8:by * 8 + 8, bx * 8:bx * 8 + 8] = np.where(H[k] > 0, 0, 255).reshape(8, 8)
    for c, f in ((cur, np.pad(ref, rng_px, mode="edge")),
                 (r.choice([0, 255], (h, w)).astype(np.uint8), r.choice([0, 255], (h + 2 * rng_px, w + 2 * rng_px)).astype(np.uint8))):
        cost, best = x266.xSatd8x8Search(c, f, rng_px)
        wc, wb = orc.satd_search(c, f, rng_px, 0, (w // 8) * (h // 8))
        assert np.array_equal(cost, wc) and np.array_equal(best, wb)


def test_satd_search_config3_golden(x266):
    """Config 3 at full size against the committed known-answer file minted from the unmodified reference (no CPU checker
    involved at test time): all 32400 argmins, and the full cost surfaces of the sampled blocks by FNV-1a-64."""
    from search_frames import config3_frames, fnv1a64
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "search_kat.npz"))
    cur, refp = config3_frames()
    _, best = x266.xSatd8x8Search(cur, refp, 32, want_cost=False)
    assert np.array_equal(best, g["best"])
    interior = [by * 240 + bx for by in range(8, 127, 17) for bx in range(8, 232, 23)]
    assert all(best[b, 1] == 5 and best[b, 2] == -3 for b in interior)
    for k in range(0, len(g["sample"]), 4):
        b = int(g["sample"][k])
        cost, bb = x266.xSatd8x8Search(cur, refp, 32, b, b + 1)
        assert fnv1a64(cost[0]) == int(g["sample_fnv"][k]) and np.array_equal(bb[0], g["best"][b])


def test_satd_search_1080p_sample(x266, orc):
    """config 3 at full size: whole 1920x1080 frame, +-32; argmins for all 32400 blocks come back in one
    call; the oracle checks a deterministic sample of 256 full cost surfaces + their argmins, and the
    interior blocks must find the planted motion (+5,-3)."""
    cur, refp = make_frames(1920, 1080, 32)
    nblk = 240 * 135
    _, best = x266.xSatd8x8Search(cur, refp, 32, want_cost=False)
    interior = [by * 240 + bx for by in range(8, 127, 17) for bx in range(8, 232, 23)]
    assert all(best[b, 1] == 5 and best[b, 2] == -3 for b in interior)
    sample = sorted(set(int(v) for v in np.random.default_rng(1).integers(0, nblk, 256)))
    for b in sample[:256]:
        cost, bb = x266.xSatd8x8Search(cur, refp, 32, b, b + 1)
        wc, wb = orc.satd_search(cur, refp, 32, b, b + 1)
        assert np.array_equal(cost, wc) and np.array_equal(bb, wb) and np.array_equal(best[b], wb[0])


# --------------------------------------------------------------------------------------- intra
@pytest.mark.parametrize("swar", [0, 1])
def test_intra32_all_modes(x266, orc, swar):
    """swar = 0: fractional angles on the tensor cores (W x Hankel product, 8 IMMA per prediction), copies / DC / planar on CUDA
    cores; swar = 1: every angular mode through the CUDA-core SWAR interpolation"""
    r = np.random.default_rng(5)
    refs = r.integers(0, 256, (35 * 6, 129)).astype(np.uint8)
    modes = np.tile(np.arange(35, dtype=np.uint8), 6)
    refs[0:35] = 0
    refs[35:70] = 255
    refs[70:105] = r.choice([0, 255], (35, 129)).astype(np.uint8)      # extremes: every weighted sum at its limits
    x266.tune(8, swar)
    try:
        pred = x266.xIntra32Pred(refs, modes)
    finally:
        x266.tune(8, 0)
    for i in range(modes.size):
        assert np.array_equal(pred[i], orc.intra32(refs[i, :64], refs[i, 64:], int(modes[i]))), int(modes[i])


# ------------------------------------------------------------------------- device-pointer entry
def test_intra32_pred_modes(x266, orc):
    """mode-major entry: all 35 modes of every block, a sparse mask, a single mode, odd block counts, misaligned device refs"""
    import torch
    rng = np.random.default_rng(77)
    for nb, mask in ((1, (1 << 35) - 1), (9, (1 << 35) - 1), (8, 0b1000000000100000000010000000111), (3, 1 << 34), (5, 1 << 0)):
        refs = rng.integers(0, 256, (nb, 129), dtype=np.uint8)
        got = x266.xIntra32PredModes(refs, mask)
        ms = [m for m in range(35) if mask >> m & 1]
        assert got.shape == (nb, len(ms), 32, 32)
        for b in range(nb):
            for j, m in enumerate(ms):
                assert np.array_equal(got[b, j], orc.intra32(refs[b, :64], refs[b, 64:], m)), (nb, b, m)
    refs = rng.integers(0, 256, (4, 129), dtype=np.uint8)
    buf = torch.zeros(4 * 129 + 8, dtype=torch.uint8, device="cuda")
    out = torch.empty((4, 35, 1024), dtype=torch.uint8, device="cuda")
    for off in (1, 2, 3):
        buf[off:off + 4 * 129] = torch.from_numpy(refs.ravel()).cuda()
        x266.xIntra32PredModesDev(buf.data_ptr() + off, 4, (1 << 35) - 1, out.data_ptr())
        torch.cuda.synchronize()
        got = out.cpu().numpy().reshape(4, 35, 32, 32)
        for b in range(4):
            for m in range(35):
                assert np.array_equal(got[b, m], orc.intra32(refs[b, :64], refs[b, 64:], m)), (off, b, m)
    with pytest.raises(x266.X266Error):
        x266.xIntra32PredModes(refs, 1 << 35)


def test_device_pointer_entry_points(x266, orc):
    import torch
    dev = torch.device("cuda:0")
    x = orc.residual(1000 * 1024, 9, 1)
    xs = torch.from_numpy(x).to(dev)
    ys = torch.empty_like(xs)
    st = torch.cuda.current_stream().cuda_stream
    for v in (x266.DCT_BFLY, x266.DCT_IMMA):
        x266.set_dct_variant(v)
        ys.zero_()
        x266.xDct32BatchDev(xs.data_ptr(), ys.data_ptr(), 1000, 6, 11, st)
        torch.cuda.synchronize()
        assert np.array_equal(ys.cpu().numpy(), orc.dct(x.reshape(-1, 32, 32), 5, 6, 11, threads=8).ravel())
    x266.set_dct_variant(x266.DCT_AUTO)
    d = torch.from_numpy(orc.residual(5000 * 64, 3, 0)).to(dev)
    o = torch.empty(5000, dtype=torch.int32, device=dev)
    x266.xSatd8x8BatchDev(d.data_ptr(), o.data_ptr(), 5000, st)
    torch.cuda.synchronize()
    assert np.array_equal(o.cpu().numpy(), orc.satd(d.cpu().numpy()))
    launches = x266.kernel_launches()
    assert launches > 0


@pytest.mark.parametrize("rng_px", [8, 32])
def test_search_and_intra_misaligned_device_pointers(x266, orc, rng_px):
    """The search prologues use 32/64-bit loads when the planes allow it and byte loads otherwise; the intra kernel
    reads its 129-byte reference records as aligned words when the array is 4-byte aligned.  Odd base addresses and an
    odd reference stride must take the fallbacks and give the same answers."""
    import torch
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    w, h = 136, 24
    cur, refp = make_frames(w, h, rng_px, seed=21)
    nb = (w // 8) * (h // 8)
    side = 2 * rng_px + 1
    strd = refp.shape[1] + 3                                   # odd stride
    refw = np.zeros((refp.shape[0], strd), np.uint8)
    refw[:, :refp.shape[1]] = refp
    dcur = torch.zeros(cur.size + 1, dtype=torch.uint8, device=dev)
    dref = torch.zeros(refw.size + 1, dtype=torch.uint8, device=dev)
    dcur[1:] = torch.from_numpy(cur.ravel()).to(dev)            # base address + 1
    dref[1:] = torch.from_numpy(refw.ravel()).to(dev)
    cost = torch.empty((nb, side, side), dtype=torch.int32, device=dev)
    best = torch.empty((nb, 3), dtype=torch.int32, device=dev)
    for fn, oracle in ((x266.xSatd8x8SearchDev, orc.satd_search), (x266.xSad8x8SearchDev, orc.sad_search)):
        cost.zero_(); best.zero_()
        fn(dcur.data_ptr() + 1, dref.data_ptr() + 1, strd, w, h, rng_px, 0, nb, cost.data_ptr(), best.data_ptr(), st)
        torch.cuda.synchronize()
        wc, wb = oracle(cur, refp, rng_px, 0, nb)
        assert np.array_equal(cost.cpu().numpy().astype(np.uint32), wc) and np.array_equal(best.cpu().numpy(), wb)
    r = np.random.default_rng(3)
    n = 71
    refs = r.integers(0, 256, (n, 129)).astype(np.uint8)
    modes = (np.arange(n) % 35).astype(np.uint8)
    drefs = torch.zeros(n * 129 + 3, dtype=torch.uint8, device=dev)
    pred = torch.empty((n, 32, 32), dtype=torch.uint8, device=dev)
    dmodes = torch.from_numpy(modes).to(dev)
    for off in (0, 1, 2, 3):                                    # off = 0 is the aligned word path (array ends mid-word)
        drefs[off:off + n * 129] = torch.from_numpy(refs.ravel()).to(dev)
        x266.xIntra32PredDev(drefs.data_ptr() + off, dmodes.data_ptr(), pred.data_ptr(), n, st)
        torch.cuda.synchronize()
        got = pred.cpu().numpy()
        for i in range(n):
            assert np.array_equal(got[i], orc.intra32(refs[i, :64], refs[i, 64:], int(modes[i]))), (off, i)


# ------------------------------------------------------------------------- tiled frames ("next" N2)
@pytest.mark.parametrize("w,h", [(32, 32), (96, 64), (1920, 1088)])
def test_frame_residual_dct32(x266, orc, w, h):
    """fused residual + DCT32 straight from ref_block_t-tiled frames vs gather + oracle transform"""
    rng = np.random.default_rng(w)
    def frame(lo, hi):
        return orc.conv_input_fmt(rng.integers(lo, hi, (h, w)).astype(np.uint8), rng.integers(0, 256, (h // 2, w // 2)).astype(np.uint8),
                                  rng.integers(0, 256, (h // 2, w // 2)).astype(np.uint8))
    for (a, b) in (((0, 256), (0, 256)), ((255, 256), (0, 1)), ((0, 1), (255, 256))):     # random, +255 flat, -255 flat
        cur, pred = frame(*a), frame(*b)
        for sh in ((4, 11), (1, 1)):
            got = x266.xFrameResiDct32(cur, pred, w, h, *sh)
            assert np.array_equal(got, orc.frame_resi_dct32(cur, pred, w, h, *sh)), (a, b, sh)


def test_conv_input_output_fmt_dev(x266, orc):
    import torch
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(9)
    w, h = 176, 144
    Y, U, V = (rng.integers(0, 256, s).astype(np.uint8) for s in ((h, w), (h // 2, w // 2), (h // 2, w // 2)))
    dY, dU, dV = (torch.from_numpy(a).to(dev) for a in (Y, U, V))
    tiles = torch.zeros((w // 16) * (h // 16) * 512, dtype=torch.uint8, device=dev)
    x266.xConvInputFmtDev(tiles.data_ptr(), dY.data_ptr(), dU.data_ptr(), dV.data_ptr(), w, w, h)
    torch.cuda.synchronize()
    assert np.array_equal(tiles.cpu().numpy(), orc.conv_input_fmt(Y, U, V))
    oY, oU, oV = torch.zeros_like(dY), torch.zeros_like(dU), torch.zeros_like(dV)
    x266.xConvOutput420Dev(tiles.data_ptr(), oY.data_ptr(), w, oU.data_ptr(), oV.data_ptr(), w // 2, w, h)
    torch.cuda.synchronize()
    assert np.array_equal(oY.cpu().numpy(), Y) and np.array_equal(oU.cpu().numpy(), U) and np.array_equal(oV.cpu().numpy(), V)


def test_host_converters_reference_signatures(x266):
    """xConvInputFmt / xConvOutput420 with the reference's names and signatures (host pointers) against the COMPILED reference functions
    (oracle/_ref/libx266conv.so), strided planes, m_I preserved"""
    from oracle import RefConv, have_ref_conv
    if not have_ref_conv():
        pytest.skip("oracle/_ref/libx266conv.so not available")
    rc = RefConv()
    rng = np.random.default_rng(3)
    for w, h, strd in ((16, 16, 16), (176, 144, 200), (1920, 1088, 1920)):
        Yb = rng.integers(0, 256, strd * h, dtype=np.uint8)
        Ub = rng.integers(0, 256, (strd >> 1) * (h // 2), dtype=np.uint8)
        Vb = rng.integers(0, 256, (strd >> 1) * (h // 2), dtype=np.uint8)
        want = np.full((w // 16) * (h // 16) * 512, 0x5A, np.uint8)
        got = want.copy()
        rc.input_fmt(Yb, Ub, Vb, strd, w, h, want)
        x266.xConvInputFmt(got, Yb, Ub, Vb, strd, w, h)
        assert np.array_equal(got, want), (w, h, strd)
        wy = np.full(strd * h, 7, np.uint8); wu = np.full((strd >> 1) * (h // 2), 7, np.uint8); wv = wu.copy()
        gy, gu, gv = wy.copy(), wu.copy(), wv.copy()
        rc.output420(want, wy, strd, wu, wv, strd >> 1, w, h)
        x266.xConvOutput420(got, gy, strd, gu, gv, strd >> 1, w, h)
        assert np.array_equal(gy, wy) and np.array_equal(gu, wu) and np.array_equal(gv, wv), (w, h, strd)   # bytes past `width` untouched too


def test_tiled_search_equals_planar_search(x266, orc):
    """search on the encoder's own frame stores: ref_block_t frames in, in-call edge replication == planar search on the config-3 frames
    (1920x1088, the height the encoder pads 1080 to) and on small ragged cases, SATD and SAD"""
    import torch
    from search_frames import config3_frames
    for (w, h, r, full) in ((1920, 1088, 32, True), (48, 32, 8, False), (64, 16, 16, False)):
        cur, refp = config3_frames(w, h, r) if full else (np.random.default_rng(w).integers(0, 256, (h, w), dtype=np.uint8), None)
        if refp is None:
            ref = np.random.default_rng(h).integers(0, 256, (h, w), dtype=np.uint8)
            refp = np.pad(ref, r, mode="edge")
        ref = refp[r:r + h, r:r + w]
        zc = np.zeros((h // 2, w // 2), np.uint8)
        ct, rt = orc.conv_input_fmt(cur, zc, zc), orc.conv_input_fmt(np.ascontiguousarray(ref), zc, zc)
        nb = (w // 8) * (h // 8)
        if full:
            _, wb = x266.xSatd8x8Search(cur, refp, r, want_cost=False)
            _, gb = x266.xSatd8x8SearchTiled(ct, rt, w, h, r, want_cost=False)
            assert np.array_equal(gb, wb)
            sub = (1000, 1300)
            wc, _ = x266.xSatd8x8Search(cur, refp, r, *sub, want_best=False)
            gc, _ = x266.xSatd8x8SearchTiled(ct, rt, w, h, r, *sub, want_best=False)
            assert np.array_equal(gc, wc)
        else:
            wc, wb = orc.satd_search(cur, refp, r, 0, nb)
            gc, gb = x266.xSatd8x8SearchTiled(ct, rt, w, h, r)
            assert np.array_equal(gc, wc) and np.array_equal(gb, wb)
            dct, drt = torch.from_numpy(ct).cuda(), torch.from_numpy(rt).cuda()
            dbest = torch.empty((nb, 3), dtype=torch.int32, device="cuda")
            x266.xSad8x8SearchTiledDev(dct.data_ptr(), drt.data_ptr(), w, h, r, 0, nb, 0, dbest.data_ptr())
            torch.cuda.synchronize()
            assert np.array_equal(dbest.cpu().numpy(), orc.sad_search(cur, refp, r, 0, nb)[1])


# --------------------------------------------------------------------------------- SAD ("next" N4)
def test_sad_reference_golden_dataset(x266):
    import os
    from conftest import GOLDEN
    d = np.load(os.path.join(GOLDEN, "sad_dataset.npz"))
    assert x266.sad(d["a"], d["b"]) == 344807             # riscv/programs/benchmarks/sad/dataset1.h:423-426


@pytest.mark.parametrize("n", [1, 3, 64, 65, 1000])
def test_sad_region_sizes(x266, orc, n):
    r = np.random.default_rng(n)
    a, b = r.integers(0, 256, (n, n)).astype(np.uint8), r.integers(0, 256, (n, n)).astype(np.uint8)
    assert x266.sad(a, b) == orc.sad(a, b)
    assert x266.sad(a, a) == 0
    assert x266.sad(np.zeros((n, n), np.uint8), np.full((n, n), 255, np.uint8)) == 255 * n * n


@pytest.mark.parametrize("rng_px", [0, 3, 8, 16, 32])
@pytest.mark.parametrize("v1", [0, 1])
def test_sad_search(x266, orc, rng_px, v1):
    """v1 = 0: position-tile kernel (two reference windows per thread, R in {8,16,32}); v1 = 1: one CTA per block (any R)"""
    x266.tune(7, v1)
    try:
        cur, refp = make_frames(72, 40, rng_px, seed=7)
        cost, best = x266.xSad8x8Search(cur, refp, rng_px)
        wc, wb = orc.sad_search(cur, refp, rng_px, 0, 45)
        assert np.array_equal(cost, wc) and np.array_equal(best, wb)
        c, b = x266.xSad8x8Search(cur, refp, rng_px, 7, 20)
        assert np.array_equal(c, wc[7:20]) and np.array_equal(b, wb[7:20])
        _, b = x266.xSad8x8Search(cur, refp, rng_px, 7, 20, want_cost=False)
        assert np.array_equal(b, wb[7:20])
    finally:
        x266.tune(7, 0)


@pytest.mark.parametrize("rng_px", [8, 32])
def test_sad_search_wide_frame_and_ties(x266, orc, rng_px):
    """width 200 = three full position tiles + a ragged one; flat frames: every cost ties and the argmin rule decides"""
    cur, refp = make_frames(200, 24, rng_px, seed=9)
    wc, wb = orc.sad_search(cur, refp, rng_px, 0, 75)
    cost, best = x266.xSad8x8Search(cur, refp, rng_px)
    assert np.array_equal(cost, wc) and np.array_equal(best, wb)
    flat = np.full((24, 200), 200, np.uint8)
    flatp = np.full((24 + 2 * rng_px, 200 + 2 * rng_px), 10, np.uint8)
    c, b = x266.xSad8x8Search(flat, flatp, rng_px)
    assert (b[:, 1] == 0).all() and (b[:, 2] == 0).all() and (c == 64 * 190).all()


@pytest.mark.parametrize("rng_px", [3, 8, 32])
def test_search_u16_cost_surface(x266, orc, rng_px):
    """xSatd8x8SearchU16 / xSad8x8SearchU16 (+Dev, +TiledDev): the same costs as 16-bit words -- exact, since an 8x8 SATD of 8-bit pixels is
    <= 32640 and a SAD <= 16320; the all-0 vs all-255 frame reaches the SAD maximum and the largest SATD DC"""
    import torch
    cur, refp = make_frames(72, 40, rng_px, seed=11)
    for fn, oracle in ((x266.xSatd8x8Search, orc.satd_search), (x266.xSad8x8Search, orc.sad_search)):
        wc, wb = oracle(cur, refp, rng_px, 0, 45)
        c, b = fn(cur, refp, rng_px, u16=True)
        assert c.dtype == np.uint16 and np.array_equal(c, wc) and np.array_equal(b, wb)
        c, b = fn(cur, refp, rng_px, 7, 20, u16=True)
        assert np.array_equal(c, wc[7:20]) and np.array_equal(b, wb[7:20])
    zero = np.zeros((16, 32), np.uint8)
    full = np.full((16 + 2 * rng_px, 32 + 2 * rng_px), 255, np.uint8)
    c, _ = x266.xSad8x8Search(zero, full, rng_px, u16=True)
    assert (c == 16320).all()
    c, _ = x266.xSatd8x8Search(zero, full, rng_px, u16=True)
    assert (c == 4080).all()
    # the largest SATD an 8-bit search can produce: +-255 in a bent pattern (flat Hadamard spectrum) = 32640, still a uint16_t
    from test_oracle import bent_block
    pat = np.where(np.tile(bent_block(), (2, 4)) > 0, 255, 0).astype(np.uint8)                  # 16 x 32, aligned to the block grid
    inv = np.pad(255 - pat, rng_px, mode="wrap") if rng_px % 8 == 0 else np.pad(255 - pat, rng_px, mode="edge")
    c, _ = x266.xSatd8x8Search(pat, inv, rng_px, u16=True)
    wc, _ = orc.satd_search(pat, inv, rng_px, 0, 8)
    assert np.array_equal(c, wc) and c[:, rng_px, rng_px].max() == 32640 and c.max() == 32640
    # device forms, misaligned cost pointer (2-byte aligned only), and the tiled device forms
    w, h = 72, 40
    side, nb = 2 * rng_px + 1, 45
    dcur, dref = torch.from_numpy(cur).cuda(), torch.from_numpy(refp).cuda()
    for fn, oracle in ((x266.xSatd8x8SearchU16Dev, orc.satd_search), (x266.xSad8x8SearchU16Dev, orc.sad_search)):
        buf = torch.zeros(nb * side * side + 1, dtype=torch.int16, device="cuda")
        dbest = torch.empty((nb, 3), dtype=torch.int32, device="cuda")
        fn(dcur.data_ptr(), dref.data_ptr(), refp.shape[1], w, h, rng_px, 0, nb, buf.data_ptr() + 2, dbest.data_ptr())
        torch.cuda.synchronize()
        wc, wb = oracle(cur, refp, rng_px, 0, nb)
        assert np.array_equal(buf[1:].cpu().numpy().view(np.uint16).reshape(nb, side, side), wc)
        assert np.array_equal(dbest.cpu().numpy(), wb)
    if rng_px == 8:
        w, h = 48, 32
        cur = np.random.default_rng(w).integers(0, 256, (h, w), dtype=np.uint8)
        ref = np.random.default_rng(h).integers(0, 256, (h, w), dtype=np.uint8)
        refp = np.pad(ref, rng_px, mode="edge")
        zc = np.zeros((h // 2, w // 2), np.uint8)
        dct, drt = (torch.from_numpy(orc.conv_input_fmt(p, zc, zc)).cuda() for p in (cur, ref))
        nb = (w // 8) * (h // 8)
        for fn, oracle in ((x266.xSatd8x8SearchTiledU16Dev, orc.satd_search), (x266.xSad8x8SearchTiledU16Dev, orc.sad_search)):
            dcost = torch.zeros((nb, side, side), dtype=torch.int16, device="cuda")
            dbest = torch.empty((nb, 3), dtype=torch.int32, device="cuda")
            fn(dct.data_ptr(), drt.data_ptr(), w, h, rng_px, 0, nb, dcost.data_ptr(), dbest.data_ptr())
            torch.cuda.synchronize()
            wc, wb = oracle(cur, refp, rng_px, 0, nb)
            assert np.array_equal(dcost.cpu().numpy().view(np.uint16), wc) and np.array_equal(dbest.cpu().numpy(), wb)


# ------------------------------------------------------------------ fused intra mode decision ("next" N1)
@pytest.mark.parametrize("v1", [0])
def test_intra32_decide(x266, orc, v1):
    """the tensor-core decision kernel (horizontal modes on the transposed problem)"""
    x266.tune(5, v1)
    r = np.random.default_rng(12)
    n = 41                                     # odd: the last pass of the kernel decides a single block
    refs = r.integers(0, 256, (n, 129)).astype(np.uint8)
    cur = r.integers(0, 256, (n, 32, 32)).astype(np.uint8)
    # plant exact predictions so that specific modes must win with cost 0
    for i, m in enumerate((0, 1, 2, 10, 18, 26, 34, 7, 13, 21, 29)):
        cur[i] = orc.intra32(refs[i, :64], refs[i, 64:], m)
    refs[n - 1] = 255; cur[n - 1] = 0          # extreme flat residual -255
    cost, best = x266.xIntra32Decide(cur, refs)
    x266.tune(5, 0)
    for i in range(n):
        wc, wb = orc.intra32_decide(cur[i], refs[i, :64], refs[i, 64:])
        assert np.array_equal(cost[i], wc), i
        assert best[i] == wb
    for i, m in enumerate((0, 1, 2, 10, 18, 26, 34, 7, 13, 21, 29)):
        assert cost[i, m] == 0 and cost[i, best[i]] == 0


# ------------------------------------------------------------------------ inverse transform ("next" N3)
def test_idct32_vs_oracle_and_roundtrip(x266, orc, vectors):
    r = np.random.default_rng(4)
    blocks = [orc.residual(64 * 1024, 8, 2).reshape(-1, 32, 32), r.integers(-2000, 2000, (64, 32, 32)).astype(np.int16),
              np.stack([np.full((32, 32), v, np.int16) for v in (0, 1, -1, 32767, -32768)]), vectors["dct_out_4_11"]]
    for x in blocks:
        for sh in ((7, 12), (7, 10), (1, 1), (16, 16)):
            assert np.array_equal(x266.xIdct32Batch(x, *sh), orc.idct(x, 5, *sh)), sh
    # round trip at the 8-bit operating point (forward 4/11, inverse 7/12): |x - IDCT(DCT(x))| <= 4 on a full 1080p frame
    # of white-noise residuals (the integer transform pair is not lossless; 3 is the worst case seen on the oracle)
    x = orc.residual(2040 * 1024, 266, 0)
    back = x266.xIdct32Batch(x266.xDct32Batch(x, 4, 11), 7, 12)
    assert int(np.abs(back.astype(np.int32) - x).max()) <= 4
    # and at the 10-bit operating point of the bench workload (forward 6/11, inverse 7/10)
    x = orc.residual(2040 * 1024, 266, 1)
    back = x266.xIdct32Batch(x266.xDct32Batch(x, 6, 11), 7, 10)
    assert int(np.abs(back.astype(np.int32) - x).max()) <= 16


def test_multi_gpu_host_entry(x266, orc):
    """xDct32BatchMultiGpu: contiguous shards, one host thread per device (all visible devices; 1 is fine)."""
    import torch
    x = orc.residual(20000 * 1024, 77, 1)
    want = orc.dct(x.reshape(-1, 32, 32), 5, 6, 11, threads=8).ravel()
    for n in sorted({1, torch.cuda.device_count()}):
        assert np.array_equal(x266.xDct32BatchMultiGpu(x, 6, 11, n_gpus=n), want)
    with pytest.raises(x266.X266Error):
        x266.xDct32BatchMultiGpu(x, 6, 11, n_gpus=torch.cuda.device_count() + 1)


# ------------------------------------------------------------------------------ full-size properties
def test_full_size_config5_two_implementations_agree(x266, orc):
    """BASELINE config 5 at the bench size (64 frames of 8K = 2 073 600 blocks, 4.25 GB) is too big for the CPU
    oracle, so: (1) the tensor-core kernel and the independent CUDA-core butterfly kernel must agree bit for bit
    on the whole batch, (2) a deterministic sample of blocks (first, last, strided) is checked against the oracle,
    (3) the run is repeatable bit for bit."""
    import torch
    dev = torch.device("cuda:0")
    n = 64 * 32400
    g = torch.Generator(device=dev); g.manual_seed(266)
    src = (torch.randint(0, 1024, (n, 32, 32), device=dev, generator=g, dtype=torch.int16)
           - torch.randint(0, 1024, (n, 32, 32), device=dev, generator=g, dtype=torch.int16))
    a, b = torch.empty_like(src), torch.empty_like(src)
    st = torch.cuda.current_stream().cuda_stream
    x266.set_dct_variant(x266.DCT_IMMA); x266.xDct32BatchDev(src.data_ptr(), a.data_ptr(), n, 6, 11, st)
    x266.set_dct_variant(x266.DCT_BFLY); x266.xDct32BatchDev(src.data_ptr(), b.data_ptr(), n, 6, 11, st)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    x266.set_dct_variant(x266.DCT_IMMA); x266.xDct32BatchDev(src.data_ptr(), b.data_ptr(), n, 6, 11, st)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    x266.set_dct_variant(x266.DCT_AUTO)
    idx = torch.cat([torch.arange(0, 512), torch.arange(n - 512, n), torch.arange(0, n, 4099)]).to(dev)
    xs, ys = src[idx].cpu().numpy(), a[idx].cpu().numpy()
    assert np.array_equal(ys, orc.dct(xs, 5, 6, 11, threads=8))


@pytest.mark.parametrize("cfg", [0, 3, 6, 7, 12])
def test_every_imma_instantiation_is_bit_exact(x266, orc, cfg):
    """every kept (warps, stages, CTAs/SM, staging) instantiation of the tensor-core kernel (TMA rings and direct loads), ragged batch"""
    x = orc.residual(5003 * 1024, 40 + cfg, 2)
    want = orc.dct(x.reshape(-1, 32, 32), 5, 4, 11, threads=8).ravel()
    x266.set_dct_variant(x266.DCT_IMMA)
    x266.tune(0, cfg)
    try:
        assert np.array_equal(x266.xDct32Batch(x, 4, 11), want)
    finally:
        x266.tune(0, -1)
        x266.set_dct_variant(x266.DCT_AUTO)


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 5000])
def test_transpose32x32_stage(x266, n):
    """row A5 as a stand-alone op (mkTranspose32x32 on bytes): streaming 32x32 corner-turn, involution"""
    x = np.random.default_rng(n).integers(0, 256, (n, 32, 32)).astype(np.uint8)
    y = x266.xTranspose32x32Batch(x)
    assert np.array_equal(y, x.transpose(0, 2, 1))
    if n:
        assert np.array_equal(x266.xTranspose32x32Batch(y), x)


# ----------------------------------------------------------------------------- plain-C drop-in proof
def test_plain_c_bdpi_dropin(x266, ref, tmp_path):
    """tests/c/bdpi_dropin.c (C only, dlopen): the testbench call sequence gives the same digest through the
    unmodified reference library and through libx266_b200.so."""
    import os, subprocess
    from conftest import ROOT
    exe = str(tmp_path / "bdpi_dropin")
    subprocess.check_call(["gcc", "-O2", os.path.join(ROOT, "tests", "c", "bdpi_dropin.c"), "-o", exe, "-ldl"])
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libx266ref.so")
    for seed in ("1", "266"):
        want = subprocess.check_output([exe, ref_so, seed], text=True).strip()
        got = subprocess.check_output([exe, x266.LIB_PATH, seed], text=True).strip()
        assert got == want and len(got) == 16


# ------------------------------------------------------------------------------------- error paths
def test_error_convention(x266):
    """src/x266.cpp convention: int 0 / -1, no exceptions across the ABI; the binding turns -1 into X266Error."""
    import torch
    L = x266.lib()
    z = np.zeros(2048, np.int16)
    assert L.xDct32Batch(z.ctypes.data, z.ctypes.data, 2, 0, 11) == -1 and b"xDct32Batch" in L.xGpuLastError()
    assert L.xDct32Batch(None, None, 1, 4, 11) == -1
    assert L.xDct32Batch(None, None, 0, 4, 11) == 0                      # empty batch is fine
    assert L.xDctNBatch(6, z.ctypes.data, z.ctypes.data, 1, 4, 11) == -1
    d = torch.zeros(4096, dtype=torch.int16, device="cuda")
    assert L.xDct32BatchDev(d.data_ptr() + 2, d.data_ptr(), 1, 4, 11, None) == -1 and b"alignment" in L.xGpuLastError()
    with pytest.raises(x266.X266Error):
        x266.xIntra32Pred(np.zeros((1, 129), np.uint8), np.array([35], np.uint8))
    with pytest.raises(x266.X266Error):
        x266.xSatd8x8Search(np.zeros((8, 12), np.uint8), np.zeros((8, 12), np.uint8), 0)     # width not a multiple of 8
    assert L.xGpuTune(99, 0) == -1
