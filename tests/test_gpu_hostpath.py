"""GPU suite, host-pointer pipeline (csrc/ffi.cu: run_chunked): the result must not depend on the kind of caller
memory (pageable as src/x266.cpp:505,647-649 allocates it, pinned, registered), on the staging mode, on the chunk
size, or on how many host threads share the device."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _aligned(nbytes, align=4096, skew=0):
    raw = np.empty(nbytes + align + skew, np.uint8)
    off = (-raw.ctypes.data) % align + skew
    return raw[off:off + nbytes]


@pytest.fixture
def small_chunks(x266):
    x266.tune(4, 64)          # 64 blocks per chunk: many chunks, several laps of the slot ring
    yield
    x266.tune(4, 0)
    x266.tune(12, 0)
    x266.tune(13, 0)
    x266.tune(14, 1)


@pytest.mark.parametrize("n", [1, 63, 64, 65, 127, 128, 129, 64 * 4, 64 * 4 + 1, 64 * 9 + 17, 64 * 40])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_dct32_pageable_modes_chunk_laps(x266, orc, small_chunks, n, mode):
    """staged ring / driver staging / register-per-call give the oracle's bytes for 1, LAG, SLOTS and many chunks"""
    x266.tune(12, mode)
    x = _aligned(n * 2048, skew=2 if n % 2 else 0).view(np.int16)
    x[:] = orc.residual(n * 1024, 500 + n, 2)
    y = _aligned(n * 2048).view(np.int16)
    y[:] = 0x5555
    x266.xDct32Batch(x, 6, 11, out=y)
    assert np.array_equal(y, orc.dct(x.reshape(-1, 32, 32), 5, 6, 11, threads=8).ravel())


@pytest.mark.parametrize("threads,nt", [(1, 1), (3, 1), (8, 0), (8, 1)])
def test_dct32_staging_threads_and_copy_kind(x266, orc, threads, nt):
    x266.tune(13, threads)
    x266.tune(14, nt)
    try:
        n = 3 * 16384 + 777                                   # default chunking, > SLOTS chunks, multi-MiB slices
        x = orc.residual(n * 1024, 77, 1)
        y = x266.xDct32Batch(x, 6, 11)
        assert np.array_equal(y, orc.dct(x.reshape(-1, 32, 32), 5, 6, 11, threads=8).ravel())
    finally:
        x266.tune(13, 0)
        x266.tune(14, 1)


def test_pinned_registered_and_mixed_buffers(x266, orc):
    import torch
    n = 2 * 16384 + 5
    x = orc.residual(n * 1024, 5, 1)
    want = orc.dct(x.reshape(-1, 32, 32), 5, 4, 11, threads=8).ravel()
    pin_in = torch.from_numpy(x.copy()).pin_memory()
    pin_out = torch.empty_like(pin_in).pin_memory()
    for src, dst in ((pin_in.numpy(), pin_out.numpy()), (x, pin_out.numpy()), (pin_in.numpy(), np.empty_like(x))):
        dst[:] = 0
        x266.xDct32Batch(src, 4, 11, out=dst)
        assert np.array_equal(dst, want)
    a, b = _aligned(n * 2048).view(np.int16), _aligned(n * 2048).view(np.int16)
    a[:] = x
    b[:] = 0
    x266.host_register(a)
    x266.host_register(b)
    try:
        x266.xDct32Batch(a, 4, 11, out=b)
        assert np.array_equal(b, want)
    finally:
        x266.host_unregister(a)
        x266.host_unregister(b)
    with pytest.raises(x266.X266Error):
        x266.host_unregister(a)                                   # not registered any more -> -1 + message, no crash


def test_multi_array_entry_points_pageable(x266, orc, small_chunks):
    """intra (2 inputs), decide (2 in, 2 out), SATD batch (128 B in / 4 B out) through the same staged pipeline"""
    rng = np.random.default_rng(9)
    n = (1 << 15) * 2 + 333
    refs = rng.integers(0, 256, (n, 129), dtype=np.uint8)
    modes = (np.arange(n) % 35).astype(np.uint8)
    got = x266.xIntra32Pred(refs, modes)
    idx = np.r_[0:70, (1 << 15) - 3:(1 << 15) + 3, n - 70:n]
    for i in idx:
        assert np.array_equal(got[i], orc.intra32(refs[i, :64], refs[i, 64:], int(modes[i]))), i
    m = (1 << 14) + 100
    cur = rng.integers(0, 256, (m, 1024), dtype=np.uint8)
    cost, best = x266.xIntra32Decide(cur, refs[:m])
    jdx = np.r_[0:4, (1 << 14) - 2:(1 << 14) + 2, m - 4:m]
    for j in jdx:
        wc, wb = orc.intra32_decide(cur[j], refs[j, :64], refs[j, 64:])
        assert np.array_equal(cost[j], wc) and best[j] == wb, j
    d = orc.residual(((1 << 17) * 4 + 9) * 64, 3, 2)
    assert np.array_equal(x266.xSatd8x8Batch(d), orc.satd(d, threads=8))


def test_search_outputs_staged(x266, orc):
    rng = np.random.default_rng(5)
    w, h, r = 64, 48, 8
    cur = rng.integers(0, 256, (h, w), dtype=np.uint8)
    refp = rng.integers(0, 256, (h + 2 * r, w + 2 * r), dtype=np.uint8)
    nb = (w // 8) * (h // 8)
    cost, best = x266.xSatd8x8Search(cur, refp, r)
    wc, wb = orc.satd_search(cur, refp, r, 0, nb)
    assert np.array_equal(cost, wc) and np.array_equal(best, wb)
    c2, b2 = x266.xSad8x8Search(cur, refp, r, want_cost=False)
    assert c2 is None and np.array_equal(b2, orc.sad_search(cur, refp, r, 0, nb)[1])


def test_two_host_threads_share_one_gpu(x266, orc):
    """each call leases its own pipeline: concurrent callers on one device neither corrupt each other nor deadlock"""
    n = 16384 + 11
    xs = [orc.residual(n * 1024, 900 + i, 1) for i in range(6)]
    want = [orc.dct(x.reshape(-1, 32, 32), 5, 6, 11, threads=8).ravel() for x in xs]
    got = [None] * len(xs)
    errs = []

    def work(i):
        try:
            import torch
            torch.cuda.set_device(0)
            for _ in range(3):
                got[i] = x266.xDct32Batch(xs[i], 6, 11)
        except Exception as e:                                     # surfaced below
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(xs))]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not errs, errs
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_dev_mode_check_flag(x266):
    import torch
    refs = torch.randint(0, 256, (64, 129), dtype=torch.uint8, device="cuda")
    modes = torch.full((64,), 35, dtype=torch.uint8, device="cuda")
    pred = torch.empty((64, 1024), dtype=torch.uint8, device="cuda")
    x266.xIntra32PredDev(refs.data_ptr(), modes.data_ptr(), pred.data_ptr(), 64, 0)          # default: trusted
    x266.tune(15, 1)
    try:
        with pytest.raises(x266.X266Error, match="mode"):
            x266.xIntra32PredDev(refs.data_ptr(), modes.data_ptr(), pred.data_ptr(), 64, 0)
    finally:
        x266.tune(15, 0)


def test_free_then_reuse(x266, orc):
    """xGpuFree releases pipelines, pool and tables; the next call re-initialises the device"""
    x = orc.residual(100 * 1024, 1, 1)
    want = orc.dct(x.reshape(-1, 32, 32), 5, 4, 11).ravel()
    assert np.array_equal(x266.xDct32Batch(x, 4, 11), want)
    x266.lib().xGpuFree()
    assert np.array_equal(x266.xDct32Batch(x, 4, 11), want)
    refs = np.random.default_rng(2).integers(0, 256, (35, 129), dtype=np.uint8)
    modes = np.arange(35, dtype=np.uint8)
    got = x266.xIntra32Pred(refs, modes)
    for i in range(35):
        assert np.array_equal(got[i], orc.intra32(refs[i, :64], refs[i, 64:], i))
