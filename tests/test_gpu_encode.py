"""GPU suite, the closed intra block loop ("next" rows N1 + N3): fused decide -> predict -> residual -> DCT32 -> quantiser stub ->
IDCT32 -> reconstruction (csrc/encode.cu) against the composed oracle, the Recon channel alone, and the quantiser on its own."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _blocks(rng, n, planted=True, orc=None):
    refs = rng.integers(0, 256, (n, 129), dtype=np.uint8)
    cur = rng.integers(0, 256, (n, 1024), dtype=np.uint8)
    if planted and orc is not None:
        # most blocks: a real prediction of a known mode plus small noise, so that every mode gets decided somewhere and the
        # residual is small (the quantiser sees zeros, small and large levels)
        for i in range(n):
            if i % 4 == 3:
                continue
            m = i % 35
            p = orc.intra32(refs[i, :64], refs[i, 64:], m).astype(np.int32).ravel()
            cur[i] = np.clip(p + rng.integers(-6, 7, 1024) * (1 + i % 3), 0, 255).astype(np.uint8)
    return cur, refs


@pytest.mark.parametrize("qp", [0, 17, 22, 37, 51])
@pytest.mark.parametrize("n", [1, 7, 8, 9, 75])
def test_encode_block_fused_vs_oracle(x266, orc, qp, n):
    rng = np.random.default_rng(100 * qp + n)
    cur, refs = _blocks(rng, n, orc=orc)
    level, recon, best, cost = x266.xIntra32EncodeBlock(cur, refs, qp)
    for i in range(n):
        wl, wr, wc, wb = orc.intra32_encode(cur[i], refs[i, :64], refs[i, 64:], qp)
        assert best[i] == wb and np.array_equal(cost[i], wc), (i, best[i], wb)
        assert np.array_equal(level[i], wl), (i, "level")
        assert np.array_equal(recon[i], wr), (i, "recon")


def test_encode_block_extremes(x266, orc):
    """flat 0 / 255 blocks against 255 / 0 references (largest residual, levels near the clip), checkerboards, qp 0 and 51"""
    cur = np.stack([np.zeros(1024, np.uint8), np.full(1024, 255, np.uint8),
                    np.where(np.arange(1024) % 2 == 0, 255, 0).astype(np.uint8),
                    np.where((np.arange(1024) // 32 + np.arange(1024)) % 2 == 0, 0, 255).astype(np.uint8)])
    refs = np.stack([np.full(129, 255, np.uint8), np.zeros(129, np.uint8), np.full(129, 128, np.uint8), np.full(129, 7, np.uint8)])
    for qp in (0, 51):
        level, recon, best, cost = x266.xIntra32EncodeBlock(cur, refs, qp, want_cost=False)
        assert cost is None
        for i in range(4):
            wl, wr, _, wb = orc.intra32_encode(cur[i], refs[i, :64], refs[i, 64:], qp)
            assert best[i] == wb and np.array_equal(level[i], wl) and np.array_equal(recon[i], wr), (qp, i)


def test_recon_channel_every_mode(x266, orc):
    """mode given (the RTL's Recon channel): all 35 modes, two quantisers"""
    rng = np.random.default_rng(8)
    n = 70
    cur, refs = _blocks(rng, n, orc=orc)
    modes = (np.arange(n) % 35).astype(np.uint8)
    for qp in (10, 30):
        level, recon = x266.xIntra32Recon(cur, refs, modes, qp)
        for i in range(n):
            wl, wr = orc.intra32_recon(cur[i], refs[i, :64], refs[i, 64:], int(modes[i]), qp)
            assert np.array_equal(level[i], wl) and np.array_equal(recon[i], wr), (qp, i, int(modes[i]))
    with pytest.raises(x266.X266Error):
        x266.xIntra32Recon(cur[:1], refs[:1], np.array([35], np.uint8), 22)


def test_fused_equals_decide_then_recon_at_frame_size(x266):
    """size-independent property on a 1080p frame of blocks (2040): the fused kernel = xIntra32Decide followed by xIntra32Recon with
    its modes, device pointers, no CPU checker in the loop"""
    import torch
    n = 2040
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    cur = torch.randint(0, 256, (n, 1024), device="cuda", generator=g, dtype=torch.uint8)
    refs = torch.randint(0, 256, (n, 129), device="cuda", generator=g, dtype=torch.uint8)
    level = torch.empty((n, 1024), device="cuda", dtype=torch.int16); recon = torch.empty((n, 1024), device="cuda", dtype=torch.uint8)
    best = torch.empty(n, device="cuda", dtype=torch.int32); cost = torch.empty((n, 35), device="cuda", dtype=torch.int32)
    x266.xIntra32EncodeBlockDev(cur.data_ptr(), refs.data_ptr(), n, 27, level.data_ptr(), recon.data_ptr(), best.data_ptr(), cost.data_ptr())
    cost2 = torch.empty_like(cost); best2 = torch.empty_like(best)
    x266.xIntra32DecideDev(cur.data_ptr(), refs.data_ptr(), cost2.data_ptr(), best2.data_ptr(), n)
    level2 = torch.empty_like(level); recon2 = torch.empty_like(recon)
    x266.xIntra32ReconDev(cur.data_ptr(), refs.data_ptr(), best2.to(torch.uint8).data_ptr(), n, 27, level2.data_ptr(), recon2.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(best, best2) and torch.equal(cost, cost2) and torch.equal(level, level2) and torch.equal(recon, recon2)
    # the loop is closed: at qp 27 the reconstruction is near the source where the prediction is any good at all
    err = (recon.int() - cur.int()).abs().float().mean().item()
    assert err < 40.0


@pytest.mark.parametrize("qp", [0, 5, 22, 36, 51])
def test_quant_dequant_stub(x266, orc, qp):
    import torch
    c = np.concatenate([orc.residual(4096 * 8, 3 + qp, 2), np.array([0, 1, -1, 32767, -32768, 255, -255, 12], np.int16)])
    d = torch.from_numpy(c).cuda()
    lv = torch.empty_like(d); dq = torch.empty_like(d)
    x266.xQuantDequantDev(d.data_ptr(), lv.data_ptr(), dq.data_ptr(), c.size, qp)
    torch.cuda.synchronize()
    wl = orc.quant(c, qp)
    assert np.array_equal(lv.cpu().numpy(), wl) and np.array_equal(dq.cpu().numpy(), orc.dequant(wl, qp))
    lv.zero_()
    x266.xQuantDequantDev(d.data_ptr(), lv.data_ptr(), 0, c.size, qp)          # either output may be NULL
    torch.cuda.synchronize()
    assert np.array_equal(lv.cpu().numpy(), wl)
