#!/usr/bin/env python
"""Sweeps the IMMA kernel instantiations (xGpuTune key 0) on the config-5 workload; prints GB/s per config."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import x266_b200 as xb

CFG = {0: "W8 S4 B2 TMA ring", 3: "W8 S6 B2 TMA ring", 6: "W8 B2 direct (shipped)", 7: "W8 B3 direct", 12: "W8 B2 direct, two blocks ahead"}
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
n = frames * 32400
src = (torch.randint(0, 1024, (n, 32, 32), device=dev, dtype=torch.int16) - torch.randint(0, 1024, (n, 32, 32), device=dev, dtype=torch.int16))
dst = torch.empty_like(src)
ref = torch.empty_like(src)
st = torch.cuda.current_stream().cuda_stream
xb.set_dct_variant(xb.DCT_BFLY)
xb.xDct32BatchDev(src.data_ptr(), ref.data_ptr(), n, 6, 11, st)
xb.set_dct_variant(xb.DCT_IMMA)
res = {}
for rep in range(1):
    for cfg, name in CFG.items():
        xb.tune(0, cfg)
        dst.zero_()
        for _ in range(3):
            xb.xDct32BatchDev(src.data_ptr(), dst.data_ptr(), n, 6, 11, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            xb.xDct32BatchDev(src.data_ptr(), dst.data_ptr(), n, 6, 11, st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        ok = bool(torch.equal(dst, ref))
        gbs = n * 4096 / (ms * 1e-3) / 1e9
        res.setdefault(cfg, []).append(gbs)
        print(f"rep{rep} cfg {cfg:2d} {name:16s} {ms:7.3f} ms  {gbs:7.1f} GB/s  {gbs / 6459.3 * 100:5.1f}% of measured  bit-exact={ok}", flush=True)
xb.tune(0, -1)
json.dump({CFG[k]: v for k, v in res.items()}, open("gpurun_out/tune_dct.json", "w"), indent=1)
