import os, sys
sys.path.insert(0, "/root/repo")
import torch, x266_b200 as xb
dev = torch.device("cuda:0"); n = 1 << 20
refs = torch.randint(0, 256, (n, 129), device=dev, dtype=torch.uint8)
modes = (torch.arange(n, device=dev) % 35).to(torch.uint8)
pred = torch.empty((n, 1024), device=dev, dtype=torch.uint8)
st = torch.cuda.current_stream().cuda_stream
for rep in range(2):
    for c in (8, 5, 4, 10, 16, 3):
        xb.tune(9, c)
        for _ in range(3): xb.xIntra32PredDev(refs.data_ptr(), modes.data_ptr(), pred.data_ptr(), n, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): xb.xIntra32PredDev(refs.data_ptr(), modes.data_ptr(), pred.data_ptr(), n, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"ctas/SM {c:2d}: {ms:.3f} ms  {n/ms/1e6:.3f} G pred/s", flush=True)
