import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import x266_b200 as xb
torch.cuda.set_device(0)
n = 1 << 20
g = torch.Generator(device="cuda"); g.manual_seed(1)
refs = torch.randint(0, 256, (n, 129), device="cuda", generator=g, dtype=torch.uint8)
pred = torch.empty((n, 1024), device="cuda", dtype=torch.uint8)
inter = (torch.arange(n, device="cuda") % 35).to(torch.uint8)
m26 = torch.full((n,), 26, device="cuda", dtype=torch.uint8)
for modes in (inter, m26, inter, m26):
    xb.xIntra32PredDev(refs.data_ptr(), modes.data_ptr(), pred.data_ptr(), n, 0)
torch.cuda.synchronize()
