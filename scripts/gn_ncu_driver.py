"""One launch group of each GroupNorm kernel at the L0 layer shape (batch 32, 256x256, 64 channels) for `ncu --set full`:
    ncu --set full --clock-control none --import-source on -k regex:gn -o gpurun_out/gn python scripts/gn_ncu_driver.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hdiff_b200.ops as hops  # noqa: E402

ops = hops.get()
dev = torch.device("cuda")
bf = torch.bfloat16
N, HW, C = 32, 65536, 64
x = torch.randn(N, HW, 1, C, device=dev).to(bf)
gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
sums = torch.empty(N, 32, 2, dtype=torch.float64, device=dev)
out = torch.empty_like(x)
dy = torch.randn(N, HW, 1, C, device=dev).to(bf)
gs = torch.empty_like(sums)
dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
dx = torch.empty_like(x)
for rep in range(2):
    ops.gn_stats(x, None, N, HW, 32, sums)
    for p in (0.0, 0.1):
        ops.gn_apply(x, None, N, HW, 32, sums, gamma, beta, 1e-5, 1, p, 123, out)
        ops.gn_bwd(x, None, N, HW, 32, sums, gamma, beta, 1e-5, 1, p, 123, dy, gs, dg, db, None, None, None, dx, None)
torch.cuda.synchronize()
print("done")
