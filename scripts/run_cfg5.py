"""BASELINE.json configs[4]: wider UNet ch=128 ch_mult=[1,1,2,2,4] at 512x512 (attention at the 64x64 / C=256 and 32x32 / C=512 levels,
attn=[3,4]; 16x16 does not occur with 5 levels from 512).  Functional + timing check of one configuration outside the
bench line: training steps on one GPU, loss must fall and stay finite.
    python scripts/run_cfg5.py [batch] [steps]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hdiff_b200.diffusion.Model import UNet  # noqa: E402
from hdiff_b200.diffusion.Diffusion import GaussianDiffusionTrainer  # noqa: E402
from hdiff_b200.optim import FlatAdamW  # noqa: E402
import hdiff_b200.ops as hops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda")
torch.manual_seed(0)
net = UNet(T=1000, ch=128, ch_mult=[1, 1, 2, 2, 4], attn=[3, 4], num_res_blocks=2, dropout=0.1).to(dev)
nparam = sum(p.numel() for p in net.parameters())
tr = GaussianDiffusionTrainer(net, 1e-4, 0.02, 1000).to(dev)
opt = FlatAdamW(net, lr=1e-4, weight_decay=1e-4, max_grad_norm=1.0)
x = torch.rand(B, 3, 512, 512, device=dev) * 2 - 1
losses, ts = [], []
for s in range(steps):
    torch.cuda.synchronize(); t0 = time.time()
    opt.zero_grad()
    loss = tr(x).sum() / 1000.
    loss.backward()
    opt.step()
    torch.cuda.synchronize(); ts.append(time.time() - t0)
    losses.append(float(loss))
ops = hops.get()
# one more step with per-launch CUDA events: where the step goes, by kernel family
ops.prof = {}
opt.zero_grad(); loss = tr(x).sum() / 1000.; loss.backward(); opt.step()
torch.cuda.synchronize()
prof, ops.prof = ops.prof, None
fam = {k: {"launches": len(v), "ms": round(sum(r[0].elapsed_time(r[1]) for r in v), 3)} for k, v in prof.items()}
print(json.dumps({"config": "cfg5 uncond UNet ch=128 [1,1,2,2,4] attn=[3,4] nrb=2, 512x512", "params": nparam, "batch": B,
                  "ms_per_step_median": sorted(ts[1:])[len(ts[1:]) // 2] * 1e3, "images_per_s": B / sorted(ts[1:])[len(ts[1:]) // 2],
                  "losses": losses, "finite": all(l == l and abs(l) < 1e9 for l in losses),
                  "tcgen05_launches": ops.tc_launches, "launches": ops.launches,
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                  "families_ms": dict(sorted(fam.items(), key=lambda kv: -kv[1]["ms"]))}))
