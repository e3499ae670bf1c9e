"""Per-kernel-family time of the hybrid training step (DynamicUNet ch=128, the product's configuration, batch 16, 256x256):
CUDA events around every launch (ops.prof), so the families are serialised — shares, not the step time.

    python scripts/prof_hybrid.py [batch]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hdiff_b200.ops as hops  # noqa: E402
from hdiff_b200.diffusion.Diffusion import GaussianDiffusionTrainer  # noqa: E402
from hdiff_b200.diffusion.Model import DynamicUNet  # noqa: E402
from hdiff_b200.optim import FlatAdamW  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = torch.device("cuda")
    ops = hops.get()
    torch.manual_seed(0)
    net = DynamicUNet(T=1000, ch=128, ch_mult=[1, 2, 2, 2], num_res_blocks=2, dropout=0.15).to(dev)
    net.train()
    tr = GaussianDiffusionTrainer(net, 1e-4, 0.02, 1000).to(dev)
    opt = FlatAdamW(net, lr=1e-4, weight_decay=1e-4, max_grad_norm=1.0)
    g = torch.Generator(device=dev).manual_seed(1)
    gt = torch.randint(0, 256, (batch, 3, 256, 256), generator=g, device=dev, dtype=torch.uint8)
    inp = torch.randint(0, 256, (batch, 3, 256, 256), generator=g, device=dev, dtype=torch.uint8)

    def step():
        opt.zero_grad()
        (tr(gt, inp, 0)[0].sum() / 1000.).backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"step {e0.elapsed_time(e1) / 3:.2f} ms (no per-launch events)")
    ops.prof = {}
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1)
    fam = {k: (sum(r[0].elapsed_time(r[1]) for r in v), len(v), sum(r[2] for r in v)) for k, v in ops.prof.items()}
    ops.prof = None
    print(f"step with events {total:.2f} ms; families sum {sum(v[0] for v in fam.values()):.2f} ms")
    for k, (ms, n, w) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
        print(f"{k:24s} {ms:8.3f} ms  {n:4d} launches  work/ms {w / ms / 1e9 if ms else 0:10.2f} G/ms")


if __name__ == "__main__":
    main()
