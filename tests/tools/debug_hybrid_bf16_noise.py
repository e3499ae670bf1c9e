"""Which DynamicUNet parameter gradients are noisiest in bf16, and is that the arithmetic or the kernels?  Compares, against the
reference module in fp32: (a) the CUDA library in bf16, (b) the same reference module under torch.autocast(bfloat16).

    python tests/tools/debug_hybrid_bf16_noise.py            # needs a GPU and oracle/_ref
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def main():
    from hdiff_b200.diffusion.Model import DynamicUNet
    dev = torch.device("cuda")
    dm = ref_loader.diffusion_model()
    cfg = dict(T=1000, ch=64, ch_mult=[1, 2, 2, 2], num_res_blocks=2, dropout=0.0)
    torch.manual_seed(32)
    ref = dm.DynamicUNet(**cfg).to(dev)
    with torch.no_grad():
        ref.tail[-1].weight.mul_(3e4)
    net = DynamicUNet(compute_dtype=torch.bfloat16, **cfg)
    net.load_state_dict(ref.state_dict())
    net.to(dev)
    x = torch.rand(2, 6, 64, 64, device=dev) * 2 - 1
    x[:, 2] += 0.3
    t = torch.tensor([9, 600], device=dev)
    lab = torch.rand(2, 3, 64, 64, device=dev) * 2 - 1
    gy = None
    grads = {}
    for name in ("fp32", "autocast", "hdiff"):
        for p in ref.parameters():
            p.grad = None
        if name == "hdiff":
            e = net(x, t, lab, context_zero=False)
        elif name == "autocast":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                e = ref(x, t, lab, context_zero=False)
            e = e.float()
        else:
            e = ref(x, t, lab, context_zero=False)
        if gy is None:
            gy = torch.randn_like(e)
        e.backward(gy)
        src = net if name == "hdiff" else ref
        grads[name] = {k: p.grad.detach().clone() for k, p in src.named_parameters() if p.grad is not None}
    gscale = max(float(g.norm()) for g in grads["fp32"].values())
    rows = []
    for k, g in grads["fp32"].items():
        rows.append((rel(grads["hdiff"][k], g), rel(grads["autocast"][k], g), float(g.norm()) / gscale, k))
    rows.sort(reverse=True)
    print(f"{'hdiff bf16':>11s} {'autocast':>9s} {'|g|/max|g|':>11s}  parameter")
    for r in rows[:12]:
        print(f"{r[0]:11.4f} {r[1]:9.4f} {r[2]:11.2e}  {r[3]}")
    import statistics
    print("median", statistics.median(r[0] for r in rows), statistics.median(r[1] for r in rows))


if __name__ == "__main__":
    main()
