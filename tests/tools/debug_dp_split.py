"""Gradients of a batch of 4 against the sum of the gradients of its two halves (what data parallelism computes), both
against the fp32 oracle: which side of tests/dp_nccl_worker.py's comparison is off?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hdiff_b200.DiffusionFreeGuidence.ModelCondition import UNet
from oracle import ref_torch as R
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda")
cfg = dict(T=1000, ch=64, ch_mult=[1, 2, 2], attn=[1], num_res_blocks=1, dropout=0.0)
res, b, world = 64, 2, 2
torch.manual_seed(100)
net = UNet(num_labels=10, **cfg).to(dev).train()
ref = R.UNet(num_labels=10, **cfg).to(dev).train()
ref.load_state_dict(net.state_dict())
g = torch.Generator(device="cpu").manual_seed(5)
x = (torch.rand(world * b, 3, res, res, generator=g) * 2 - 1).to(dev)
t = torch.randint(0, 1000, (world * b,), generator=g).to(dev)
lab = (torch.randint(0, 10, (world * b,), generator=g) + 1).to(dev)
lab[:b] = 0
def grads(model, xs, ts, ls, scale):
    model.zero_grad()
    ((model(xs, ts, ls) ** 2).sum() * scale).backward()
    return {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
full = grads(net, x, t, lab, 1.0 / b ** 2 / world)
h0 = grads(net, x[:b], t[:b], lab[:b], 1.0 / b ** 2)
h1 = grads(net, x[b:], t[b:], lab[b:], 1.0 / b ** 2)
halves = {k: (h0[k] + h1[k]) / world for k in h0}
oracle = grads(ref, x, t, lab, 1.0 / b ** 2 / world)
def rel(a, b_): return float((a - b_).norm() / (b_.norm() + 1e-30))
print(f"{'parameter':44s} full-vs-halves  full-vs-oracle  halves-vs-oracle")
for k in full:
    r1, r2, r3 = rel(full[k], halves[k]), rel(full[k], oracle[k]), rel(halves[k], oracle[k])
    if max(r1, r2, r3) > 5e-3:
        print(f"{k:44s} {r1:12.4f} {r2:14.4f} {r3:16.4f}")
