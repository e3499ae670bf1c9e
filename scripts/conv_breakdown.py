"""Where the convolution / wgrad time of one cfg2 training step goes, by shape (CUDA events per launch).
    python scripts/conv_breakdown.py [batch] > gpurun_out/conv_breakdown.json"""
import collections
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hdiff_b200.diffusion.Model import UNet  # noqa: E402
from hdiff_b200.diffusion.Diffusion import GaussianDiffusionTrainer  # noqa: E402
import hdiff_b200.ops as hops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as _f:
    _pk = json.load(_f)
PEAK_TFLOPS, PEAK_GBS = _pk["bf16_tflops_sustained"], _pk["hbm_gbs"]      # GFLOP / (TFLOP/s) = ms, MB / (GB/s) = ms
dev = torch.device("cuda")
torch.manual_seed(0)
net = UNet(T=1000, ch=64, ch_mult=[1, 2, 2, 2], attn=[1], num_res_blocks=2, dropout=0.1).to(dev)
tr = GaussianDiffusionTrainer(net, 1e-4, 0.02, 1000).to(dev)
x = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1
ops = hops.get()
shapes = {"conv_tc": [], "wgrad_tc": []}
conv0, wgrad0 = ops.conv, ops.wgrad


def conv(x0, x1, P_in, w, bias, emb, res, out, P_out, N, H, W, k, **kw):
    n0 = len(ops.prof.get("conv_tc", [])) if ops.prof is not None else 0
    r = conv0(x0, x1, P_in, w, bias, emb, res, out, P_out, N, H, W, k, **kw)
    if ops.prof is not None and len(ops.prof.get("conv_tc", [])) > n0:
        shapes["conv_tc"].append(f"{k}x{k} {H}x{W} {x0.shape[-1]}+{0 if x1 is None else x1.shape[-1]}(P{P_in})->{out.shape[-1] if out.dim() == 4 and out.shape[-1] > 8 else 'nchw'}(P{P_out})"
                                 + ("+res" if res is not None else "") + ("+emb" if emb is not None else ""))
    return r


def wgrad(x0, x1, P_in, dy, P_dy, dw, N, H, W, k, dtype, **kw):
    n0 = len(ops.prof.get("wgrad_tc", [])) if ops.prof is not None else 0
    r = wgrad0(x0, x1, P_in, dy, P_dy, dw, N, H, W, k, dtype, **kw)
    if ops.prof is not None and len(ops.prof.get("wgrad_tc", [])) > n0:
        shapes["wgrad_tc"].append(f"{k}x{k} {H}x{W} {x0.shape[-1]}+{0 if x1 is None else x1.shape[-1]}(P{P_in})->{dy.shape[-1]}(P{P_dy})")
    return r


ops.conv, ops.wgrad = conv, wgrad
for it in range(3):
    ops.prof = {} if it == 2 else None
    for k in shapes: shapes[k].clear()
    net.zero_grad()
    loss = tr(x).sum() / 1000.
    loss.backward()
torch.cuda.synchronize()
prof, ops.prof = ops.prof, None
out = {}
for fam in shapes:
    agg = collections.OrderedDict()
    assert len(shapes[fam]) == len(prof[fam]), (fam, len(shapes[fam]), len(prof[fam]))
    for s, r in zip(shapes[fam], prof[fam]):
        a, b, w = r[:3]
        e = agg.setdefault(s, {"n": 0, "ms": 0.0, "gflop": 0.0, "mbyte": 0.0})
        e["n"] += 1; e["ms"] += a.elapsed_time(b); e["gflop"] += w / 1e9; e["mbyte"] += (r[3] if len(r) > 3 else 0.0) / 1e6
    bound_total = 0.0
    for e in agg.values():
        # the roofline that binds this shape: the longer of FLOPs / sustained tensor peak and compulsory bytes / copy bandwidth
        t_tensor, t_hbm = e["gflop"] / PEAK_TFLOPS, e["mbyte"] / PEAK_GBS          # both in ms
        e["bound"] = "hbm" if t_hbm > t_tensor else "tensor"
        e["frac_of_binding_roofline"] = round(max(t_tensor, t_hbm) / e["ms"], 3)
        bound_total += max(t_tensor, t_hbm)
        e["tflops"] = round(e["gflop"] / e["ms"], 1); e["gbs"] = round(e["mbyte"] / e["ms"], 1)
        e["ms"] = round(e["ms"], 3); e["gflop"] = round(e["gflop"], 1); e["mbyte"] = round(e["mbyte"], 1)
    out[fam + "_frac_of_binding_roofline"] = round(bound_total / sum(e["ms"] for e in agg.values()), 3)
    out[fam] = dict(sorted(agg.items(), key=lambda kv: -kv[1]["ms"]))
    out[fam + "_total_ms"] = round(sum(e["ms"] for e in agg.values()), 3)
print(json.dumps(out, indent=1))
