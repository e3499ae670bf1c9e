"""Every tensor-core kernel of the cfg2 step at its layer shapes (batch 32), each case launched twice, for
    ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none -k regex:"conv_tc_kernel|wgrad_tc_kernel|attn_fwd2|attn_bwd_tc" --csv --log-file gpurun_out/tc.csv python scripts/tc_ncu_driver.py
The case names go to gpurun_out/tc_cases.txt in launch order (scripts/tc_ncu_table.py joins them with the CSV)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hdiff_b200.ops as hops  # noqa: E402

ops = hops.get()
dev = torch.device("cuda")
bf = torch.bfloat16
cases = []
N = 32
CONVS = [  # H, C0, C1, Cout, k, res, emb   (the shapes of profiles/r01_conv_breakdown_by_shape.json that carry most of the time)
    (256, 64, 0, 64, 3, False, False), (256, 64, 0, 64, 3, True, False), (256, 64, 0, 64, 3, False, True),
    (128, 128, 0, 128, 3, False, False), (128, 128, 0, 128, 3, True, False), (256, 128, 0, 128, 3, False, False),
    (256, 64, 0, 128, 3, False, False), (256, 128, 0, 64, 3, False, True), (256, 128, 64, 64, 3, False, True),
    (128, 256, 0, 128, 3, False, True), (64, 128, 0, 128, 3, False, False), (32, 128, 0, 128, 3, False, False),
    (256, 64, 0, 128, 1, False, False), (256, 64, 64, 64, 1, False, False), (128, 128, 0, 384, 1, False, False),
    (128, 384, 0, 128, 1, False, False), (128, 128, 0, 128, 1, True, False),
]
_only = os.environ.get("HDIFF_TC_CASES")
if _only:
    CONVS = [CONVS[int(i)] for i in _only.split(",")]
for (H, C0, C1, Cout, k, res, emb) in CONVS:
    W = H
    x0 = torch.randn(N, H, W, C0, device=dev).to(bf)
    x1 = torch.randn(N, H, W, C1, device=dev).to(bf) if C1 else None
    Cin = C0 + C1
    w = (torch.randn(Cout * k * k * Cin, device=dev) / (k * k * Cin) ** 0.5).to(bf)
    bias = torch.randn(Cout, device=dev)
    out = torch.empty(N, H, W, Cout, device=dev, dtype=bf)
    r = torch.randn(N, H, W, Cout, device=dev).to(bf) if res else None
    e = torch.randn(N, Cout, device=dev) if emb else None
    for _ in range(2):
        ops.conv(x0, x1, 1, w, bias, e, r, out, 1, N, H, W, k)
    cases.append(f"conv {k}x{k} {H}x{W} {C0}+{C1}->{Cout}{'+res' if res else ''}{'+emb' if emb else ''} gflop={2.0 * N * H * W * Cout * k * k * Cin / 1e9:.1f}")
    if not res and not emb:
        dy = torch.randn(N, H, W, Cout, device=dev).to(bf)
        dw = torch.empty(Cout * k * k * Cin, device=dev)
        need = ops.lib.hd_wgrad_tc_workspace(C0, C1, 1, Cout, 1, N, H, W, k)
        ws = torch.empty((need + 3) // 4, device=dev)
        for _ in range(2):
            ops.wgrad(x0, x1, 1, dy, 1, dw, N, H, W, k, bf, workspace=ws)
        cases.append(f"wgrad {k}x{k} {H}x{W} {C0}+{C1}->{Cout} gflop={2.0 * N * H * W * Cout * k * k * Cin / 1e9:.1f}")
        del dy
    del x0, x1, out, r
for S in (() if _only else (16384, 1024)):
    C = 128
    qkv = torch.randn(N, S, 3 * C, device=dev).to(bf)
    out = torch.empty(N, S, C, dtype=bf, device=dev)
    dout = torch.randn(N, S, C, device=dev).to(bf)
    dqkv = torch.empty_like(qkv)
    lse = torch.empty(N, S, device=dev)
    for _ in range(2):
        ops.attn_fwd(qkv, out, lse, N, S, C)
    cases.append(f"attn_fwd N{N} S{S} gflop={4.0 * N * S * S * C / 1e9:.1f}")
    for _ in range(2):
        ops.attn_bwd(qkv, out, dout, lse, None, dqkv, N, S, C)
    cases.append(f"attn_bwd(dKdV) N{N} S{S} gflop={8.0 * N * S * S * C / 1e9:.1f} (both passes)")
    cases.append(f"attn_bwd(dQ) N{N} S{S}")
torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "tc_cases.txt"), "w") as f:
    f.write("\n".join(cases) + "\n")
print("done", len(cases))
