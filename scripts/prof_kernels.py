"""Micro-driver for ncu: one launch group of each hot kernel at its cfg2 shape (batch 32, 256x256), CUDA-event timed.
    python scripts/prof_kernels.py [conv|wgrad|attn|gn|all]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hdiff_b200.ops as hops  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ops = hops.get()
dev = torch.device("cuda")
bf = torch.bfloat16


def timeit(name, fn, work, unit):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:44s} {ms:8.3f} ms   {work / ms / 1e6:9.1f} {unit}", flush=True)


if what in ("conv", "wgrad", "all"):
    for (N, H, W, C0, C1, Cout, k) in ((32, 256, 256, 64, 0, 64, 3), (32, 128, 128, 128, 0, 128, 3), (32, 256, 256, 128, 64, 64, 3),
                                      (32, 64, 64, 128, 0, 128, 3), (32, 128, 128, 128, 0, 384, 1)):
        x0 = torch.randn(N, H, W, C0, device=dev).to(bf)
        x1 = torch.randn(N, H, W, C1, device=dev).to(bf) if C1 else None
        Cin = C0 + C1
        w = (torch.randn(Cout * k * k * Cin, device=dev) / (k * k * Cin) ** 0.5).to(bf)
        bias = torch.randn(Cout, device=dev)
        out = torch.empty(N, H, W, Cout, device=dev, dtype=bf)
        fl = 2.0 * N * H * W * Cout * k * k * Cin
        if what in ("conv", "all"):
            timeit(f"conv{k}x{k} N{N} {H}x{W} {C0}+{C1}->{Cout}", lambda: ops.conv(x0, x1, 1, w, bias, None, None, out, 1, N, H, W, k), fl / 1e3, "TFLOP/s")
        if what in ("conv", "all") and k == 3:
            res_t = torch.randn(N, H, W, Cout, device=dev).to(bf)
            emb_t = torch.randn(N, Cout, device=dev)
            timeit(f"conv{k}x{k}+res+emb N{N} {H}x{W} {C0}+{C1}->{Cout}", lambda: ops.conv(x0, x1, 1, w, bias, emb_t, res_t, out, 1, N, H, W, k), fl / 1e3, "TFLOP/s")
        if what in ("conv", "all") and k == 3 and False:
            cs = torch.zeros(N, Cout, 2, dtype=torch.float64, device=dev)
            timeit(f"conv{k}x{k}+stats N{N} {H}x{W} {C0}+{C1}->{Cout}", lambda: ops.conv(x0, x1, 1, w, bias, None, None, out, 1, N, H, W, k, chan_sums=cs), fl / 1e3, "TFLOP/s")
        if what in ("wgrad", "all"):
            dy = torch.randn(N, H, W, Cout, device=dev).to(bf)
            dw = torch.empty(Cout * k * k * Cin, device=dev)
            need = ops.lib.hd_wgrad_tc_workspace(C0, C1, 1, Cout, 1, N, H, W, k)
            ws = torch.empty((need + 3) // 4, device=dev)
            timeit(f"wgrad{k}x{k} N{N} {H}x{W} {C0}+{C1}->{Cout}", lambda: ops.wgrad(x0, x1, 1, dy, 1, dw, N, H, W, k, bf, workspace=ws), fl / 1e3, "TFLOP/s")
        del x0, x1, out

if what in ("attn", "attn32", "all"):
    for N, S in (((32, 16384),) if what == "attn32" else ((8, 16384), (32, 1024))):
        C = 128
        qkv = torch.randn(N, S, 3 * C, device=dev).to(bf)
        out = torch.empty(N, S, C, dtype=bf, device=dev)
        dout = torch.randn(N, S, C, device=dev).to(bf)
        dqkv = torch.empty_like(qkv)
        lse = torch.empty(N, S, device=dev)
        timeit(f"attn fwd N{N} S{S}", lambda: ops.attn_fwd(qkv, out, lse, N, S, C), 4.0 * N * S * S * C / 1e3, "TFLOP/s")
        timeit(f"attn bwd N{N} S{S}", lambda: ops.attn_bwd(qkv, out, dout, lse, None, dqkv, N, S, C), 8.0 * N * S * S * C / 1e3, "TFLOP/s")

if what in ("attnwide",):
    for N, S, C in ((4, 4096, 256), (4, 1024, 512), (32, 4096, 256), (8, 4096, 128)):
        qkv = torch.randn(N, S, 3 * C, device=dev).to(bf)
        out = torch.empty(N, S, C, dtype=bf, device=dev)
        dout = torch.randn(N, S, C, device=dev).to(bf)
        dqkv = torch.empty_like(qkv)
        lse = torch.empty(N, S, device=dev)
        timeit(f"attn fwd N{N} S{S} C{C}", lambda: ops.attn_fwd(qkv, out, lse, N, S, C), 4.0 * N * S * S * C / 1e3, "TFLOP/s")
        timeit(f"attn bwd N{N} S{S} C{C}", lambda: ops.attn_bwd(qkv, out, dout, lse, None, dqkv, N, S, C), 8.0 * N * S * S * C / 1e3, "TFLOP/s")

if what in ("gn", "all"):
    for (N, HW, C0, C1, p_drop) in ((32, 65536, 64, 0, 0.0), (32, 65536, 64, 0, 0.1), (32, 16384, 128, 0, 0.1), (32, 65536, 128, 64, 0.0)):
        C = C0 + C1
        x0 = torch.randn(N, HW, 1, C0, device=dev).to(bf)
        x1 = torch.randn(N, HW, 1, C1, device=dev).to(bf) if C1 else None
        gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
        sums = torch.empty(N, 32, 2, dtype=torch.float64, device=dev)
        out = torch.empty(N, HW, 1, C, dtype=bf, device=dev)
        numel = N * HW * C
        timeit(f"gn_stats N{N} HW{HW} C{C0}+{C1}", lambda: ops.gn_stats(x0, x1, N, HW, 32, sums), numel * 2.0, "GB/s")
        timeit(f"gn_apply N{N} HW{HW} C{C0}+{C1} p{p_drop}", lambda: ops.gn_apply(x0, x1, N, HW, 32, sums, gamma, beta, 1e-5, 1, p_drop, 123, out), numel * 4.0, "GB/s")
        dy = torch.randn(N, HW, 1, C, device=dev).to(bf)
        gs = torch.empty_like(sums)
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        dx0 = torch.empty_like(x0)
        dx1 = None if x1 is None else torch.empty_like(x1)
        timeit(f"gn_bwd N{N} HW{HW} C{C0}+{C1} p{p_drop}",
               lambda: ops.gn_bwd(x0, x1, N, HW, 32, sums, gamma, beta, 1e-5, 1, p_drop, 123, dy, gs, dg, db, None, None, None, dx0, dx1),
               numel * 10.0, "GB/s")
        t = torch.randn(N, HW, C, device=dev).to(bf)
        pn = torch.zeros(N, C, device=dev)
        tot = torch.zeros(C, device=dev)
        timeit(f"colsum N{N} HW{HW} C{C}", lambda: ops.colsum(t, N, HW, C, pn, tot), numel * 2.0, "GB/s")
        del x0, x1, out, dy, dx0, dx1, t
if what in ("mha", "all"):
    # the 8-head attention core at DynamicUNet's middle blocks (product configuration: 32x32, C = 256 -> hd = 32, batch 16) and at the
    # live ModelCondition UNet's CIFAR-sized levels (32x32 C = 128 -> hd = 16; 16x16 C = 256)
    for (N, S, C) in ((16, 1024, 256), (32, 1024, 128), (32, 256, 256), (8, 4096, 64)):
        qkv = torch.randn(N, S, 3 * C, device=dev).to(bf)
        out = torch.empty(N, S, C, dtype=bf, device=dev)
        dout = torch.randn(N, S, C, device=dev).to(bf)
        dqkv = torch.empty_like(qkv)
        lse = torch.empty(N, 8, S, device=dev)
        delta = torch.empty(N, 8, S, device=dev)
        timeit(f"mha fwd N{N} S{S} C{C} (8 heads)", lambda: ops.mha_fwd(qkv, out, lse, N, S, C, 8), 4.0 * N * S * S * C / 1e3, "TFLOP/s")
        timeit(f"mha bwd N{N} S{S} C{C} (8 heads)", lambda: ops.mha_bwd(qkv, out, dout, lse, delta, dqkv, N, S, C, 8), 8.0 * N * S * S * C / 1e3, "TFLOP/s")

if what in ("metrics", "all"):
    from hdiff_b200 import metrics
    src = torch.randint(0, 256, (64, 720, 1280, 3), device=dev, dtype=torch.uint8)
    timeit("resize_u8 64 x 720x1280 -> 256x256 (CHW)", lambda: metrics.resize_u8(src, 256, 256), (src.numel() + 64 * 3 * 256 * 256) * 1.0, "GB/s")
    img = torch.randint(0, 256, (64, 256, 256, 3), device=dev, dtype=torch.uint8)
    timeit("uiqm_u8 64 x 256x256", lambda: metrics.uiqm_u8(img), img.numel() * 3.0, "GB/s")
    timeit("psnr_u8 64 x 256x256", lambda: metrics.psnr_u8(img, img.flip(0).contiguous()), img.numel() * 2.0, "GB/s")
print("done")
