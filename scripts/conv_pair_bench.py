"""3x3 convolutions with 128 / 256 output channels per tile: CTA-pair mode (cta_group::2, half a weight tile per CTA) against the
one-CTA kernel.  One child process per HDIFF_CONV_PAIR setting (the switch is read once); every case is checked against the
operator test double before it is timed.

    python scripts/conv_pair_bench.py
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = ((32, 128, 128, 128, 0, 128, 0), (32, 128, 128, 128, 0, 128, 1), (32, 256, 256, 128, 0, 128, 0), (32, 256, 256, 64, 0, 128, 0),
         (32, 128, 128, 256, 0, 128, 1), (32, 128, 128, 128, 0, 256, 0), (16, 128, 128, 256, 0, 256, 1), (16, 256, 256, 128, 0, 128, 1),
         (16, 128, 128, 256, 128, 256, 0), (32, 64, 64, 128, 0, 128, 0), (16, 64, 64, 256, 0, 256, 1), (3, 128, 128, 128, 0, 128, 1),
         (32, 256, 256, 64, 0, 64, 0), (32, 256, 256, 64, 0, 64, 1), (32, 256, 256, 128, 64, 64, 1), (32, 128, 128, 64, 0, 64, 0),
         (32, 256, 256, 128, 0, 64, 1), (32, 256, 256, 192, 0, 64, 1), (32, 128, 128, 128, 0, 64, 0))
# CONV_BENCH_KSIZE=1: the 1x1 layers of the cfg2 step instead (shortcuts, q/k/v, proj and their gradients), e.g.
#   CONV_BENCH_KSIZE=1 python scripts/conv_pair_bench.py 1,HDIFF_CONV_STAGE_BUFS=2 1,HDIFF_CONV_STAGE_BUFS=4
KS = int(os.environ.get("CONV_BENCH_KSIZE", "3"))
if KS == 1:
    CASES = ((32, 256, 256, 64, 0, 128, 0), (32, 256, 256, 64, 0, 192, 0), (32, 128, 128, 128, 0, 384, 0), (32, 128, 128, 128, 0, 256, 0),
             (32, 128, 128, 128, 0, 128, 0), (32, 128, 128, 128, 0, 128, 1), (32, 128, 128, 384, 0, 128, 0), (32, 256, 256, 128, 64, 64, 0),
             (32, 256, 256, 64, 64, 64, 0), (32, 64, 64, 128, 0, 256, 0), (16, 256, 256, 64, 0, 128, 0), (3, 128, 128, 128, 0, 384, 0))
KK = KS * KS


def one():
    import torch
    import hdiff_b200.ops as hops
    from tests.emu_backend import EmuOps
    ops, emu = hops.get(), EmuOps()
    dev, bf = torch.device("cuda"), torch.bfloat16
    rows = {}
    for (N, H, W, C0, C1, Cout, extra) in CASES:
        torch.manual_seed(N + H + C0 + Cout)
        x0 = torch.randn(N, H, W, C0, device=dev).to(bf)
        x1 = torch.randn(N, H, W, C1, device=dev).to(bf) if C1 else None
        Cin = C0 + C1
        w = (torch.randn(Cout, KK, Cin, device=dev) / (KK * Cin) ** 0.5).to(bf)
        bias = torch.randn(Cout, device=dev)
        emb = torch.randn(N, Cout, device=dev) if extra else None
        res = torch.randn(N, H, W, Cout, device=dev).to(bf) if extra else None
        out = torch.full((N, H, W, Cout), float("nan"), device=dev, dtype=bf)
        ops.conv(x0, x1, 1, w, bias, emb, res, out, 1, N, H, W, KS)
        torch.cuda.synchronize()
        nchk = min(N, 3)                                   # the first images and the last one
        idx = list(range(nchk - 1)) + [N - 1]
        ref = torch.empty(len(idx), H, W, Cout, device=dev)
        emu.conv(x0[idx].float(), None if x1 is None else x1[idx].float(), 1, w.float(), bias, None if emb is None else emb[idx],
                 None if res is None else res[idx].float(), ref, 1, len(idx), H, W, KS)
        err = float((out[idx].float() - ref).norm() / ref.norm())
        finite = bool(torch.isfinite(out.float()).all())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            ops.conv(x0, x1, 1, w, bias, emb, res, out, 1, N, H, W, KS)
        e0.record()
        for _ in range(5):
            ops.conv(x0, x1, 1, w, bias, emb, res, out, 1, N, H, W, KS)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        rows[f"N{N} {H}x{W} {C0}+{C1}->{Cout}{' +res+emb' if extra else ''}"] = (ms, 2.0 * N * H * W * Cout * KK * Cin / ms / 1e9, err, finite)
        del x0, x1, out, ref
    print(json.dumps(rows))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        return one()
    table = {}
    for v in sys.argv[1:] or ("0", "1", "2"):
        env = dict(os.environ)
        for kv in v.split(","):                      # "2" or "2,HDIFF_CONV_STAGE=0,HDIFF_CONV_BUDGET=222"
            k, _, val = kv.partition("=")
            env.update({k: val} if val else {"HDIFF_CONV_PAIR": k})
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, capture_output=True, text=True, timeout=300)
        except subprocess.TimeoutExpired:
            print(f"PAIR={v}: TIMEOUT")
            continue
        if r.returncode != 0:
            print(f"PAIR={v}: FAILED", r.stdout[-1500:], r.stderr[-2500:])
            continue
        table[v] = json.loads(r.stdout.strip().splitlines()[-1])
    if not table:
        return
    keys = list(next(iter(table.values())))
    print(f"{'case':44s}" + "".join(f"{'PAIR=' + v[:28]:>34s}" for v in table))
    for k in keys:
        print(f"{k:44s}" + "".join(f"{table[v][k][0]:9.3f} ms {table[v][k][1]:7.0f} TF/s err {table[v][k][2]:.1e}{'' if table[v][k][3] else ' NaN!'}" for v in table))


if __name__ == "__main__":
    main()
