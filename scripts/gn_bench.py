"""GroupNorm kernel family at the cfg2 layer shapes (batch 32), each kernel timed alone with CUDA events, against the
compulsory-byte roofline (every operand once).  Explores the (CTAs per SM, pixels in flight) variants of
csrc/hd_gn.cu through their environment switches, one child process per setting (the switches are read once).

    python scripts/gn_bench.py                 # the default build + HDIFF_GN_V1=1 (first generation)
    python scripts/gn_bench.py sweep           # every variant
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = ((32, 65536, 64, 0), (32, 65536, 128, 64), (32, 16384, 128, 0), (32, 16384, 128, 128), (32, 4096, 128, 0))


def one():
    import torch
    import hdiff_b200.ops as hops
    ops = hops.get()
    dev = torch.device("cuda")
    bf = torch.bfloat16
    reps = 5
    peak = 6535.4
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    rows = {}

    def timeit(name, fn, nbytes):
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
        rows[name] = (ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / peak)

    for (N, HW, C0, C1) in SHAPES:
        C = C0 + C1
        tag = f"N{N} HW{HW} C{C0}+{C1}"
        x0 = torch.randn(N, HW, 1, C0, device=dev).to(bf)
        x1 = torch.randn(N, HW, 1, C1, device=dev).to(bf) if C1 else None
        gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
        sums = torch.empty(N, 32, 2, dtype=torch.float64, device=dev)
        out = torch.empty(N, HW, 1, C, dtype=bf, device=dev)
        numel = N * HW * C
        timeit(f"stats {tag}", lambda: ops.gn_stats(x0, x1, N, HW, 32, sums), numel * 2.0)
        for p in (0.0, 0.1):
            timeit(f"apply p{p} {tag}", lambda: ops.gn_apply(x0, x1, N, HW, 32, sums, gamma, beta, 1e-5, 1, p, 123, out), numel * 4.0)
        dy = torch.randn(N, HW, 1, C, device=dev).to(bf)
        gs = torch.empty_like(sums)
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        dx0 = torch.empty_like(x0)
        dx1 = None if x1 is None else torch.empty_like(x1)
        lib, P, S = ops.lib, hops._p, hops._stream
        for p in (0.0, 0.1):
            a = (1, P(x0), C0, P(x1), C1, N, HW, 32, P(sums), P(gamma), P(beta), 1e-5, 1, p, 123, P(dy))
            timeit(f"bwd_reduce p{p} {tag}", lambda: lib.hd_gn_bwd_reduce(*a, P(gs), P(dg), P(db), None, S()), numel * 4.0)
            timeit(f"bwd_apply p{p} {tag}", lambda: lib.hd_gn_bwd_apply(*a, P(gs), None, None, None, P(dx0), P(dx1), None, None, 0, C, 0, S()), numel * 6.0)
            timeit(f"bwd (both) p{p} {tag}", lambda: ops.gn_bwd(x0, x1, N, HW, 32, sums, gamma, beta, 1e-5, 1, p, 123, dy, gs, dg, db, None, None, None, dx0, dx1), numel * 6.0)
        add = torch.randn(N, HW, 1, C, device=dev).to(bf)
        acc0 = torch.randn(N, HW, 1, C0, device=dev).to(bf)
        cs = torch.zeros(C0, device=dev)
        timeit(f"bwd (both) +add+acc0+cs {tag}", lambda: ops.gn_bwd(x0, x1, N, HW, 32, sums, gamma, beta, 1e-5, 1, 0.0, 0, dy, gs, dg, db, add, acc0, None, dx0, dx1,
                                                                     cs_total=cs, cs_n=C0), numel * 8.0 + N * HW * C0 * 2.0)
        del x0, x1, out, dy, dx0, dx1, add, acc0
    print(json.dumps(rows))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        return one()
    settings = [("default", {}), ("gen1", {"HDIFF_GN_V1": "1"})]
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":
        for v in ("34", "32", "42", "52", "62", "24"):
            settings.append((f"FWD/RED/APP={v}", {"HDIFF_GN_FWD": v, "HDIFF_GN_RED": v, "HDIFF_GN_APP": v}))
    table = {}
    for name, env in settings:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=dict(os.environ, **env), capture_output=True, text=True)
        if r.returncode != 0:
            print(name, "FAILED", r.stderr[-2000:])
            continue
        table[name] = json.loads(r.stdout.strip().splitlines()[-1])
    names = list(table)
    keys = list(next(iter(table.values())))
    print(f"{'kernel @ shape':58s}" + "".join(f"{n:>22s}" for n in names))
    for k in keys:
        print(f"{k:58s}" + "".join(f"{table[n][k][0]:9.3f} ms {table[n][k][2]:6.2f}   " for n in names))
    tot = {n: sum(v[0] for v in table[n].values()) for n in names}
    print(f"{'sum':58s}" + "".join(f"{tot[n]:9.3f} ms          " for n in names))


if __name__ == "__main__":
    main()
