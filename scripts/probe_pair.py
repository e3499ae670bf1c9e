"""CTA-pair MMA probe (csrc/hd_probe2.cu): correctness of tcgen05.mma.cta_group::2 against torch, and cycles per MMA of the
one-CTA and the pair form for the operand shapes of the convolution kernels."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hdiff_b200 import _lib
lib = _lib.load_lab()          # probes live in the lab library (python -m hdiff_b200.build --lab)
dev = torch.device("cuda")
torch.manual_seed(0)
for N in (64, 128, 256):
    for K in (64, 192):
        a = torch.randn(256, K, device=dev).to(torch.bfloat16)
        b = torch.randn(N, K, device=dev).to(torch.bfloat16)
        ref = a.float() @ b.float().t()
        line = [f"N={N:3d} K={K:3d}"]
        for mode in (1, 2):
            for reps in (1, 64):
                out = torch.full((256, N), float("nan"), device=dev)
                cyc = torch.zeros(4, dtype=torch.int64, device=dev)
                _lib.check(lib.hd_probe_pair(a.data_ptr(), b.data_ptr(), out.data_ptr(), N, K, mode, reps, cyc.data_ptr(), 0, 0, 0, 1, 0,
                                             torch.cuda.current_stream().cuda_stream), "probe_pair")
                torch.cuda.synchronize()
                err = float((out - reps * ref).norm() / (reps * ref).norm())
                n_mma = reps * (K // 16)
                line.append(f"mode{mode} reps{reps}: err {err:.1e} cyc/MMA {float(cyc[:2].max()) / n_mma:7.1f}")
        print(" | ".join(line), flush=True)

print("cycles per MMA with the A operand starting `shift` rows into the box (K=192, reps=64):")
for N in (64, 128):
    K = 192
    a = torch.randn(256, K, device=dev).to(torch.bfloat16)
    b = torch.randn(N, K, device=dev).to(torch.bfloat16)
    for mode in (1, 2):
        row = []
        for shift in (0, 1, 2, 4, 8):
            out = torch.empty(256, N, device=dev)
            cyc = torch.zeros(4, dtype=torch.int64, device=dev)
            _lib.check(lib.hd_probe_pair(a.data_ptr(), b.data_ptr(), out.data_ptr(), N, K, mode, 64, cyc.data_ptr(), shift, 0, 0, 1, 0,
                                         torch.cuda.current_stream().cuda_stream), "probe_pair")
            torch.cuda.synchronize()
            row.append(f"shift {shift}: {float(cyc[:2].max()) / (64 * K // 16):6.1f}")
        print(f"N={N:3d} mode{mode} " + "  ".join(row), flush=True)

print("MMA cycles with a concurrent TMA fill of shared memory (mode 1, K=192, reps=64 -> 768 MMAs):")
for N in (64, 128):
    K = 192
    a = torch.randn(256, K, device=dev).to(torch.bfloat16)
    b = torch.randn(N, K, device=dev).to(torch.bfloat16)
    for fill in (0, 64, 256, 1024, 4096):
        out = torch.empty(256, N, device=dev)
        cyc = torch.zeros(4, dtype=torch.int64, device=dev)
        _lib.check(lib.hd_probe_pair(a.data_ptr(), b.data_ptr(), out.data_ptr(), N, K, 1, 64, cyc.data_ptr(), 0, fill, 0, 1, 0,
                                     torch.cuda.current_stream().cuda_stream), "probe_pair")
        torch.cuda.synchronize()
        mma, fl = float(cyc[:2].max()), float(cyc[2:].max())
        print(f"N={N:3d} fill {fill:5d} x 16 KB: MMA {mma / 768:6.1f} cyc/MMA ({mma:9.0f} total)   fill {fl:9.0f} cycles"
              + (f" = {fill * 16384 / fl:5.1f} B/clk" if fill else ""), flush=True)

print("cycles per MMA with a tcgen05.commit every `cper` MMAs (K=192, reps=64):")
for N in (64, 128):
    K = 192
    a = torch.randn(256, K, device=dev).to(torch.bfloat16)
    b = torch.randn(N, K, device=dev).to(torch.bfloat16)
    for mode in (1, 2):
        row = []
        for cper in (0, 48, 24, 12, 8, 4, 1):
            out = torch.empty(256, N, device=dev)
            cyc = torch.zeros(4, dtype=torch.int64, device=dev)
            _lib.check(lib.hd_probe_pair(a.data_ptr(), b.data_ptr(), out.data_ptr(), N, K, mode, 64, cyc.data_ptr(), 0, 0, cper, 1, 0,
                                         torch.cuda.current_stream().cuda_stream), "probe_pair")
            torch.cuda.synchronize()
            row.append(f"cper {cper}: {float(cyc[:2].max()) / (64 * K // 16):6.1f}")
        print(f"N={N:3d} mode{mode} " + "  ".join(row), flush=True)

print("cycles per MMA when `nclusters` pairs run the same MMA stream at once (K=192, reps=512):")
for N in (64, 128, 256):
    K = 192 if N < 256 else 128
    a = torch.randn(256, K, device=dev).to(torch.bfloat16)
    b = torch.randn(N, K, device=dev).to(torch.bfloat16)
    for mode in (1, 2):
        row = []
        for ncl in (1, 8, 37, 74):
            out = torch.empty(256, N, device=dev)
            cyc = torch.zeros(4, dtype=torch.int64, device=dev)
            for _ in range(3):
                _lib.check(lib.hd_probe_pair(a.data_ptr(), b.data_ptr(), out.data_ptr(), N, K, mode, 512, cyc.data_ptr(), 0, 0, 0, ncl, 0,
                                             torch.cuda.current_stream().cuda_stream), "probe_pair")
            torch.cuda.synchronize()
            row.append(f"{ncl:2d} pairs: {float(cyc[:2].max()) / (512 * K // 16):6.1f}")
        print(f"N={N:3d} mode{mode} " + "  ".join(row), flush=True)

print("cycles per MMA with the full/empty handshake of a ring of `ring` stages, 12 MMAs per stage, no data moved (K=192, reps=64):")
for N in (64, 128):
    K = 192
    a = torch.randn(256, K, device=dev).to(torch.bfloat16)
    b = torch.randn(N, K, device=dev).to(torch.bfloat16)
    row = []
    for ring in (0, 1, 2, 3, 4, 5, 8):
        out = torch.empty(256, N, device=dev)
        cyc = torch.zeros(4, dtype=torch.int64, device=dev)
        _lib.check(lib.hd_probe_pair(a.data_ptr(), b.data_ptr(), out.data_ptr(), N, K, 1, 64, cyc.data_ptr(), 0, 0, 0, 1, ring,
                                     torch.cuda.current_stream().cuda_stream), "probe_pair")
        torch.cuda.synchronize()
        row.append(f"ring {ring}: {float(cyc[:2].max()) / (64 * K // 16):6.1f}")
    print(f"N={N:3d} " + "  ".join(row), flush=True)

print("cycles per 12 MMAs when the issuing thread idles `gap` cycles after every 12 (K=192, reps=64): queue depth of the tensor pipe")
for N in (64, 128):
    K = 192
    a = torch.randn(256, K, device=dev).to(torch.bfloat16)
    b = torch.randn(N, K, device=dev).to(torch.bfloat16)
    row = []
    for gap in (0, 50, 100, 150, 200, 300, 400, 600):
        out = torch.empty(256, N, device=dev)
        cyc = torch.zeros(4, dtype=torch.int64, device=dev)
        _lib.check(lib.hd_probe_pair(a.data_ptr(), b.data_ptr(), out.data_ptr(), N, K, 1, 64, cyc.data_ptr(), 0, 0, 0, 1, -gap if gap else 0,
                                     torch.cuda.current_stream().cuda_stream), "probe_pair")
        torch.cuda.synchronize()
        row.append(f"gap {gap}: {float(cyc[:2].max()) / (64 * K // 16) * 12:6.0f}")
    print(f"N={N:3d} " + "  ".join(row), flush=True)
