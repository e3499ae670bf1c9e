"""Queue depth of the tensor pipe (csrc/hd_probe2.cu, probe_queue_kernel): cycles per group of 12 MMAs against the idle time of
the issuing thread after each group."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hdiff_b200 import _lib
lib = _lib.load_lab()          # probes live in the lab library (python -m hdiff_b200.build --lab)
dev = torch.device("cuda")
torch.manual_seed(0)
for N in (64, 128, 256):
    a = torch.randn(128, 192, device=dev).to(torch.bfloat16)
    b = torch.randn(N, 192, device=dev).to(torch.bfloat16)
    row = []
    for gap in (0, 100, 200, 300, 400, 500, 600, 800, 1000, 1500):
        cyc = torch.zeros(2, dtype=torch.int64, device=dev)
        _lib.check(lib.hd_probe_queue(a.data_ptr(), b.data_ptr(), N, 256, gap, 1, 2, cyc.data_ptr(), torch.cuda.current_stream().cuda_stream), "probe_queue")
        torch.cuda.synchronize()
        row.append(f"{gap}:{float(cyc[0]) / 256:6.0f}")
    print(f"N={N:3d} cycles per 12 MMAs by gap  " + "  ".join(row), flush=True)

print("two issuing threads (each its own accumulator, same operands): cycles per 12 MMAs PER THREAD; total work doubles")
for N in (64, 128, 256):
    a = torch.randn(128, 192, device=dev).to(torch.bfloat16)
    b = torch.randn(N, 192, device=dev).to(torch.bfloat16)
    row = []
    for issuers, sw in ((1, 2), (2, 2), (2, 4), (2, 5)):
        for gap in (0, 300):
            cyc = torch.zeros(2, dtype=torch.int64, device=dev)
            _lib.check(lib.hd_probe_queue(a.data_ptr(), b.data_ptr(), N, 256, gap, issuers, sw, cyc.data_ptr(), torch.cuda.current_stream().cuda_stream), "probe_queue")
            torch.cuda.synchronize()
            row.append(f"{issuers} thr (warp {sw}) gap {gap}: {float(cyc.max()) / 256:6.0f}")
    print(f"N={N:3d}  " + "  ".join(row), flush=True)
