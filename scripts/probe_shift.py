"""Does a tcgen05 shared-memory operand work when it starts at an arbitrary 128-byte row of a 128B-swizzled TMA box?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hdiff_b200 import _lib
lib = _lib.load_lab()          # probes live in the lab library (python -m hdiff_b200.build --lab)
dev = torch.device("cuda")
torch.manual_seed(0)
x = torch.randn(144, 64, device=dev).to(torch.bfloat16)
w = torch.randn(64, 64, device=dev).to(torch.bfloat16)
for mode in (0, 1):
    res = []
    for shift in range(0, 17):
        out = torch.zeros(128, 64, device=dev)
        _lib.check(lib.hd_probe_shift(x.data_ptr(), w.data_ptr(), out.data_ptr(), shift, mode, torch.cuda.current_stream().cuda_stream), "probe")
        torch.cuda.synchronize()
        ref = x[shift:shift + 128].float() @ w.float().t()
        res.append(float((out - ref).norm() / ref.norm()))
    print("mode", mode, " ".join(f"{r:.1e}" for r in res))
