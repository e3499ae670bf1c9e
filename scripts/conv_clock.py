"""SM clock DURING a convolution kernel (HDIFF_CONV_DBG=4: CTA 0 stamps clock64 and globaltimer at its start and end),
against the clock NVML reports, for a back-to-back stream of launches.  python scripts/conv_clock.py"""
import ctypes, os, sys, threading, time
os.environ["HDIFF_CONV_DBG"] = os.environ.get("HDIFF_CONV_DBG", "4")
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hdiff_b200.ops as hops
from hdiff_b200 import _lib
import pynvml
_lib._lib = _lib.load_lab()          # the timing modes exist only in the lab build (python -m hdiff_b200.build --lab)
ops = hops.get()
dev = torch.device("cuda")
bf = torch.bfloat16
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
for (N, H, W, C, Cout) in ((32, 256, 256, 64, 64), (32, 128, 128, 128, 128)):
    x = torch.randn(N, H, W, C, device=dev).to(bf)
    w = (torch.randn(Cout * 9 * C, device=dev) / (9 * C) ** 0.5).to(bf)
    bias = torch.randn(Cout, device=dev)
    out = torch.empty(N, H, W, Cout, device=dev, dtype=bf)
    for warm, n_launch in ((0, 3), (1, 400)):
        clocks = []
        stop = False
        def sample():
            while not stop:
                clocks.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)); time.sleep(0.002)
        th = threading.Thread(target=sample); th.start()
        torch.cuda.synchronize(); t0 = time.time()
        for _ in range(n_launch):
            ops.conv(x, None, 1, w, bias, None, None, out, 1, N, H, W, 3)
        torch.cuda.synchronize(); dt = time.time() - t0
        stop = True; th.join()
        buf = (ctypes.c_longlong * 4)()
        assert ops.lib.hd_conv_dbg_read(ctypes.cast(buf, ctypes.c_void_p)) == 0
        cyc, ns = buf[2] - buf[0], buf[3] - buf[1]
        print(f"conv3x3 {C}->{Cout} {H}x{W}: {n_launch} launches {dt / n_launch * 1e3:.3f} ms each; last launch CTA0: {cyc} cycles in {ns} ns = "
              f"{cyc / ns * 1e3:.0f} MHz in-kernel; NVML SM clock median {sorted(clocks)[len(clocks) // 2]} MHz (n={len(clocks)})", flush=True)
