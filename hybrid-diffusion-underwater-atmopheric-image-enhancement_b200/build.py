"""In-tree nvcc build of the C-ABI library (sm_100a only).  `python -m hdiff_b200.build`."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhdiff_b200.so")
SOURCES = ["hd_simt.cu", "hd_fused.cu", "hd_gn.cu", "hd_mha.cu", "hd_image.cu", "hd_metrics.cu", "hd_conv_tc.cu", "hd_wgrad_tc.cu", "hd_attn_tc.cu", "hd_attn_wide_tc.cu"]
# lab build (`--lab`): the same sources with -DHDIFF_LAB (timing modes of the conv kernel) + the hardware probes -> libhdiff_b200_lab.so
LAB_SOURCES = ["hd_probe.cu", "hd_probe2.cu"]
LAB_LIB = os.path.join(HERE, "libhdiff_b200_lab.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def _stale(obj, deps):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose: bool = False, force: bool = False, lab: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    sources, lib, suffix, extra = (SOURCES + LAB_SOURCES, LAB_LIB, ".lab.o", ["-DHDIFF_LAB"]) if lab else (SOURCES, LIB, ".o", [])
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "hdiff_b200.h"))
    hdrs = [h for h in hdrs if os.path.exists(h)]
    objs = []
    procs = []
    for s in sources:
        src = os.path.join(CSRC, s)
        if not os.path.exists(src):
            continue
        obj = os.path.join(CSRC, s.replace(".cu", suffix))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
    if force or procs or _stale(lib, objs):
        cmd = [nvcc, "-shared", "-o", lib] + objs + ["-cudart", "static"]
        subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv, lab="--lab" in sys.argv))
