#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for g in 4 3 2 1; do echo "== G=$g"; HDIFF_WGRAD_G=$g timeout 300 python scripts/prof_kernels.py wgrad 5 2>&1 | grep wgrad; done
