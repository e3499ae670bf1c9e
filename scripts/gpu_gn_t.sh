#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -2
timeout 300 python scripts/prof_kernels.py gn 5 2>&1 | grep "gn_"
