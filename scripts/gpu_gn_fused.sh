#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3
for f in 0 1; do echo "== fused=$f"; HDIFF_GN_FUSED=$f timeout 300 python scripts/prof_kernels.py gn 5 2>&1 | grep "gn_bwd"; done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_256.err
