#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for mb in 0 32 48 64 96; do echo "== L2 MB=$mb"; HDIFF_GN_L2_MB=$mb timeout 300 python scripts/prof_kernels.py gn 5 2>&1 | grep "gn_bwd"; done
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "groupnorm or dropout" 2>&1 | tail -1
HDIFF_GN_L2_MB=48 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_l2.json 2> gpurun_out/bench_l2.err; echo "bench rc=$?"
