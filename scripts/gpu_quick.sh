#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/all_gpu.log 2>&1
echo "== all gpu tests: $(tail -1 gpurun_out/all_gpu.log)"
grep -h "^FAILED\|^E  .*Error" gpurun_out/all_gpu.log | head -20
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err
echo "== bench 256: rc=$?"; cat gpurun_out/bench_256.json; tail -5 gpurun_out/bench_256.err
timeout 600 python scripts/prof_kernels.py ${1:-all} 3 > gpurun_out/prof_plain.log 2>&1; echo "== prof plain rc=$?"; cat gpurun_out/prof_plain.log
