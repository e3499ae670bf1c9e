#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/all_gpu.log 2>&1
echo "== all gpu tests: $(tail -1 gpurun_out/all_gpu.log)"
grep -h "^FAILED\|^E  .*Error" gpurun_out/all_gpu.log | head -20
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err
echo "== bench 256: rc=$?"; tail -5 gpurun_out/bench_256.err
timeout 1500 python scripts/parity_trajectory.py > gpurun_out/parity_trajectory.json 2> gpurun_out/parity.err
echo "== parity rc=$?"; cat gpurun_out/parity_trajectory.json | cut -c1-1500; tail -5 gpurun_out/parity.err
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "== smoke: $(tail -1 gpurun_out/smoke.log)"
