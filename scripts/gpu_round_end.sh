#!/bin/bash
# round-end evidence of the last build: tests, bench, launch list of the same command, trajectory parity, smoke
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/all_gpu.log 2>&1
echo "== all gpu tests: $(tail -1 gpurun_out/all_gpu.log)"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err
echo "== bench 256: rc=$?"; cut -c1-260 gpurun_out/bench_256.json
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sample-steps 0 > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sample-steps 0 > gpurun_out/ncu.log 2>&1
echo "== launch list rc=$?"
timeout 1200 python scripts/parity_trajectory.py > gpurun_out/parity_trajectory.json 2> gpurun_out/parity.err
echo "== parity rc=$?"; cut -c1-300 gpurun_out/parity_trajectory.json
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "== smoke: $(tail -1 gpurun_out/smoke.log)"
