#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sample-steps 0 > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sample-steps 0 > gpurun_out/ncu.log 2>&1
echo "== ncu rc=$?"; wc -l gpurun_out/launches.csv
