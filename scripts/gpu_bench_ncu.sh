#!/bin/bash
# full GPU test suite, the bench line, then the ncu launch list of a short bench run
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/all_gpu.log 2>&1
echo "== all gpu tests: $(tail -1 gpurun_out/all_gpu.log)"
grep -h "^FAILED\|^E  .*Error" gpurun_out/all_gpu.log | head -20
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err
echo "== bench 256: rc=$?"; cat gpurun_out/bench_256.json; tail -5 gpurun_out/bench_256.err
if [ "$1" = "ncu" ]; then
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "== ncu rc=$?"; tail -3 gpurun_out/ncu.log; wc -l gpurun_out/launches.csv
fi
