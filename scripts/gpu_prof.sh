#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/all_gpu.log 2>&1
echo "== all gpu tests: $(tail -1 gpurun_out/all_gpu.log)"
grep -h "^FAILED\|^E  .*Error" gpurun_out/all_gpu.log | head -20
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err
echo "== bench 256: rc=$?"; cat gpurun_out/bench_256.json; tail -5 gpurun_out/bench_256.err
timeout 600 python scripts/prof_kernels.py all 3 > gpurun_out/prof_plain.log 2>&1; echo "== prof plain rc=$?"; cat gpurun_out/prof_plain.log
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 python scripts/prof_kernels.py conv 1 > /dev/null 2>&1 && timeout 900 $NCU -k regex:conv_tc_kernel -c 6 -f -o gpurun_out/r01_conv python scripts/prof_kernels.py conv 1 > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 300 python scripts/prof_kernels.py wgrad 1 > /dev/null 2>&1 && timeout 900 $NCU -k regex:wgrad_tc_kernel -c 6 -f -o gpurun_out/r01_wgrad python scripts/prof_kernels.py wgrad 1 > gpurun_out/ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
timeout 300 python scripts/prof_kernels.py attn 1 > /dev/null 2>&1 && timeout 900 $NCU -k "regex:attn_(fwd|bwd)_tc" -c 9 -f -o gpurun_out/r01_attn python scripts/prof_kernels.py attn 1 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 300 python scripts/prof_kernels.py gn 1 > /dev/null 2>&1 && timeout 900 $NCU -k "regex:gn_|colsum" -c 15 -f -o gpurun_out/r01_gn python scripts/prof_kernels.py gn 1 > gpurun_out/ncu_gn.log 2>&1; echo "ncu gn rc=$?"
ls -la gpurun_out/*.ncu-rep
