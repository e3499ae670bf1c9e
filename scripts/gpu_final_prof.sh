#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sample-steps 0 > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sample-steps 0 > gpurun_out/ncu.log 2>&1
echo "== launch list rc=$?"
NCU="ncu --set full --clock-control none"
timeout 300 python scripts/prof_kernels.py conv 1 > /dev/null 2>&1 && timeout 900 $NCU -k regex:conv_tc_kernel -s 2 -c 4 -f -o gpurun_out/r01_conv_final python scripts/prof_kernels.py conv 1 > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 300 python scripts/prof_kernels.py wgrad 1 > /dev/null 2>&1 && timeout 900 $NCU -k regex:wgrad_tc_kernel -s 2 -c 4 -f -o gpurun_out/r01_wgrad_final python scripts/prof_kernels.py wgrad 1 > gpurun_out/ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
timeout 300 python scripts/prof_kernels.py attn32 1 > /dev/null 2>&1 && timeout 900 $NCU --import-source on -k "regex:attn_(fwd2|bwd)_tc" -s 2 -c 3 -f -o gpurun_out/r01_attn_final python scripts/prof_kernels.py attn32 1 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
du -sh gpurun_out
