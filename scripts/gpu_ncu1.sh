#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 python scripts/prof_kernels.py conv 1 > /dev/null 2>&1 && timeout 900 $NCU -k regex:conv_tc_kernel -s 2 -c 4 -f -o gpurun_out/r01_conv python scripts/prof_kernels.py conv 1 > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 300 python scripts/prof_kernels.py wgrad 1 > /dev/null 2>&1 && timeout 900 $NCU -k regex:wgrad_tc_kernel -s 2 -c 4 -f -o gpurun_out/r01_wgrad python scripts/prof_kernels.py wgrad 1 > gpurun_out/ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
timeout 300 python scripts/prof_kernels.py gn 1 > /dev/null 2>&1 && timeout 900 ncu --set full --clock-control none -k "regex:gn_|colsum" -s 2 -c 13 -f -o gpurun_out/r01_gn python scripts/prof_kernels.py gn 1 > gpurun_out/ncu_gn.log 2>&1; echo "ncu gn rc=$?"
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
