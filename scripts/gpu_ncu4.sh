#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
NCU="ncu --set full --clock-control none"
timeout 300 python scripts/prof_kernels.py conv 1 > /dev/null 2>&1 && timeout 900 $NCU -k regex:conv_tc_kernel -s 2 -c 1 -f -o gpurun_out/r01_conv5 python scripts/prof_kernels.py conv 1 > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
