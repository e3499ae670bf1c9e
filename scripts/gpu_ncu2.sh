#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
NCU="ncu --set full --clock-control none"
timeout 300 python scripts/prof_kernels.py gn 1 > /dev/null 2>&1 && timeout 900 $NCU --import-source on -k "regex:gn_" -s 6 -c 6 -f -o gpurun_out/r01_gn2 python scripts/prof_kernels.py gn 1 > gpurun_out/ncu_gn.log 2>&1; echo "ncu gn rc=$?"
timeout 300 python scripts/prof_kernels.py attn32 1 > /dev/null 2>&1 && timeout 900 $NCU -k "regex:attn_(fwd|bwd)_tc" -s 2 -c 3 -f -o gpurun_out/r01_attn32 python scripts/prof_kernels.py attn32 1 > gpurun_out/ncu_attn32.log 2>&1; echo "ncu attn32 rc=$?"
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
