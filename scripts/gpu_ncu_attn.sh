#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 python scripts/prof_kernels.py attn 1 > /dev/null 2>&1 && timeout 900 $NCU -k "regex:attn_(fwd|bwd)_tc" -s 2 -c 3 -f -o gpurun_out/r01_attn python scripts/prof_kernels.py attn 1 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
