#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for c in "256-3-1.0" "1024-2-1.0" "4096-1-1.0" "1024-2-6.0" "2048-1-12.0"; do
  timeout 120 python -m pytest "tests/test_kernels_gpu.py::test_attention_tcgen05_forward[$c]" -m gpu -q -p no:cacheprovider > gpurun_out/at_$c.log 2>&1
  echo "attn fwd2 $c: rc=$? $(tail -1 gpurun_out/at_$c.log)"; grep -h "AssertionError\|Error\|watchdog" gpurun_out/at_$c.log | head -3
done
echo "== v2"; timeout 300 python scripts/prof_kernels.py attn32 5 2>&1 | grep "attn fwd"; timeout 300 python scripts/prof_kernels.py attn 5 2>&1 | grep "attn fwd"
echo "== v1"; HDIFF_ATTN_FWD_V1=1 timeout 300 python scripts/prof_kernels.py attn32 5 2>&1 | grep "attn fwd"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -2
