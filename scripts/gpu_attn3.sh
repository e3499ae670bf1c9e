#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for c in "128-2-1.0" "256-3-1.0" "1024-2-1.0" "4096-1-1.0" "1024-2-4.0"; do
  timeout 120 python -m pytest "tests/test_kernels_gpu.py::test_attention_tcgen05_backward[$c]" -m gpu -q -p no:cacheprovider > gpurun_out/ab_$c.log 2>&1
  echo "attn bwd $c: rc=$? $(tail -1 gpurun_out/ab_$c.log)"; grep -h "AssertionError\|Error\|watchdog" gpurun_out/ab_$c.log | head -3
done
timeout 300 python scripts/prof_kernels.py attn32 5 2>&1 | grep "attn"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err; echo "bench rc=$?"
