#!/bin/bash
# wgrad tcgen05 cases one process each (a trap must not poison the rest), then the model tests and a plumbing bench.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
: > gpurun_out/wg_tc.log
for c in 3x3_64_64 3x3_128_128_w32 3x3_concat_128+64_to_64 1x1_concat_shortcut 1x1_qkv_128_384 down_s2d_64 convT_s2d_64 convT_s2d_128 3x3_wide_w256 3x3_ragged_h; do
  timeout 120 python -m pytest "tests/test_kernels_gpu.py::test_conv_wgrad[bf16_tc-$c]" -m gpu -q -p no:cacheprovider > gpurun_out/wg_$c.log 2>&1
  echo "wgrad tc $c: rc=$? $(tail -1 gpurun_out/wg_$c.log)" | tee -a gpurun_out/wg_tc.log
  grep -h "AssertionError\|Error" gpurun_out/wg_$c.log | head -3
done
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "not wgrad" > gpurun_out/k_all.log 2>&1
echo "== kernels (non-wgrad): $(tail -1 gpurun_out/k_all.log)"
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/m_all.log 2>&1
echo "== model: $(tail -1 gpurun_out/m_all.log)"
grep -h "^FAILED\|^E  " gpurun_out/k_all.log gpurun_out/m_all.log | head -30
timeout 600 python bench.py --res 64 --batch 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r64.json 2> gpurun_out/bench_r64.err
echo "== bench r64: rc=$?"; cat gpurun_out/bench_r64.json; tail -5 gpurun_out/bench_r64.err
