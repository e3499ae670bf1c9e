#!/bin/bash
# First GPU pass: CUDA-core kernels, then each tcgen05 conv case in its own process (a trap must not poison the rest).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "not bf16_tc" -p no:cacheprovider > gpurun_out/k_simt.log 2>&1
echo "== simt kernels: $(tail -1 gpurun_out/k_simt.log)"
: > gpurun_out/k_tc.log
for c in 3x3_64_64 3x3_128_128_w32 3x3_concat_128+64_to_64 1x1_concat_shortcut 1x1_qkv_128_384 down_s2d_64 convT_s2d_64 convT_s2d_128 3x3_wide_w256 3x3_ragged_h; do
  timeout 120 python -m pytest "tests/test_kernels_gpu.py::test_conv_forward[bf16_tc-$c]" -m gpu -q -p no:cacheprovider > gpurun_out/tc_$c.log 2>&1
  echo "tc $c: rc=$? $(tail -1 gpurun_out/tc_$c.log)" | tee -a gpurun_out/k_tc.log
done
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -p no:cacheprovider -k "float32 or identity" > gpurun_out/m_fp32.log 2>&1
echo "== model fp32: $(tail -1 gpurun_out/m_fp32.log)"
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -p no:cacheprovider -k "bfloat16 or curve" > gpurun_out/m_bf16.log 2>&1
echo "== model bf16: $(tail -1 gpurun_out/m_bf16.log)"
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
echo "== smoke: $(tail -1 gpurun_out/smoke.log)"
grep -h "FAILED\|Error\|error" gpurun_out/k_simt.log gpurun_out/m_fp32.log gpurun_out/m_bf16.log | head -40
