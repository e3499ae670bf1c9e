#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for c in "128-2-1.0" "256-3-1.0" "1024-2-1.0" "4096-1-1.0" "1024-2-6.0" "2048-1-12.0"; do
  timeout 120 python -m pytest "tests/test_kernels_gpu.py::test_attention_tcgen05_forward[$c]" -m gpu -q -p no:cacheprovider > gpurun_out/at_$c.log 2>&1
  echo "attn fwd tc $c: rc=$? $(tail -1 gpurun_out/at_$c.log)"
  grep -h "AssertionError\|Error\|watchdog" gpurun_out/at_$c.log | head -3
done
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "attention" > gpurun_out/k_attn.log 2>&1
echo "== attention tests: $(tail -1 gpurun_out/k_attn.log)"
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/m_all.log 2>&1
echo "== model: $(tail -1 gpurun_out/m_all.log)"
grep -h "^FAILED\|^E  .*Error" gpurun_out/k_attn.log gpurun_out/m_all.log | head -20
python - <<'PY' > gpurun_out/attn_perf.txt 2>&1
import torch, sys
sys.path.insert(0, '.')
import hdiff_b200.ops as hops
ops = hops.get()
dev = torch.device('cuda')
for N, S in ((32, 16384), (32, 1024), (8, 16384)):
    C = 128
    qkv = torch.randn(N, S, 3 * C, device=dev).to(torch.bfloat16)
    out = torch.empty(N, S, C, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(N, S, device=dev)
    for _ in range(2): ops.attn_fwd(qkv, out, lse, N, S, C)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.attn_fwd(qkv, out, lse, N, S, C)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"attn fwd N={N} S={S}: {ms:.3f} ms  {4.0*N*S*S*C/ms/1e9:.1f} TFLOP/s")
PY
cat gpurun_out/attn_perf.txt
