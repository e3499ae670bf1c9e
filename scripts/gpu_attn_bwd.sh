#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for c in "128-2-1.0" "256-3-1.0" "1024-2-1.0" "4096-1-1.0" "1024-2-4.0"; do
  timeout 120 python -m pytest "tests/test_kernels_gpu.py::test_attention_tcgen05_backward[$c]" -m gpu -q -p no:cacheprovider > gpurun_out/ab_$c.log 2>&1
  echo "attn bwd tc $c: rc=$? $(tail -1 gpurun_out/ab_$c.log)"
  grep -h "AssertionError\|Error\|watchdog" gpurun_out/ab_$c.log | head -3
done
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/all_gpu.log 2>&1
echo "== all gpu tests: $(tail -1 gpurun_out/all_gpu.log)"
grep -h "^FAILED\|^E  .*Error" gpurun_out/all_gpu.log | head -20
python - <<'PY' > gpurun_out/attn_perf.txt 2>&1
import torch, sys
sys.path.insert(0, '.')
import hdiff_b200.ops as hops
ops = hops.get()
dev = torch.device('cuda')
for N, S in ((32, 16384), (32, 1024)):
    C = 128
    qkv = torch.randn(N, S, 3 * C, device=dev).to(torch.bfloat16)
    out = torch.empty(N, S, C, dtype=torch.bfloat16, device=dev)
    dout = torch.randn(N, S, C, device=dev).to(torch.bfloat16)
    dqkv = torch.empty_like(qkv)
    lse = torch.empty(N, S, device=dev)
    for name, fn, fl in (("fwd", lambda: ops.attn_fwd(qkv, out, lse, N, S, C), 4.0), ("bwd", lambda: ops.attn_bwd(qkv, out, dout, lse, None, dqkv, N, S, C), 8.0)):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"attn {name} N={N} S={S}: {ms:.3f} ms  {fl*N*S*S*C/ms/1e9:.1f} algorithmic TFLOP/s")
PY
cat gpurun_out/attn_perf.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err
echo "== bench 256: rc=$?"; cat gpurun_out/bench_256.json; tail -5 gpurun_out/bench_256.err
