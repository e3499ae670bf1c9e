#!/bin/bash
# Last evidence of round 2 on one B200: the two 200-step loss-curve tests on the final build, the default bench line,
# the A/B of the weight-gradient side stream, the per-shape convolution breakdown with the binding roofline.
mkdir -p gpurun_out
(timeout 260 python -m pytest tests/test_trajectory_gpu.py -x -q -k "loss_curve" 2>&1 | tail -3) > gpurun_out/r02z_slow.log 2>&1
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err
HDIFF_WGRAD_STREAM=1 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-hybrid --no-gpu-reference --sample-steps 0 > gpurun_out/r02z_bench_wstream.json 2> gpurun_out/r02z_bench_wstream.err
timeout 120 python scripts/conv_breakdown.py > gpurun_out/r02z_conv_breakdown.json 2> gpurun_out/r02z_conv_breakdown.err
tail -3 gpurun_out/r02z_slow.log
python - <<'P'
import json
for f in ("gpurun_out/r02z_bench.json", "gpurun_out/r02z_bench_wstream.json"):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), round(d["ms_per_step"], 2), round(d["e2e"]["value"], 1), (d.get("sampling") or {}).get("value"))
        for k, v in d["kernel_families"].items():
            print("  ", k, round(v["ms_per_step"], 2), round(v["frac"], 3), v.get("frac_of_binding_roofline"))
    except Exception as e:
        print(f, "unreadable:", e)
P
