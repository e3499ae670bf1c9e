"""bench.py — DDPM UNet training throughput (images/s) at 256x256 on N B200s, the metric BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 our arm (hdiff_b200 CUDA path)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  reference arm: the fp32 PyTorch path on host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

A step is one pass of the hot path over one batch: GaussianDiffusionTrainer.forward(x_0) -> `.sum()/1000.` ->
backward -> clip_grad_norm_(1.0) -> AdamW (diffusion/Train.py:49-55 of the reference).  Workload at every N:
BASELINE.json configs[1] (UNet ch=64 ch_mult=[1,2,2,2] attn=[1] num_res_blocks=2 T=1000 dropout=0.1, 256x256 RGB,
batch 32 per GPU, bf16 compute with fp32 master weights, synthetic data, random-init weights).

One JSON line on stdout (rank 0).  `value`: inputs resident in HBM; `e2e`: the same step driven with pinned HOST
batches (host->device copy of x_0 and device->host read of the loss inside the timed region).
`oracle/` is imported only by the cpu_baseline leg and by `--impl reference`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG2 = dict(T=1000, ch=64, ch_mult=[1, 2, 2, 2], attn=[1], num_res_blocks=2, dropout=0.1)
BETA_1, BETA_T = 1e-4, 0.02
METRIC = "ddpm_train_images_per_sec_256"
UNIT = "images/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        hi = sm[len(sm) // 2:] if sm else []           # samples under load = upper half
        return {"sm_mhz": (hi[len(hi) // 2] if hi else None), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (fp32 PyTorch restatement of the reference path) on host cores
# ---------------------------------------------------------------------------------------------------
def cpu_reference_steps(steps, warmup, res, batch, threads=None):
    """Times `steps` training steps of the reference fp32 path on the host.  Returns (images/s, seconds per step)."""
    import torch
    from oracle import ref_torch as R                    # CPU baseline only: the thing timed here is the reference path
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = R.UNet(**CFG2)
    tr = R.GaussianDiffusionTrainer(net, BETA_1, BETA_T, CFG2["T"])
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=1e-4)
    x = torch.rand(batch, 3, res, res) * 2 - 1
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        R.train_step(tr, opt, x)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return batch / med, med, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res, batch = args.res, args.ref_batch
    v, sec, threads = cpu_reference_steps(args.steps, max(1, min(args.warmup, 1)), res, batch)
    sample = (f"{args.steps} timed training steps (median) of the fp32 PyTorch reference path on the host, batch {batch} at {res}x{res} "
              f"(the reference materialises [S,S] attention scores: ~1 GB per image per block at S=16384), reported per image")
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"cfg2 DDPM UNet ch=64 [1,2,2,2] attn=[1] nrb=2 T=1000 dropout=0.1 train step, {res}x{res}",
                      "global_batch": batch, "resolution": res},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from hdiff_b200.diffusion.Model import UNet
    from hdiff_b200.diffusion.Diffusion import GaussianDiffusionTrainer
    from hdiff_b200.optim import FlatAdamW
    from hdiff_b200 import parallel
    import hdiff_b200.ops as hops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # keep stdout to the one JSON line: anything libraries print meanwhile (NCCL's version banner, warnings) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, res = args.batch, args.res
    torch.manual_seed(0)                                   # identical replicas
    net = UNet(**CFG2).to(dev)
    net.train()
    if world > 1:
        parallel.enable_data_parallel(net)
    trainer = GaussianDiffusionTrainer(net, BETA_1, BETA_T, CFG2["T"]).to(dev)
    opt = FlatAdamW(net, lr=1e-4, weight_decay=1e-4, max_grad_norm=1.0)
    ops = hops.get()
    torch.manual_seed(1000 + rank)                         # per-rank data and RNG stream
    n_pool = 6                                             # 6 x 25 MB of distinct batches > the 126 MB L2
    host_pool = [(torch.rand(B, 3, res, res) * 2 - 1).pin_memory() for _ in range(n_pool)]
    dev_pool = [h.to(dev) for h in host_pool]

    def step_resident(i):
        opt.zero_grad()
        loss = trainer(dev_pool[i % n_pool]).sum() / 1000.
        loss.backward()
        opt.step()
        return loss

    def step_e2e(i):
        x = host_pool[i % n_pool].to(dev, non_blocking=True)
        opt.zero_grad()
        loss = trainer(x).sum() / 1000.
        loss.backward()
        opt.step()
        return float(loss.item())                          # device -> host read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, K):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for i in range(args.warmup):
        step_resident(i)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = ops.launches
    ops.prof = {}                                          # per-launch CUDA events on the launching stream
    ms = timed(step_resident, args.steps)
    prof, ops.prof = ops.prof, None
    launches = ops.launches - l0
    clk = clocks.stop() if rank == 0 else None
    # ---- end-to-end through the public API with host batches ----
    step_e2e(0)
    ms_e2e = timed(step_e2e, args.steps)
    loss_val = step_e2e(0)

    # ---- roofline of the dominant kernel family (device time from the events recorded above) ----
    fam = {}
    for name, lst in prof.items():
        t = sum(a.elapsed_time(b) for a, b, _ in lst)
        fam[name] = {"ms": t, "work": sum(w for _, _, w in lst), "launches": len(lst)}
    peaks = _peaks()
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)
    roof, fam_out = None, {}
    tot = sum(f["ms"] for f in fam.values()) or 1.0
    for name, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        tensor = name.startswith(("conv", "wgrad", "attn"))
        ach = f["work"] / (f["ms"] * 1e-3) / (1e12 if tensor else 1e9) if f["ms"] > 0 else 0.0
        peak = peaks["bf16_tflops_sustained"] if tensor else peaks["hbm_gbs"]
        fam_out[name] = {"ms_per_step": f["ms"] / args.steps, "launches_per_step": f["launches"] / args.steps,
                         "achieved": ach, "unit": "TFLOP/s" if tensor else "GB/s", "frac": ach / peak}
        if roof is None:
            roof = {"kernel": name, "bound": "tensor" if tensor else "hbm", "achieved": ach, "peak": peak,
                    "unit": "TFLOP/s" if tensor else "GB/s", "frac": ach / peak,
                    "traffic": traffic.get(name, {}).get("traffic_bytes"), "traffic_detail": traffic.get(name),
                    "peak_source": peaks["source"] + (" sustained bf16 (kernel timed inside a long step)" if tensor else " copy bandwidth"),
                    "share_of_timed_kernels": f["ms"] / tot}
    # ---- CFG sampling (the second half of BASELINE.json's metric): configs[3] shape, a bounded number of the 1000 steps ----
    sampling = None
    if args.sample_steps > 0:
        del trainer, opt, dev_pool
        net._state = None
        del net
        torch.cuda.empty_cache()
        from hdiff_b200.DiffusionFreeGuidence.ModelCondition import UNet as CondUNet
        from hdiff_b200.DiffusionFreeGuidence.DiffusionCondition import GaussianDiffusionSampler
        torch.manual_seed(0)
        cfg = dict(CFG2); cfg["dropout"] = 0.0
        snet = CondUNet(num_labels=10, **cfg).to(dev)
        snet.eval()
        sampler = GaussianDiffusionSampler(snet, BETA_1, BETA_T, CFG2["T"], w=1.8).to(dev)
        Bs = args.sample_batch
        torch.manual_seed(2000 + rank)
        x = torch.randn(Bs, 3, res, res, device=dev)
        labels = (torch.arange(Bs, device=dev) + rank * Bs) % 10 + 1
        step = torch.full((1,), CFG2["T"] - 1, dtype=torch.int32, device=dev)
        nan_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        sampler.run_steps(x, labels, step, nan_flag, 3)                     # warm-up (includes graph capture cost once)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l1 = ops.launches
        s0.record()
        sampler.run_steps(x, labels, step, nan_flag, args.sample_steps)
        s1.record()
        barrier()
        sms = torch.tensor([s0.elapsed_time(s1)], device=dev)
        if world > 1:
            dist.all_reduce(sms, op=dist.ReduceOp.MAX)
        per_step = float(sms.item()) / args.sample_steps
        sampling = {"metric": "cfg_sampling_images_per_sec_256", "value": world * Bs / (per_step * 1e-3 * CFG2["T"]), "unit": UNIT,
                    "ms_per_sampler_step": per_step, "sampler_steps_timed": args.sample_steps,
                    "note": f"{args.sample_steps} of the {CFG2['T']} ancestral steps timed (each = one 2B-batch conditional+null UNet forward "
                            "+ fused CFG/posterior update) as replays of the step's CUDA graph (captured once during warm-up, as it is once per 1000-step "
                            "chain), extrapolated to the full chain",
                    "ddim100_images_per_s": world * Bs / (per_step * 1e-3 * 100),
                    "ddim_note": "the same step (same launches; another coefficient table) run as the 100-step deterministic DDIM sampler "
                                 "of the hybrid pipeline (sampler(x_T, labels, ddim=True, ddim_step=100)); derived from the step time above",
                    "batch_per_gpu": Bs, "global_batch": world * Bs, "guidance_w": 1.8, "cuda_graph": bool(sampler.use_cuda_graph),
                    "nan_flag": int(nan_flag.item())}
    value = world * B * args.steps / (ms * 1e-3)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
           "data": "synthetic",
           "config": {"workload": f"cfg2 DDPM UNet ch=64 [1,2,2,2] attn=[1] nrb=2 T=1000 dropout=0.1 train step "
                                  f"(trainer fwd + sum/1000 + bwd + clip 1.0 + AdamW), {res}x{res}",
                      "global_batch": world * B, "per_gpu_batch": B, "resolution": res, "parallelism": f"dp{world}",
                      "l2": f"{n_pool} rotating input batches ({n_pool * B * 3 * res * res * 4 >> 20} MiB) and a multi-GB activation "
                            "working set per step, both larger than the 126 MB L2"},
           "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": world * B * 3 * res * res * 4, "d2h_bytes_per_step": world * 4,
                   "ms_per_step": ms_e2e / args.steps},
           "gpu_launches": launches, "tcgen05_launches_total": ops.tc_launches, "loss": loss_val,
           "roofline": roof, "kernel_families": fam_out, "clocks": clk, "sampling": sampling}
    if world == 1 and not args.no_cpu_baseline:
        v, sec, threads = cpu_reference_steps(args.cpu_steps, 1, res, args.ref_batch)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"{args.cpu_steps} training steps (median, after 1 warm-up) of the fp32 PyTorch reference path "
                                         f"(oracle/ref_torch.py) at batch {args.ref_batch}, {res}x{res}, {threads} torch threads; per image"}
    else:
        out["cpu_baseline"] = None
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (configs[1]: 32)")
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--ref-batch", type=int, default=1, help="batch of the CPU reference sample")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sample-steps", type=int, default=20, help="ancestral sampler steps timed after the training measurement (0: skip)")
    ap.add_argument("--sample-batch", type=int, default=8, help="per-GPU sampling batch (configs[3]: 64 images over 8 GPUs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        assert world == args.gpus or world == 1 and args.gpus == 1, \
            f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})"
        run_ours(args)


if __name__ == "__main__":
    main()
