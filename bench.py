"""bench.py — DDPM UNet training throughput (images/s) at 256x256 on N B200s, the metric BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 our arm (hdiff_b200 CUDA path)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  reference arm: the reference's fp32 PyTorch path on host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

A step is one pass of the hot path over one batch.
  N = 1: BASELINE.json configs[1] — unconditional UNet ch=64 ch_mult=[1,2,2,2] attn=[1] num_res_blocks=2 T=1000 dropout=0.1,
         256x256 RGB, batch 32, bf16 compute / fp32 master weights:
         GaussianDiffusionTrainer.forward(x_0) -> `.sum()/1000.` -> backward -> clip_grad_norm_(1.0) -> AdamW (diffusion/Train.py:49-55).
  N > 1: BASELINE.json configs[2] — the CONDITIONAL UNet (num_labels=10) of the same size, labels+1, whole-batch label dropout
         with a per-rank host coin (p = 0.1), `.sum()/b**2` (DiffusionFreeGuidence/TrainCondition.py:53-63), batch 32 per GPU,
         data parallel over NCCL.
Synthetic data, random-init weights.

One JSON line on stdout (rank 0):
  value      timed region = K steps with inputs resident in HBM, NO per-launch instrumentation inside it;
  e2e        the same step driven with pinned HOST batches (host->device copy of x_0 [and labels], device->host read of the loss);
  roofline / kernel_families   from a SEPARATE profiling pass after the timed ones (CUDA events around every launch);
  sampling   the second half of the metric: the full 1000-step CFG chain (w = 1.8) on 8 images per GPU through
             GaussianDiffusionSampler.forward, its own roofline and CPU baseline;
  gpu_reference  (N = 1) the same training step by stock PyTorch on the same B200: the oracle in fp32 (TF32 off) at the largest
             batch that fits, and under bf16 autocast with F.scaled_dot_product_attention at batch 32;
  cpu_baseline   (N = 1) the reference's CPU path on the host cores, bounded sample.
`oracle/` is imported only by the cpu_baseline / gpu_reference legs and by `--impl reference`.
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
NUM_LABELS = 10
BETA_1, BETA_T = 1e-4, 0.02
METRIC = "ddpm_train_images_per_sec_256"
UNIT = "images/s"
FWD_GFLOP_PER_IMAGE_256 = 488.0          # conv 212.6 + attention 275.4 GFLOP forward per image at 256x256


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
# reference arm / cpu_baseline: the reference's fp32 PyTorch path on host cores
# ---------------------------------------------------------------------------------------------------
def _reference_model(cond, dropout=None):
    """(net, trainer class, sampler class, kind): the reference's OWN classes when its sources are reachable (/root/reference in
    the build container, oracle/_ref on the GPU box — `oracle/make_ref.py` puts them there), else the restatement."""
    from oracle import ref_loader, ref_torch as R
    cfg = dict(CFG2)
    if dropout is not None:
        cfg["dropout"] = dropout
    if ref_loader.available():
        net = ref_loader.assemble_unet(num_labels=NUM_LABELS if cond else None, **cfg)
        dc = ref_loader.diffusion_condition()
        return net, _RefTrainer(dc), dc.GaussianDiffusionSampler, "reference"
    return R.UNet(num_labels=NUM_LABELS if cond else None, **cfg), R.GaussianDiffusionTrainer, R.GaussianDiffusionSampler, "port"


class _RefTrainer:
    """GaussianDiffusionTrainer of DiffusionCondition.py:19-46 takes (x_0, labels); the unconditional twin
    (diffusion/Diffusion.py:304-314, commented there) is the same algorithm without labels — driven here through a
    one-argument model adapter so that the reference class itself runs."""

    def __init__(self, dc):
        self.dc = dc

    def __call__(self, net, beta_1, beta_T, T):
        import torch.nn as nn
        dc = self.dc

        class Uncond(nn.Module):
            def __init__(self, m):
                super().__init__()
                self.m = m

            def forward(self, x, t, labels):
                return self.m(x, t) if labels is None else self.m(x, t, labels)

        tr = dc.GaussianDiffusionTrainer(Uncond(net), beta_1, beta_T, T)
        tr.hd_net = net
        return tr


def cpu_reference_steps(steps, warmup, res, batch, threads=None, cond=False):
    """Times `steps` training steps of the reference fp32 path on the host.  Returns (images/s, s per step, threads, kind)."""
    import numpy as np
    import torch
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net, Trainer, _, kind = _reference_model(cond)
    tr = Trainer(net, BETA_1, BETA_T, CFG2["T"])
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=1e-4)
    x = torch.rand(batch, 3, res, res) * 2 - 1
    rng = np.random.RandomState(0)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        if cond:
            labels = torch.randint(0, NUM_LABELS, (batch,)) + 1
            if rng.rand() < 0.1:
                labels = torch.zeros_like(labels)
            loss = tr(x, labels).sum() / batch ** 2.
        else:
            loss = (tr(x, None) if kind == "reference" else tr(x)).sum() / 1000.
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return batch / med, med, threads, kind


def cpu_reference_sampling(n_steps, res, batch, threads=None):
    """`n_steps` of the 1000 CFG ancestral steps (2 UNet evaluations each) of the reference on the host, extrapolated.
    Returns (images/s for the full chain, s per sampler step, threads, kind)."""
    import torch
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net, _, Sampler, kind = _reference_model(True, dropout=0.0)
    net.eval()
    smp = Sampler(net, BETA_1, BETA_T, n_steps, w=1.8)
    x = torch.randn(batch, 3, res, res)
    labels = torch.arange(batch) % NUM_LABELS + 1
    import contextlib
    import io
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):   # the reference prints the step index (DiffusionCondition.py:88)
        Sampler(net, BETA_1, BETA_T, 2, w=1.8)(x, labels)             # warm-up: a 2-step chain
        t0 = time.perf_counter()
        smp(x, labels)
        sec = (time.perf_counter() - t0) / n_steps
    return batch / (sec * CFG2["T"]), sec, threads, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res, batch = args.res, args.ref_batch
    cond = args.gpus > 1
    v, sec, threads, kind = cpu_reference_steps(args.steps, max(1, min(args.warmup, 1)), res, batch, cond=cond)
    what = "the reference's own classes (oracle/ref_loader.assemble_unet + DiffusionCondition.GaussianDiffusionTrainer)" \
        if kind == "reference" else "the oracle restatement (oracle/ref_torch.py; the reference sources are not reachable here)"
    sample = (f"{args.steps} timed training steps (median) of the fp32 PyTorch reference path on the host — {what} — batch {batch} at "
              f"{res}x{res} (the reference materialises [S,S] attention scores: ~1 GB per image per block at S=16384), reported per image")
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": _workload(cond, res), "global_batch": batch, "resolution": res},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def _workload(cond, res):
    if cond:
        return (f"cfg3 conditional DDPM UNet (num_labels=10) ch=64 [1,2,2,2] attn=[1] nrb=2 T=1000 dropout=0.1, label-dropout CFG train step "
                f"(labels+1, whole-batch drop p=0.1, trainer fwd + sum/b^2 + bwd + clip 1.0 + AdamW), {res}x{res}")
    return (f"cfg2 DDPM UNet ch=64 [1,2,2,2] attn=[1] nrb=2 T=1000 dropout=0.1 train step "
            f"(trainer fwd + sum/1000 + bwd + clip 1.0 + AdamW), {res}x{res}")


# ---------------------------------------------------------------------------------------------------
# same-GPU reference: stock PyTorch (cuDNN / cuBLAS) running the oracle
# ---------------------------------------------------------------------------------------------------
def gpu_reference(dev, res, steps=3):
    """cfg2 'vs reference fp32': the oracle's training step on this GPU in fp32 with TF32 off at the largest batch that fits
    (it materialises the [S,S] attention scores), and the fastest stock-PyTorch form of the same step (bf16 autocast,
    F.scaled_dot_product_attention, batch 32).  Reported, not the target."""
    import torch
    from oracle import ref_torch as R
    out = {}

    def run(batch, autocast, sdpa):
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        torch.manual_seed(0)
        R.AttnBlock.use_sdpa = sdpa
        net = R.UNet(**CFG2).to(dev).train()
        tr = R.GaussianDiffusionTrainer(net, BETA_1, BETA_T, CFG2["T"]).to(dev)
        opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=1e-4)
        x = torch.rand(batch, 3, res, res, device=dev) * 2 - 1
        ts = []
        for i in range(steps + 1):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                R.train_step(tr, opt, x)
            e1.record()
            torch.cuda.synchronize()
            if i:
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        for b in (8, 4, 2, 1):
            try:
                ms = run(b, False, False)
                out["fp32"] = {"value": b / (ms * 1e-3), "unit": UNIT, "batch": b, "ms_per_step": ms,
                               "what": "oracle/ref_torch.py (restatement of the reference modules) in fp32, TF32 off, cuDNN/cuBLAS, "
                                       "largest batch of (8, 4, 2, 1) that fits with materialised [S,S] attention scores"}
                break
            except torch.cuda.OutOfMemoryError:
                torch.cuda.empty_cache()
        try:
            ms = run(32, True, True)
            out["bf16_sdpa"] = {"value": 32 / (ms * 1e-3), "unit": UNIT, "batch": 32, "ms_per_step": ms,
                                "what": "the same modules under torch.autocast(bfloat16) with F.scaled_dot_product_attention in AttnBlock "
                                        "(cuDNN convolutions + the library flash attention), batch 32: the fastest stock-PyTorch form of the step"}
        except Exception as e:                           # noqa: BLE001  (reported, never fatal)
            out["bf16_sdpa"] = {"unavailable": repr(e)[:200]}
    finally:
        R.AttnBlock.use_sdpa = False
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
        torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def hybrid_leg(dev, res, ops, steps=5, batch=16):
    """SURVEY 8(f) rank 2, the product's own model and settings (Main.py:16-35): DynamicUNet ch=128 ch_mult=[1,2,2,2] num_res_blocks=2
    dropout=0.15, batch 16, 256x256 uint8 image pairs — one hybrid training step (trainer forward -> sum/1000 -> backward -> clip ->
    AdamW) and the 100-step DDIM sampler.  Extra information next to the headline metric, not part of it."""
    import torch
    from hdiff_b200.diffusion.Model import DynamicUNet
    from hdiff_b200.diffusion.Diffusion import GaussianDiffusionTrainer, GaussianDiffusionSampler
    from hdiff_b200.optim import FlatAdamW
    torch.manual_seed(0)
    net = DynamicUNet(T=1000, ch=128, ch_mult=[1, 2, 2, 2], num_res_blocks=2, dropout=0.15).to(dev)
    net.train()
    tr = GaussianDiffusionTrainer(net, BETA_1, BETA_T, 1000).to(dev)
    opt = FlatAdamW(net, lr=1e-4, weight_decay=1e-4, max_grad_norm=1.0)
    g = torch.Generator(device=dev).manual_seed(1)
    gt = [torch.randint(0, 256, (batch, 3, res, res), generator=g, device=dev, dtype=torch.uint8) for _ in range(3)]
    inp = [torch.randint(0, 256, (batch, 3, res, res), generator=g, device=dev, dtype=torch.uint8) for _ in range(3)]
    for k in range(3):
        opt.zero_grad(); (tr(gt[k], inp[k], 0)[0].sum() / 1000.).backward(); opt.step()
    torch.cuda.synchronize()
    l0 = ops.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        opt.zero_grad()
        loss = tr(gt[k % 3], inp[k % 3], 0)[0].sum() / 1000.
        loss.backward()
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"model": "DynamicUNet ch=128 ch_mult=[1,2,2,2] num_res_blocks=2 dropout=0.15 (Main.py:16-35), 256x256, batch 16, uint8 pairs",
           "params_m": sum(p.numel() for p in net.parameters()) / 1e6,
           "train_images_per_s": batch / (ms * 1e-3), "train_ms_per_step": ms, "train_launches_per_step": (ops.launches - l0) / steps,
           "train_loss": float(loss)}
    del tr, opt
    net.eval()
    smp = GaussianDiffusionSampler(net, BETA_1, BETA_T, 1000).to(dev)
    img = inp[0][:8].contiguous()
    smp(img, ddim=True, unconditional_guidance_scale=1, ddim_step=4)        # warm-up: a 4-step chain (stride 250)
    torch.cuda.synchronize()
    secs = []
    for _ in range(2):                      # the public call captures its step graph every time (host-side cost: 0.1-0.8 s on these boxes)
        e0.record()
        y = smp(img, ddim=True, unconditional_guidance_scale=1, ddim_step=100)
        e1.record()
        torch.cuda.synchronize()
        secs.append(e0.elapsed_time(e1) * 1e-3)
    sec = min(secs)
    out.update({"ddim100_images_per_s": 8 / sec, "ddim100_chain_seconds": sec, "ddim100_chain_seconds_each_call": secs, "ddim_batch": 8,
                "ddim_out_abs_max": float(y.abs().max())})
    del net, smp
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from hdiff_b200.diffusion.Model import UNet
    from hdiff_b200.DiffusionFreeGuidence.ModelCondition import UNet as CondUNet
    from hdiff_b200.diffusion.Diffusion import GaussianDiffusionTrainer
    from hdiff_b200.DiffusionFreeGuidence.DiffusionCondition import GaussianDiffusionSampler
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
    cond = world > 1 or args.cond                          # N > 1: configs[2], the conditional model with label dropout
    torch.manual_seed(0)                                   # identical replicas
    net = (CondUNet(num_labels=NUM_LABELS, **CFG2) if cond else UNet(**CFG2)).to(dev)
    net.train()
    if world > 1:
        parallel.enable_data_parallel(net)
    trainer = GaussianDiffusionTrainer(net, BETA_1, BETA_T, CFG2["T"]).to(dev)
    opt = FlatAdamW(net, lr=1e-4, weight_decay=1e-4, max_grad_norm=1.0)
    ops = hops.get()
    torch.manual_seed(1000 + rank)                         # per-rank data and RNG stream
    coin = np.random.RandomState(1000 + rank)              # the reference's per-process host coin (TrainCondition.py:57)
    n_pool = 6                                             # 6 x 25 MB of distinct batches > the 126 MB L2
    host_pool = [(torch.rand(B, 3, res, res) * 2 - 1).pin_memory() for _ in range(n_pool)]
    host_lab = [torch.randint(0, NUM_LABELS, (B,)).pin_memory() for _ in range(n_pool)]
    dev_pool = [h.to(dev) for h in host_pool]
    dev_lab = [h.to(dev) for h in host_lab]

    def loss_of(x, labels):
        if not cond:
            return trainer(x).sum() / 1000.
        labels = labels + 1
        if coin.rand() < 0.1:
            labels = torch.zeros_like(labels)
        return trainer(x, labels).sum() / B ** 2.

    def step_resident(i):
        opt.zero_grad()
        loss = loss_of(dev_pool[i % n_pool], dev_lab[i % n_pool])
        loss.backward()
        opt.step()
        return loss

    def step_e2e(i):
        x = host_pool[i % n_pool].to(dev, non_blocking=True)
        lab = host_lab[i % n_pool].to(dev, non_blocking=True) if cond else None
        opt.zero_grad()
        loss = loss_of(x, lab)
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
    ms = timed(step_resident, args.steps)                  # the `value` region: no per-launch events, no host reads
    launches = ops.launches - l0
    clk = clocks.stop() if rank == 0 else None
    # ---- end-to-end through the public API with host batches ----
    step_e2e(0)
    ms_e2e = timed(step_e2e, args.steps)
    loss_val = step_e2e(0)
    # ---- separate profiling pass: CUDA events around every launch, on the launching stream ----
    n_prof = max(1, min(args.steps, args.profile_steps))
    ops.prof = {}
    timed(step_resident, n_prof)
    prof, ops.prof = ops.prof, None

    fam = {}
    peaks = _peaks()
    for name, lst in prof.items():
        t = sum(r[0].elapsed_time(r[1]) for r in lst)
        fam[name] = {"ms": t, "work": sum(r[2] for r in lst), "launches": len(lst)}
        if any(len(r) > 3 for r in lst):
            # per launch, the roofline that BINDS it: the longer of (algorithmic FLOPs / sustained tensor peak) and (compulsory bytes / copy
            # bandwidth).  The 1x1 convolutions and the 64-channel layers at 256x256 move more bytes than their FLOPs can hide.
            lb = [(r[2] / (peaks["bf16_tflops_sustained"] * 1e12), (r[3] if len(r) > 3 else 0.0) / (peaks["hbm_gbs"] * 1e9)) for r in lst]
            fam[name]["bound_ms"] = 1e3 * sum(max(a, b) for a, b in lb)
            fam[name]["hbm_bound_launches"] = sum(1 for a, b in lb if b > a)
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)
    roof, fam_out = None, {}
    tot = sum(f["ms"] for f in fam.values()) or 1.0
    for name, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        tensor = name.startswith(("conv", "wgrad", "attn", "mha"))
        ach = f["work"] / (f["ms"] * 1e-3) / (1e12 if tensor else 1e9) if f["ms"] > 0 else 0.0
        peak = peaks["bf16_tflops_sustained"] if tensor else peaks["hbm_gbs"]
        fam_out[name] = {"ms_per_step": f["ms"] / n_prof, "launches_per_step": f["launches"] / n_prof,
                         "achieved": ach, "unit": "TFLOP/s" if tensor else "GB/s", "frac": ach / peak}
        if "bound_ms" in f:
            fam_out[name]["frac_of_binding_roofline"] = f["bound_ms"] / f["ms"] if f["ms"] > 0 else 0.0
            fam_out[name]["hbm_bound_launches_per_step"] = f["hbm_bound_launches"] / n_prof
        if roof is None:
            roof = {"kernel": name, "bound": "tensor" if tensor else "hbm", "achieved": ach, "peak": peak,
                    "unit": "TFLOP/s" if tensor else "GB/s", "frac": ach / peak,
                    "traffic": traffic.get(name, {}).get("traffic_bytes"), "traffic_detail": traffic.get(name),
                    "peak_source": peaks["source"] + (" sustained bf16 (kernel timed inside a long step)" if tensor else " copy bandwidth"),
                    "share_of_timed_kernels": f["ms"] / tot,
                    "timed_in": f"separate profiling pass of {n_prof} step(s) after the timed regions (per-launch CUDA events)"}
    step_tflops = 3 * FWD_GFLOP_PER_IMAGE_256 * (res / 256.) ** 2 * B / (ms / args.steps)      # GFLOP / ms = TFLOP/s (attention share scales as res^4: 256 only)

    # ---- CFG sampling (the second half of BASELINE.json's metric): configs[3] — the full 1000-step chain ----
    sampling = None
    if args.sample_steps > 0:
        del trainer, opt, dev_pool
        net._state = None
        del net
        torch.cuda.empty_cache()
        torch.manual_seed(0)
        cfg = dict(CFG2); cfg["dropout"] = 0.0
        snet = CondUNet(num_labels=NUM_LABELS, **cfg).to(dev)
        snet.eval()
        T_s = args.sample_steps
        sampler = GaussianDiffusionSampler(snet, BETA_1, BETA_T, T_s, w=1.8).to(dev)
        Bs = args.sample_batch
        torch.manual_seed(2000 + rank)
        xT = torch.randn(Bs, 3, res, res, device=dev)
        labels = (torch.arange(Bs, device=dev) + rank * Bs) % NUM_LABELS + 1
        warm = GaussianDiffusionSampler(snet, BETA_1, BETA_T, 4, w=1.8).to(dev)
        warm(xT, labels)                                                    # warm-up: a 4-step chain (lazy state, first graph capture)
        barrier()
        l1 = ops.launches
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        x0 = sampler(xT, labels)                                            # the public call: T steps, graph capture once, NaN check
        s1.record()
        barrier()
        sms = torch.tensor([s0.elapsed_time(s1)], device=dev)
        if world > 1:
            dist.all_reduce(sms, op=dist.ReduceOp.MAX)
        chain_s = float(sms.item()) * 1e-3
        per_step = chain_s / T_s
        full = T_s == CFG2["T"]
        tf = 2 * Bs * FWD_GFLOP_PER_IMAGE_256 * (res / 256.) ** 2 / (per_step * 1e3)             # 2B forward passes per step
        sampling = {"metric": "cfg_sampling_images_per_sec_256", "value": world * Bs / (per_step * CFG2["T"]), "unit": UNIT,
                    "chain_seconds": chain_s, "ms_per_sampler_step": per_step * 1e3, "sampler_steps_timed": T_s,
                    "note": ("the FULL 1000-step ancestral chain through GaussianDiffusionSampler.forward" if full else
                             f"{T_s} of the 1000 ancestral steps (extrapolated)") +
                            " (each step = one 2B-batch conditional+null UNet forward + fused CFG/posterior update; the step's CUDA graph is "
                            "captured once inside the call and replayed)",
                    "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                 "frac": tf / peaks["bf16_tflops_sustained"],
                                 "work": f"2 x {Bs} forward passes x {FWD_GFLOP_PER_IMAGE_256} GFLOP algorithmic per sampler step"},
                    "ddim100_images_per_s": world * Bs / (per_step * 100),
                    "ddim_note": "the same step (same launches; another coefficient table) run as the 100-step deterministic DDIM sampler "
                                 "of the hybrid pipeline (sampler(x_T, labels, ddim=True, ddim_step=100)); derived from the step time above",
                    "batch_per_gpu": Bs, "global_batch": world * Bs, "guidance_w": 1.8, "cuda_graph": bool(sampler.use_cuda_graph),
                    "launches": ops.launches - l1, "out_abs_max": float(x0.abs().max())}
        del snet, sampler, warm
        torch.cuda.empty_cache()
    hybrid = None
    if world == 1 and not args.no_hybrid:
        try:
            hybrid = hybrid_leg(dev, res, ops)
        except Exception as e:                                              # noqa: BLE001
            hybrid = {"unavailable": repr(e)[:300]}
    value = world * B * args.steps / (ms * 1e-3)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)
    gpu_ref = None
    if world == 1 and not args.no_gpu_reference:
        try:
            gpu_ref = gpu_reference(dev, res)
        except Exception as e:                                              # noqa: BLE001
            gpu_ref = {"unavailable": repr(e)[:300]}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    h2d = B * 3 * res * res * 4 + (B * 8 if cond else 0)
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
           "data": "synthetic",
           "config": {"workload": _workload(cond, res),
                      "global_batch": world * B, "per_gpu_batch": B, "resolution": res, "parallelism": f"dp{world}",
                      "l2": f"{n_pool} rotating input batches ({n_pool * B * 3 * res * res * 4 >> 20} MiB) and a multi-GB activation "
                            "working set per step, both larger than the 126 MB L2"},
           "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": world * h2d, "d2h_bytes_per_step": world * 4,
                   "ms_per_step": ms_e2e / args.steps},
           "gpu_launches": launches, "tcgen05_launches_total": ops.tc_launches, "loss": loss_val,
           "step_algorithmic_tflops": step_tflops, "step_frac_of_sustained_bf16": step_tflops / peaks["bf16_tflops_sustained"],
           "roofline": roof, "kernel_families": fam_out, "clocks": clk, "sampling": sampling, "gpu_reference": gpu_ref, "hybrid": hybrid}
    if world == 1 and not args.no_cpu_baseline:
        v, sec, threads, kind = cpu_reference_steps(args.cpu_steps, 1, res, args.ref_batch, cond=cond)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
                               "sample": f"{args.cpu_steps} training steps (median, after 1 warm-up) of the fp32 PyTorch reference path "
                                         f"({'the reference classes via oracle/ref_loader' if kind == 'reference' else 'oracle/ref_torch.py'}) "
                                         f"at batch {args.ref_batch}, {res}x{res}, {threads} torch threads; per image"}
        if sampling is not None and args.cpu_sample_steps > 0:
            sv, ssec, threads, kind = cpu_reference_sampling(args.cpu_sample_steps, res, 1)
            sampling["cpu_baseline"] = {"value": sv, "unit": UNIT, "cores": threads, "kind": kind,
                                        "sample": f"{args.cpu_sample_steps} CFG sampler steps (2 UNet evaluations each, batch 1, {res}x{res}) "
                                                  f"after a 2-step warm-up chain, {ssec:.2f} s per step, extrapolated to the 1000-step chain"}
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
    ap.add_argument("--cond", action="store_true", help="run the conditional (cfg3) step at N = 1 too")
    ap.add_argument("--ref-batch", type=int, default=1, help="batch of the CPU reference sample")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--cpu-sample-steps", type=int, default=4, help="CFG sampler steps of the CPU baseline (0: skip)")
    ap.add_argument("--profile-steps", type=int, default=3, help="steps of the separate per-launch profiling pass")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hybrid", action="store_true", help="skip the DynamicUNet / hybrid pipeline leg")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--sample-steps", type=int, default=1000, help="length of the CFG sampling chain timed after training (0: skip)")
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
