"""How long / on what data must the oracle be trained before its 1000-step CFG samples (w = 1.8) stop saturating at +-1?
Feeds the pre-training recipe of tests/test_trajectory_gpu.py::test_fixed_noise_1000_step_cfg_sampling_psnr_on_a_trained_network."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_torch as R  # noqa: E402

CFG = dict(T=1000, ch=64, ch_mult=[1, 2, 2, 2], attn=[1], num_res_blocks=2)
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
res = 64
for (amp, shift, steps, lr) in ((0.6, 0.3, 400, 2e-4), (0.35, 0.15, 800, 2e-4), (0.35, 0.15, 1500, 3e-4)):
    torch.manual_seed(2)
    ref = R.UNet(num_labels=10, dropout=0.0, **CFG).to(dev)
    ref.train()
    rtr = R.GaussianDiffusionTrainer(ref, 1e-4, 0.02, 1000).to(dev)
    ropt = torch.optim.AdamW(ref.parameters(), lr=lr, weight_decay=1e-4)
    g = torch.Generator(device="cuda").manual_seed(5)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, res, device=dev), torch.linspace(-1, 1, res, device=dev), indexing="ij")
    t0 = time.time()
    for s in range(steps):
        lab = torch.randint(0, 10, (16,), generator=g, device=dev) + 1
        ph = torch.rand(16, 3, 1, 1, generator=g, device=dev) * 6.28
        fr = 1 + 3 * torch.rand(16, 3, 1, 1, generator=g, device=dev)
        x = amp * torch.sin(fr * xx + ph) * torch.cos(fr * yy - ph) + shift * ((lab.view(-1, 1, 1, 1).float() - 5.5) / 5.5)
        if s % 10 == 0:
            lab = torch.zeros_like(lab)
        loss = R.train_step(rtr, ropt, x.clamp(-1, 1), lab)
    torch.cuda.synchronize()
    t1 = time.time()
    ref.eval()
    xT = torch.randn(2, 3, res, res, device=dev)
    lab = torch.tensor([3, 8], device=dev)
    torch.manual_seed(3)
    with torch.no_grad():
        r0 = R.GaussianDiffusionSampler(ref, 1e-4, 0.02, 1000, w=1.8).to(dev)(xT, lab)
    torch.cuda.synchronize()
    print(f"amp {amp} shift {shift} steps {steps} lr {lr}: train {t1 - t0:.1f}s loss {float(loss):.4f} sample {time.time() - t1:.1f}s "
          f"clipped {float((r0.abs() >= 1.0).float().mean()):.3f} std {float(r0.std()):.3f}", flush=True)
