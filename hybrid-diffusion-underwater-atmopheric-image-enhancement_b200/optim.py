"""clip_grad_norm_ + AdamW as two kernels over the flat parameter / gradient buffers of a hdiff_b200 UNet
(reference: torch.nn.utils.clip_grad_norm_(net.parameters(), grad_clip) then torch.optim.AdamW.step(),
DiffusionFreeGuidence/TrainCondition.py:39,61-63 — about 360 parameter tensors, i.e. hundreds of tiny kernels).
Numerics follow torch.optim.AdamW (decoupled weight decay, bias correction) and clip_grad_norm_ (eps 1e-6)."""
from __future__ import annotations

import torch

from . import ops as _ops


class FlatAdamW:
    def __init__(self, net, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, max_grad_norm=0.0):
        self.net = net
        self.lr, self.betas, self.eps, self.weight_decay, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.step_count = 0
        self._m = self._v = self._sq = None

    def zero_grad(self, set_to_none=True):
        for p in self.net.parameters():
            p.grad = None

    @torch.no_grad()
    def step(self):
        st = self.net._get_state()
        g = st.flat_grad
        assert g is not None, "call backward() first"
        if self._m is None or self._m.device != st.flat.device or self._m.numel() != st.n_flat:
            self._m = torch.zeros_like(st.flat)
            self._v = torch.zeros_like(st.flat)
            self._sq = torch.zeros(1, dtype=torch.float64, device=st.flat.device)
        ops = _ops.get()
        self.step_count += 1
        if self.max_grad_norm > 0:
            ops.sqnorm(g, self._sq)
        ops.adamw_flat(st.flat, g, self._m, self._v, self._sq, self.max_grad_norm, self.lr, self.betas[0], self.betas[1],
                       self.eps, self.weight_decay, self.step_count)

    def grad_norm(self):
        return float(self._sq.sqrt())
