"""clip_grad_norm_ + AdamW as two kernels over the flat parameter / gradient buffers of a hdiff_b200 UNet
(reference: torch.nn.utils.clip_grad_norm_(net.parameters(), grad_clip) then torch.optim.AdamW.step(),
DiffusionFreeGuidence/TrainCondition.py:39,61-63 — about 360 parameter tensors, i.e. hundreds of tiny kernels).
Numerics follow torch.optim.AdamW (decoupled weight decay, bias correction formed in double) and clip_grad_norm_
(eps 1e-6).  Semantics kept from torch:
  * a parameter whose `.grad` is None is skipped entirely — no weight decay, no state update (the unconditional model's
    `cond_proj.*` never receive a gradient, ModelCondition.py:199-200): the kernels run over the contiguous ranges of the
    flat buffer that hold parameters WITH a gradient;
  * gradients are read from `p.grad`: when those are the views of the engine's flat gradient buffer (the normal case)
    nothing is copied; after gradient accumulation (a second backward without zero_grad, where autograd sums into the
    first buffer) or any other re-binding of `.grad`, they are gathered into the flat buffer first;
  * `param_groups[0]["lr"]` is what an lr scheduler drives (CosineAnnealingLR / GradualWarmupScheduler,
    TrainCondition.py:41-44); `opt.lr` is an alias of it."""
from __future__ import annotations

import torch

from . import ops as _ops


class FlatAdamW:
    def __init__(self, net, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, max_grad_norm=0.0):
        self.net = net
        self.param_groups = [{"params": list(net.parameters()), "lr": float(lr), "initial_lr": float(lr), "betas": tuple(betas),
                              "eps": float(eps), "weight_decay": float(weight_decay)}]
        self.max_grad_norm = max_grad_norm
        self.step_count = 0
        self._m = self._v = self._sq = None
        self._ranges_key = self._ranges = None

    # torch.optim.Optimizer-shaped accessors (lr schedulers read / write param_groups)
    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @lr.setter
    def lr(self, v):
        self.param_groups[0]["lr"] = float(v)

    @property
    def betas(self):
        return self.param_groups[0]["betas"]

    @property
    def eps(self):
        return self.param_groups[0]["eps"]

    @property
    def weight_decay(self):
        return self.param_groups[0]["weight_decay"]

    def zero_grad(self, set_to_none=True):
        for p in self.net.parameters():
            p.grad = None

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self._m, "exp_avg_sq": self._v,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def _active_ranges(self, st):
        """Contiguous [lo, hi) ranges of the flat buffer covering the parameters that have a gradient; gathers gradients that
        do not already live in the flat gradient buffer."""
        fg = st.flat_grad
        base = fg.data_ptr()
        key = []
        for p in st.order:
            g = p.grad
            if g is None:
                key.append(False)
                continue
            key.append(True)
            o = st.offs[id(p)]
            if g.data_ptr() != base + 4 * o or g.dtype != torch.float32 or not g.is_contiguous():
                fg[o:o + p.numel()].copy_(g.reshape(-1))            # accumulated / re-bound gradient: bring it into the flat buffer
        key = tuple(key)
        if key != self._ranges_key:
            ranges, cur = [], None
            for p, a in zip(st.order, key):
                o = st.offs[id(p)]
                n = (p.numel() + 3) // 4 * 4                        # the padding behind a parameter belongs to it
                if a:
                    if cur is not None and cur[1] == o:
                        cur[1] = o + n
                    else:
                        cur = [o, o + n]
                        ranges.append(cur)
                else:
                    cur = None
            self._ranges_key, self._ranges = key, [tuple(r) for r in ranges]
        return self._ranges

    @torch.no_grad()
    def step(self):
        st = self.net._get_state()
        assert st.flat_grad is not None, "call backward() first"
        if self._m is None or self._m.device != st.flat.device or self._m.numel() != st.n_flat:
            self._m = torch.zeros_like(st.flat)
            self._v = torch.zeros_like(st.flat)
            self._sq = torch.zeros(1, dtype=torch.float64, device=st.flat.device)
        ops = _ops.get()
        ranges = self._active_ranges(st)
        if not ranges:
            return
        self.step_count += 1
        g = st.flat_grad
        whole = len(ranges) == 1 and ranges[0] == (0, st.n_flat)
        if self.max_grad_norm > 0:
            if whole:
                ops.sqnorm(g, self._sq)
            else:
                self._sq.zero_()
                for lo, hi in ranges:
                    ops.sqnorm(g[lo:hi], self._sq, accumulate=True)
        grp = self.param_groups[0]
        for lo, hi in ranges:
            ops.adamw_flat(st.flat[lo:hi], g[lo:hi], self._m[lo:hi], self._v[lo:hi], self._sq, self.max_grad_norm, grp["lr"],
                           grp["betas"][0], grp["betas"][1], grp["eps"], grp["weight_decay"], self.step_count)

    def grad_norm(self):
        return float(self._sq.sqrt())
