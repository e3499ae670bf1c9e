"""Drop-in for the reference's `diffusion/Model.py`:

  * `UNet(T=..., ch=..., ch_mult=..., attn=..., num_res_blocks=..., dropout=...)` as its callers construct it
    (diffusion/Train.py:30-31,71-72), `forward(x[B,3,H,W] fp32, t[B] int64) -> eps[B,3,H,W] fp32`;
  * `DynamicUNet(T, ch, ch_mult, num_res_blocks, dropout)` (:382-517), the hybrid pipeline's model:
    `forward(x[B,6,H,W], t, labels=None, context_zero=True) -> eps[B,3,H,W]` with the 6-channel head (:391), the image
    `ConditionalEmbedding` (:110-167), MHA ResBlocks (:267-312; four attention middle blocks :425-431), `num_res_blocks` up
    blocks per level with nearest-interpolated skips (:506-510), the red / blue mean gate that freezes every other
    middle block (:454-474) and the reference's initialisation (:403-407).

Parameter names follow the reference (diffusion/Model.py:18-312) so reference checkpoints load."""
import torch
from torch.nn import init

from .. import ops as _ops
from ..engine import (UNetBase, Swish, TimeEmbedding, DownSample, UpSample, AttnBlock, ResBlock,  # noqa: F401
                      ImageConditionalEmbedding as ConditionalEmbedding)


class UNet(UNetBase):
    def __init__(self, T, ch, ch_mult, attn, num_res_blocks, dropout, compute_dtype=None, mha=False):
        super().__init__(T, ch, ch_mult, attn, num_res_blocks, dropout, num_labels=None, compute_dtype=compute_dtype, mha=mha)


class DynamicUNet(UNetBase):
    def __init__(self, T, ch, ch_mult, num_res_blocks, dropout, compute_dtype=None):
        super().__init__(T, ch, ch_mult, [], num_res_blocks, dropout, num_labels=None, compute_dtype=compute_dtype, mha=True,
                         in_channels=6, middle_attn=(True, True, True, True), up_extra=0, image_cond=True)
        self.initialize()

    def initialize(self):
        """diffusion/Model.py:403-407"""
        init.xavier_uniform_(self.head.weight)
        init.zeros_(self.head.bias)
        init.xavier_uniform_(self.tail[-1].weight, gain=1e-5)
        init.zeros_(self.tail[-1].bias)

    def dynamic_forward(self, x):
        """Red / blue gate (diffusion/Model.py:454-474): if the batch's mean blue exceeds its mean red ("subaquatic") the
        even-indexed middle blocks train and the odd ones are frozen, otherwise the other way round.  The two channel means
        come from one column-sum kernel; like the reference's `if is_subaquatic:` this reads one value back to the host."""
        N, C, H, W = x.shape
        if x.is_cuda:
            tot = torch.zeros(C, dtype=torch.float32, device=x.device)
            _ops.get().colsum(x.contiguous(), N, H * W, C, None, tot, nchw=True)
            red, blue = tot[0], tot[2]
        else:
            red, blue = x[:, 0].sum(), x[:, 2].sum()
        is_subaquatic = bool(blue > red)
        for i, layer in enumerate(self.middleblocks):
            train = (i % 2 == 0) if is_subaquatic else (i % 2 != 0)
            for p in layer.parameters():
                p.requires_grad = train
        return is_subaquatic

    def forward(self, x, t, labels=None, context_zero=True):
        # the gate only decides which parameters RECEIVE gradients: under no_grad (sampling) it has no effect on the result,
        # and its host read-back would break the CUDA-graph capture of a sampler step
        if torch.is_grad_enabled():
            self.dynamic_forward(x)
        return UNetBase.forward(self, x, t, labels, context_zero)
