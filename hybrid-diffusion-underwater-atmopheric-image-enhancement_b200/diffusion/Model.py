"""Drop-in for the reference's `diffusion/Model.py` UNet, as its callers construct it:
`UNet(T=..., ch=..., ch_mult=..., attn=..., num_res_blocks=..., dropout=...)` (diffusion/Train.py:30-31,71-72),
`forward(x[B,3,H,W] fp32, t[B] int64) -> eps[B,3,H,W] fp32`.  Parameter names follow the reference blocks
(diffusion/Model.py:18-265) so reference checkpoints load."""
from ..engine import UNetBase, Swish, TimeEmbedding, DownSample, UpSample, AttnBlock, ResBlock  # noqa: F401


class UNet(UNetBase):
    def __init__(self, T, ch, ch_mult, attn, num_res_blocks, dropout, compute_dtype=None):
        super().__init__(T, ch, ch_mult, attn, num_res_blocks, dropout, num_labels=None, compute_dtype=compute_dtype)
