"""Drop-in for the reference's `DiffusionFreeGuidence/ModelCondition.py` UNet:
`UNet(T, num_labels, ch, ch_mult, num_res_blocks, dropout)` (ModelCondition.py:214; callers
TrainCondition.py:33-34,89-90 pass keywords), `forward(x, t, labels[B] int64 in 0..num_labels) -> eps`,
label 0 == null condition (padding_idx=0).  `attn` is the keyword the north-star API adds: the levels
whose down-path ResBlocks get an AttnBlock (ModelCondition.py:92-120,150-153); see SURVEY.md F4."""
from ..engine import (UNetBase, Swish, TimeEmbedding, ConditionalEmbedding, DownSample, UpSample,  # noqa: F401
                      AttnBlock, ResBlock)


class UNet(UNetBase):
    def __init__(self, T, num_labels, ch, ch_mult, num_res_blocks, dropout, attn=(1,), compute_dtype=None):
        super().__init__(T, ch, ch_mult, list(attn), num_res_blocks, dropout, num_labels=num_labels,
                         compute_dtype=compute_dtype)
