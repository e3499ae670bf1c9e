"""Drop-in for the reference's `DiffusionFreeGuidence/ModelCondition.py` UNet:
`UNet(T, num_labels, ch, ch_mult, num_res_blocks, dropout)` (ModelCondition.py:214; callers
TrainCondition.py:33-34,89-90 pass keywords), `forward(x, t, labels[B] int64 in 0..num_labels) -> eps`,
label 0 == null condition (padding_idx=0).

Two block variants exist in the reference file and both are provided:
  * `attn=[levels]` (the keyword the north-star API adds; default `(1,)`): ResBlock_old + AttnBlock (ModelCondition.py:92-164) —
    GroupNorm, single-head spatial attention, residual — on the listed down levels and the first middle block.  This is
    the benchmarked configuration; its attention runs on the tcgen05 flash kernels.
  * `mha=True`: the file's LIVE `UNet` (:213-276): every down ResBlock and the first middle block carry
    `nn.MultiheadAttention(out_ch, 8)` whose output replaces h (ResBlock :166-211).  State-dict keys are the reference's
    (`attn.in_proj_weight`, `attn.in_proj_bias`, `attn.out_proj.*`), so live reference checkpoints load.  Pass `attn=[...]` as
    well to restrict the attention levels (the reference materialises [S, S] scores per head and cannot run its own
    configuration beyond small images; this implementation is flash-style and can)."""
from ..engine import (UNetBase, Swish, TimeEmbedding, ConditionalEmbedding, DownSample, UpSample,  # noqa: F401
                      AttnBlock, ResBlock)


class UNet(UNetBase):
    def __init__(self, T, num_labels, ch, ch_mult, num_res_blocks, dropout, attn=None, compute_dtype=None, mha=False):
        if attn is None:
            attn = "all" if mha else (1,)
        super().__init__(T, ch, ch_mult, attn if attn == "all" else list(attn), num_res_blocks, dropout, num_labels=num_labels,
                         compute_dtype=compute_dtype, mha=mha)
