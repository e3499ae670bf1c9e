"""Data-parallel training of the hot path: one process per GPU, identical replicas, ONE exchange per step —
the mean of the flat gradient buffer over ranks (the reference wraps the net in
torch.nn.parallel.DistributedDataParallel over NCCL, utils/rotinas.py:572-577,618-619).

All parameter gradients of the UNet live in one contiguous fp32 buffer that the backward kernels write
directly, so the exchange is a handful of large NCCL all-reduces over NVLink 5 / NVSwitch (in-switch
reduction when NCCL enables NVLS) issued on NCCL's stream from inside the backward pass; the optimizer
waits on them through the stream dependency, the host never blocks.  Sampling shards the batch and needs
no communication."""
from __future__ import annotations

import torch
import torch.distributed as dist


def enable_data_parallel(net, group=None, bucket_bytes=64 << 20, broadcast=True):
    """Attach a process group to a hdiff_b200 UNet.  Parameters are broadcast from rank 0 once."""
    assert dist.is_initialized(), "init torch.distributed first (backend nccl on GPUs, gloo on CPU tests)"
    group = group if group is not None else dist.group.WORLD
    net.dp_group = group
    net.dp_bucket_bytes = int(bucket_bytes)
    if broadcast:
        st = net._get_state()
        dist.broadcast(st.flat, src=dist.get_global_rank(group, 0), group=group)
    return net


def allreduce_flat_(flat_grad: torch.Tensor, group, bucket_bytes: int):
    """In-place mean over the ranks of `group`, bucketed; asynchronous with respect to the host."""
    world = dist.get_world_size(group)
    if world == 1:
        return
    n = flat_grad.numel()
    step = max(1, bucket_bytes // 4)
    works = []
    for lo in range(0, n, step):
        chunk = flat_grad[lo:lo + step]
        chunk.mul_(1.0 / world)
        works.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=group, async_op=True))
    for w in works:
        w.wait()          # stream dependency only (NCCL): the compute stream waits, the host does not


def shard_batch(x, rank, world):
    """Sampling: contiguous slice of the batch for this rank (no communication; DiffusionCondition.py:82-98
    is independent per sample)."""
    B = x.shape[0]
    per = (B + world - 1) // world
    return x[rank * per: min(B, (rank + 1) * per)]
