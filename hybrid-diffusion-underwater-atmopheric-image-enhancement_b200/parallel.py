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


def enable_data_parallel(net, group=None, bucket_bytes=8 << 20, broadcast=True):
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
    red = GradReducer(group, bucket_bytes)
    n = flat_grad.numel()
    step = max(1, bucket_bytes // 4)
    for lo in range(0, n, step):
        red.reduce_async(flat_grad[lo:lo + step])
    red.wait()


class GradReducer:
    """Bucketed mean all-reduce overlapped with the backward pass.

    The UNet's backward schedule calls `ready(lo)` every time the packed weight-gradient buffer is final from
    offset `lo` to its end (blocks finish in reverse order, and the buffer is laid out in block order, so the
    finished region is a growing suffix).  Whenever at least one bucket of finished bytes has accumulated it is
    all-reduced asynchronously: NCCL runs on its own stream behind the kernels already enqueued and under the ones
    that follow.  `finish()` reduces what is left and makes the compute stream wait for all of it."""

    def __init__(self, group, bucket_bytes: int):
        self.group = group
        self.world = dist.get_world_size(group)
        self.bucket = max(1, int(bucket_bytes) // 4)
        self.works = []
        self.buf = None
        self.hi = 0
        self.launched = 0                      # number of collectives issued (tests / bench read it)
        self.before_reduce = None              # optional callable run before a collective is enqueued (stream joins)
        # NCCL averages inside the collective (ReduceOp.AVG); gloo (CPU tests) has no AVG: pre-scale there
        self.avg = dist.get_backend(group) == "nccl"

    def reduce_async(self, chunk: torch.Tensor):
        if self.world == 1 or chunk.numel() == 0:
            return
        if self.before_reduce is not None:
            self.before_reduce()
        if self.avg:
            self.works.append(dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        else:
            chunk.mul_(1.0 / self.world)
            self.works.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.launched += 1

    def attach(self, buf: torch.Tensor, hi: int):
        """`buf[:hi]` is the region that `ready()` walks down."""
        self.buf, self.hi = buf, hi

    def ready(self, lo: int):
        if self.hi - lo >= self.bucket:
            self.reduce_async(self.buf[lo:self.hi])
            self.hi = lo

    def flush(self, *extra: torch.Tensor):
        """Exchange what is left of the walked region plus `extra`, without waiting (more kernels follow)."""
        if self.buf is not None and self.hi > 0:
            self.reduce_async(self.buf[:self.hi])
            self.hi = 0
        for t in extra:
            n = t.numel()
            for lo in range(0, n, self.bucket):
                self.reduce_async(t[lo:lo + self.bucket])

    def finish(self, *extra: torch.Tensor):
        self.flush(*extra)
        self.wait()

    def wait(self):
        for w in self.works:
            w.wait()      # stream dependency only (NCCL): the compute stream waits, the host does not
        self.works = []


def shard_batch(x, rank, world):
    """Sampling: contiguous slice of the batch for this rank (no communication; DiffusionCondition.py:82-98
    is independent per sample)."""
    B = x.shape[0]
    per = (B + world - 1) // world
    return x[rank * per: min(B, (rank + 1) * per)]
