"""Data-parallel path on CPU: two `gloo` ranks, the torch test double of the operator layer standing in for the CUDA
library.  Checks the one exchange of the path: after backward every rank holds the MEAN over ranks of
the per-rank gradients — equal to the single-process gradient of the concatenated batch divided by the world size —
with the packed weight-gradient buffer reduced in several overlapped buckets, plus the batch sharding of sampling."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import hdiff_b200.ops as hops
    from hdiff_b200 import parallel
    from hdiff_b200.DiffusionFreeGuidence.ModelCondition import UNet
    from tests.emu_backend import EmuOps
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    hops.set_backend(EmuOps())
    cfg = dict(T=50, ch=32, ch_mult=[1, 2], attn=[1], num_res_blocks=1, dropout=0.0)
    torch.manual_seed(123 + rank)                 # different initial weights per rank: the broadcast must fix that
    net = UNet(num_labels=4, compute_dtype=torch.float32, **cfg)
    parallel.enable_data_parallel(net, bucket_bytes=256 << 10)
    torch.manual_seed(7)
    x = torch.randn(4, 3, 16, 16)
    t = torch.tensor([3, 9, 20, 41])
    lab = torch.tensor([1, 0, 4, 2])
    gy = torch.randn(4, 3, 16, 16)
    xs = parallel.shard_batch(x, rank, world)
    sl = slice(rank * 2, rank * 2 + 2)
    assert torch.equal(xs, x[sl])
    net(xs, t[sl], lab[sl]).backward(gy[sl])
    n_coll = net.last_reducer.launched
    got = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    # single-process reference on the full batch, same weights, no process group
    ref = UNet(num_labels=4, compute_dtype=torch.float32, **cfg)
    ref.load_state_dict(sd)
    ref(x, t, lab).backward(gy)
    worst = 0.0
    for k, p in ref.named_parameters():
        if p.grad is None:
            continue
        d = float((got[k] - p.grad / world).norm())
        worst = max(worst, d / (float(p.grad.norm()) / world + 1e-6))
    q.put((rank, worst, n_coll, float(sd["head.weight"].sum())))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gradient_mean_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert abs(res[0][3] - res[1][3]) < 1e-6, "parameters were not broadcast from rank 0"
    for rank, worst, n_coll, _ in res:
        assert worst < 2e-4, (rank, worst)
        assert n_coll >= 3, "expected several overlapped buckets plus the tail exchange"


def _worker_stock_ddp(rank, world, port, q, cond):
    """The reference's own wrap: torch.nn.parallel.DistributedDataParallel(net) (utils/rotinas.py:618-619)."""
    import hdiff_b200.ops as hops
    from hdiff_b200.DiffusionFreeGuidence.ModelCondition import UNet as UNetC
    from hdiff_b200.diffusion.Model import UNet as UNetU
    from tests.emu_backend import EmuOps
    from torch.nn.parallel import DistributedDataParallel as DDP
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    hops.set_backend(EmuOps())
    cfg = dict(T=50, ch=32, ch_mult=[1, 2], attn=[1], num_res_blocks=1, dropout=0.0)
    torch.manual_seed(321 + rank)                 # DDP broadcasts rank 0's parameters at construction
    net = UNetC(num_labels=4, compute_dtype=torch.float32, **cfg) if cond else UNetU(compute_dtype=torch.float32, **cfg)
    # the unconditional model's cond_proj.* never receive a gradient (as in the reference): DDP must be told, as it must for
    # the reference's own unconditional model
    ddp = DDP(net, find_unused_parameters=not cond)
    torch.manual_seed(7)
    x = torch.randn(4, 3, 16, 16)
    t = torch.tensor([3, 9, 20, 41])
    lab = torch.tensor([1, 0, 4, 2])
    gy = torch.randn(4, 3, 16, 16)
    sl = slice(rank * 2, rank * 2 + 2)
    worst = 0.0
    for it in range(2):                           # two iterations: DDP's reducer must be re-armed correctly by our autograd node
        for p in net.parameters():
            p.grad = None
        out = ddp(x[sl], t[sl], lab[sl]) if cond else ddp(x[sl], t[sl])
        out.backward(gy[sl])
        got = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
        ref = UNetC(num_labels=4, compute_dtype=torch.float32, **cfg) if cond else UNetU(compute_dtype=torch.float32, **cfg)
        ref.load_state_dict(net.state_dict())
        (ref(x, t, lab) if cond else ref(x, t)).backward(gy)
        for k, p in ref.named_parameters():
            if p.grad is None:
                assert k not in got, k
                continue
            d = float((got[k] - p.grad / world).norm())
            worst = max(worst, d / (float(p.grad.norm()) / world + 1e-6))
    q.put((rank, worst, float(net.state_dict()["head.weight"].sum())))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("cond", [True, False], ids=["conditional", "unconditional_find_unused"])
def test_stock_distributed_data_parallel_wrap(cond):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_stock_ddp, args=(r, world, port, q, cond)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert abs(res[0][2] - res[1][2]) < 1e-6, "DDP did not broadcast rank 0's parameters into the flat buffer"
    for rank, worst, _ in res:
        assert worst < 2e-4, (rank, worst)


def _worker_dynamic(rank, world, port, q, context_zero):
    """DynamicUNet under hdiff_b200's data parallelism.  With the image condition encoder in use its convolutions are the LAST
    weight gradients of a backward pass (engine: `late_conv`): the exchange of the packed buffer must wait for them.  The gate
    freezes every other middle block: those parameters must end without a gradient on every rank."""
    import hdiff_b200.ops as hops
    from hdiff_b200 import parallel
    from hdiff_b200.diffusion.Model import DynamicUNet
    from tests.emu_backend import EmuOps
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    hops.set_backend(EmuOps())
    cfg = dict(T=50, ch=32, ch_mult=[1, 2], num_res_blocks=1, dropout=0.0)
    torch.manual_seed(55 + rank)
    net = DynamicUNet(compute_dtype=torch.float32, **cfg)
    with torch.no_grad():
        net.tail[-1].weight.mul_(3e4)
    parallel.enable_data_parallel(net, bucket_bytes=64 << 10)
    torch.manual_seed(8)
    x = torch.randn(4, 6, 16, 16)
    x[:, 2] += 1.0                                  # every shard "subaquatic": the same gate decision on both ranks
    t = torch.tensor([3, 9, 20, 41])
    lab = torch.randn(4, 3, 16, 16)
    gy = torch.randn(4, 3, 16, 16)
    sl = slice(rank * 2, rank * 2 + 2)
    net(x[sl], t[sl], lab[sl], context_zero=context_zero).backward(gy[sl])
    got = {k: (None if p.grad is None else p.grad.clone()) for k, p in net.named_parameters()}
    ref = DynamicUNet(compute_dtype=torch.float32, **cfg)
    ref.load_state_dict(net.state_dict())
    ref(x, t, lab, context_zero=context_zero).backward(gy)
    worst, gmax = 0.0, max(float(p.grad.norm()) for p in ref.parameters() if p.grad is not None)
    for k, p in ref.named_parameters():
        if p.grad is None:
            assert got[k] is None, k
            continue
        d = float((got[k] - p.grad / world).norm())
        worst = max(worst, d / (float(p.grad.norm()) / world + 1e-5 * gmax))
    q.put((rank, worst, net.last_reducer.launched))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("context_zero", [True, False], ids=["context_zero", "image_condition"])
def test_dynamic_unet_two_rank_gradient_mean(context_zero):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_dynamic, args=(r, world, port, q, context_zero)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, worst, n_coll in res:
        assert worst < 3e-4, (rank, worst)
        assert n_coll >= 3
