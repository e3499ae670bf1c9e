"""Worker of tests/test_parallel_gpu.py (one process per GPU, launched by torch.distributed.run): data-parallel gradients of
the CONDITIONAL UNet with label dropout on the CUDA kernels over NCCL against the single-GPU gradients of the concatenated
batch (BASELINE.json configs[2]; reference step DiffusionFreeGuidence/TrainCondition.py:53-63, DDP wrap utils/rotinas.py:618-619).
    mode hdiff : hdiff_b200.parallel.enable_data_parallel (bucketed all-reduce issued from inside backward)
    mode ddp   : torch.nn.parallel.DistributedDataParallel(net, device_ids=[local_rank])  — the reference's own wrap
Second argument: fp32 (check mode: the exchange logic must reproduce the single-GPU gradients to 1e-4) or bf16 (product path).
bf16 tolerance: two bf16 evaluations of this network that are not bit-identical (here: the fp32 partial sums of the GroupNorm
statistics are formed over other tile ranges when the batch per GPU is 2 instead of 4) differ by about 1 % in their weight
gradients — each is about 2 % from the fp32 oracle (tests/tools/debug_dp_split.py measures the same 1 % between a batch of 4 and the
sum of its halves on ONE GPU, no communication involved) — so bf16 is held to 4e-2, the bound of the whole-network tests.
Exit code 0 = every check passed on this rank."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    mode = sys.argv[1]
    fp32 = len(sys.argv) > 2 and sys.argv[2] == "fp32"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from hdiff_b200 import parallel
    from hdiff_b200.DiffusionFreeGuidence.ModelCondition import UNet
    from hdiff_b200.optim import FlatAdamW
    import hdiff_b200.ops as hops
    cfg = dict(T=1000, ch=64, ch_mult=[1, 2, 2], attn=[1], num_res_blocks=1, dropout=0.0)
    res, b = 64, 2
    torch.manual_seed(100 + rank)                          # replicas start different: the wrap must broadcast rank 0's
    cdt = torch.float32 if fp32 else torch.bfloat16
    net = UNet(num_labels=10, compute_dtype=cdt, **cfg).to(dev).train()
    if mode == "hdiff":
        parallel.enable_data_parallel(net, bucket_bytes=1 << 20)      # small buckets: several overlapped collectives
        model = net
    else:
        from torch.nn.parallel import DistributedDataParallel as DDP
        model = DDP(net, device_ids=[local], output_device=local)
    g = torch.Generator(device="cpu").manual_seed(5)
    x = (torch.rand(world * b, 3, res, res, generator=g) * 2 - 1).to(dev)
    t = torch.randint(0, 1000, (world * b,), generator=g).to(dev)
    lab = (torch.randint(0, 10, (world * b,), generator=g) + 1).to(dev)
    lab[:b] = 0                                            # rank 0's label-dropout coin came up (TrainCondition.py:57-58)
    sl = slice(rank * b, (rank + 1) * b)
    before = hops.get().tc_launches
    loss = (model(x[sl], t[sl], lab[sl]) ** 2).sum() / b ** 2.
    loss.backward()
    assert fp32 or hops.get().tc_launches > before, "the tcgen05 kernels must run"
    got = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    if mode == "hdiff":
        assert net.last_reducer.launched >= 4, net.last_reducer.launched
    # single-GPU reference on the concatenated batch, same weights, no process group
    ref = UNet(num_labels=10, compute_dtype=cdt, **cfg).to(dev).train()
    ref.load_state_dict(net.state_dict())
    ((ref(x, t, lab) ** 2).sum() / b ** 2. / world).backward()
    gscale = max(float(p.grad.norm()) for p in ref.parameters())
    bad = []
    for k, p in ref.named_parameters():
        d = float((got[k] - p.grad).norm())
        if d > (2e-4 if fp32 else 4e-2) * float(p.grad.norm()) + (1e-5 if fp32 else 2e-3) * gscale:
            bad.append((k, d / (float(p.grad.norm()) + 1e-30)))
    assert not bad, (rank, bad[:6])
    # one optimizer step keeps the replicas identical
    opt = FlatAdamW(net, lr=1e-3, weight_decay=1e-4, max_grad_norm=1.0)
    opt.step()
    flat = net._get_state().flat
    chk = torch.stack([flat.double().sum(), (flat.double() ** 2).sum()])
    all_chk = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(all_chk, chk)
    assert all(torch.equal(c, all_chk[0]) for c in all_chk), "replicas diverged after the step"
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok ({mode} {'fp32' if fp32 else 'bf16'})", flush=True)


if __name__ == "__main__":
    main()
