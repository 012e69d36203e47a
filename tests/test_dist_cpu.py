"""world_size-2 gloo test (CPU) of the data-parallel gradient exchange used by the training step: bucketed all-reduce
gives every rank the mean gradient, bit-identically on both ranks, and parameters broadcast from rank 0."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path[:0] = [os.path.join(ROOT, "few-shot-transformer-tts_b200"), ROOT]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tts_b200.dist import GradBuckets
    g = torch.Generator().manual_seed(100 + rank)
    shapes = [(300, 7), (5,), (64, 64, 5), (1,), (1000,)]
    params = [torch.nn.Parameter(torch.randn(*s, generator=g)) for s in shapes]
    params[1].requires_grad_(False)
    gb = GradBuckets(params, bucket_bytes=4096)    # several buckets
    assert len(gb.buckets) >= 3
    gb.broadcast_parameters(0)
    for p in params:
        if p.requires_grad:
            p.grad = torch.randn(p.shape, generator=g)
    params[3].grad = None                             # a parameter that received no gradient on this rank
    mine = [None if p.grad is None else p.grad.clone() for p in params]
    gb.allreduce_mean()
    # overlapped mode: two parameter groups, all-reduce launched from autograd hooks while backward is still running
    g2 = torch.Generator().manual_seed(7 + rank)
    wa, wb = torch.nn.Parameter(torch.randn(40, 3, generator=g2)), torch.nn.Parameter(torch.randn(3, 5, generator=g2))
    x = torch.randn(8, 40, generator=g2)
    hb = GradBuckets([[wb], [wa]], bucket_bytes=256)
    hb.attach_hooks()
    ((x @ wa) @ wb).square().sum().backward()
    local = [wa.grad.clone(), wb.grad.clone()]
    hb.finish()
    out[rank] = dict(params=[p.detach().clone() for p in params], before=mine,
                     after=[None if p.grad is None else p.grad.clone() for p in params],
                     hook_local=local, hook_after=[wa.grad.clone(), wb.grad.clone()])
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_gradient_allreduce_world2():
    port = 29000 + os.getpid() % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        r0, r1 = out[0], out[1]
        for a, b in zip(r0["params"][:1] + r0["params"][2:], r1["params"][:1] + r1["params"][2:]):
            assert torch.equal(a, b)                  # broadcast from rank 0
        for i in (0, 2, 4):
            want = (r0["before"][i] + r1["before"][i]) / 2
            assert torch.allclose(r0["after"][i], want, atol=1e-7) and torch.equal(r0["after"][i], r1["after"][i])
        assert r0["after"][1] is None and r0["after"][3] is None
        for i in range(2):   # hook-driven (overlapped) reduction: mean of the two ranks' gradients, identical on both
            want = (r0["hook_local"][i] + r1["hook_local"][i]) / 2
            assert torch.allclose(r0["hook_after"][i], want, atol=1e-5) and torch.equal(r0["hook_after"][i], r1["hook_after"][i])
