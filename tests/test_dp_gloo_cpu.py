"""N>1 host logic on CPU: two gloo ranks, flat gradient buffers, ONE all-reduce per buffer, identical
parameters after an SGD step; plus the no_sync()/re-attach behaviour of FlatParams."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from navillm_b200.llama import FlatParams, allreduce_flat_grads
    torch.manual_seed(0)                                     # identical replicas
    lin_a, lin_b = nn.Linear(64, 128, bias=False), nn.Linear(128, 64, bias=False)
    emb = nn.Embedding(10, 64)
    bf = FlatParams([p for m in (lin_a, lin_b) for p in m.parameters()], torch.device("cpu"))
    f32 = FlatParams(list(emb.parameters()), torch.device("cpu"))
    # q/k-style adjacency: the two [128,64]/[64,128] weights are consecutive blocks of one buffer
    assert lin_a.weight.data_ptr() == bf.flat.data_ptr()
    assert lin_b.weight.data_ptr() == bf.flat.data_ptr() + lin_a.weight.numel() * 4
    g = torch.Generator().manual_seed(100 + rank)            # different data per rank
    with torch.no_grad():
        for p in list(bf.params) + list(f32.params):
            p.grad.copy_(torch.randn(p.shape, generator=g))
    local = [p.grad.clone() for p in list(bf.params) + list(f32.params)]
    n = allreduce_flat_grads([bf, f32, None], average=True)
    assert n == 2                                            # one collective per flat buffer, not per parameter
    gathered = [None] * world
    dist.all_gather_object(gathered, [t.numpy() for t in local])
    for i, p in enumerate(list(bf.params) + list(f32.params)):
        mean = sum(torch.from_numpy(gathered[r][i]) for r in range(world)) / world
        assert torch.allclose(p.grad, mean, atol=1e-6), f"rank {rank} param {i}"
    # optimizer on the views, then zero_grad(set_to_none=True) + re-attach
    opt = torch.optim.SGD(list(bf.params) + list(f32.params), lr=0.1)
    opt.step()
    opt.zero_grad(set_to_none=True)
    assert bf.params[0].grad is None and not bf.intact()
    bf.reattach_grads()
    assert bf.intact() and float(bf.flat_grad.abs().sum()) == 0.0
    ws = [p.detach().clone().numpy() for p in bf.params]
    gathered = [None] * world
    dist.all_gather_object(gathered, ws)
    for r in range(1, world):
        for a, b in zip(gathered[0], gathered[r]):
            assert (a == b).all(), "replicas diverged after the reduced step"
    q.put((rank, "ok"))
    dist.destroy_process_group()


def test_flat_grad_allreduce_two_ranks_gloo():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0, f"rank failed with exit code {p.exitcode}"
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, "ok"), (1, "ok")]


def test_allreduce_is_noop_without_process_group():
    from navillm_b200.llama import FlatParams, allreduce_flat_grads
    lin = nn.Linear(8, 8)
    f = FlatParams(list(lin.parameters()), torch.device("cpu"))
    assert allreduce_flat_grads([f]) == 0
