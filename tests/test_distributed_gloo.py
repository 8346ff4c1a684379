"""world_size-2 gloo tests (CPU) of the multi-GPU choreography in mtdgan_b200/distributed.py: reduce-scatter ->
partial Gram -> all-reduce -> replicated solve -> combine on the shard -> all-gather must equal single-process
PCGrad on the rank-averaged gradients (SURVEY §8e equality target).  The arithmetic back-end is injected
(torch stand-ins here, CUDA kernels in production), so this exercises exactly the code path N > 1 runs."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mtdgan_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torch_gram(shards):
    g = torch.zeros(16, dtype=torch.float64)
    for a in range(len(shards)):
        for b in range(a, len(shards)):
            g[a * 4 + b] = shards[a].double() @ shards[b].double()
    return g


def _torch_solve_combine(shards, gram, orders, mean, scale):
    T = len(shards)
    G = np.zeros((T, T))
    for a in range(T):
        for b in range(a, T):
            G[a, b] = G[b, a] = float(gram[a * 4 + b])
    C = O.pcgrad_coefficients(G, [list(map(int, o)) for o in np.asarray(orders).reshape(T, T)])
    w = C.sum(0) / (T if mean else 1)
    return sum(float(w[k]) * scale * shards[k] for k in range(T))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from mtdgan_b200 import distributed as mdist
    mdist.init("gloo")
    assert mdist.active() and mdist.world() == world and mdist.rank() == rank
    g = torch.Generator().manual_seed(100 + rank)
    shapes = [(7,), (3, 5), (64, 9), (1,), (33, 2, 2)]              # sum = 373: not divisible by 2 -> padded shards
    base = [torch.randn(s, generator=g) for s in shapes]
    grads = [base, [-0.6 * b + 0.2 * torch.randn(b.shape, generator=g) for b in base],
             [1e-4 * torch.randn(b.shape, generator=g) for b in base]]
    random.seed(5)                                                   # every rank draws the same visit orders
    from mtdgan_b200.weight_methods import draw_visit_orders
    orders = draw_visit_orders(3)
    merged = mdist.pcgrad_sharded(grads, torch.tensor(orders).reshape(-1), False, _torch_gram, _torch_solve_combine)
    # plain gradient averaging paths
    ps = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
    for p, b in zip(ps, base):
        p.grad = b.clone()
    ps[3].grad = None
    mdist.allreduce_mean_grads(ps)
    avg = mdist.allreduce_mean_list([b.clone() for b in base])
    if rank == 0:
        torch.save({"merged": merged, "orders": orders, "param_grads": [p.grad for p in ps], "avg": avg}, out)
    # every rank must hold the identical merged gradient
    flat = torch.cat([m.flatten() for m in merged])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(flat, ref)
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_pcgrad_sharded_equals_single_process(tmp_path):
    world, port, out = 2, _free_port(), str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    res = torch.load(out, weights_only=False)
    # single-process reference: average the per-rank task gradients, then the oracle's vector-space PCGrad
    shapes = [(7,), (3, 5), (64, 9), (1,), (33, 2, 2)]
    per_rank = []
    for rank in range(world):
        g = torch.Generator().manual_seed(100 + rank)
        base = [torch.randn(s, generator=g) for s in shapes]
        per_rank.append([base, [-0.6 * b + 0.2 * torch.randn(b.shape, generator=g) for b in base],
                         [1e-4 * torch.randn(b.shape, generator=g) for b in base]])
    avg = [tuple(sum(per_rank[r][k][i] for r in range(world)) / world for i in range(len(shapes))) for k in range(3)]
    random.seed(5)
    want = O.pcgrad_project_lists(list(avg), "sum")
    for got, w in zip(res["merged"], want):
        assert got.shape == w.shape
        assert torch.allclose(got, w, rtol=1e-5, atol=1e-7)
    for i, (pg, s) in enumerate(zip(res["param_grads"], shapes)):
        if i == 3:
            assert pg is None
        else:
            assert torch.allclose(pg, (per_rank[0][0][i] + per_rank[1][0][i]) / 2, rtol=1e-6, atol=1e-7)
    for i, a in enumerate(res["avg"]):
        assert torch.allclose(a, (per_rank[0][0][i] + per_rank[1][0][i]) / 2, rtol=1e-6, atol=1e-7)


def test_shard_bounds_and_flatten():
    from mtdgan_b200 import distributed as mdist
    assert mdist.shard_bounds(10, 4, 0) == (3, 0, 3) and mdist.shard_bounds(10, 4, 3) == (3, 9, 10)
    assert mdist.shard_bounds(8, 2, 1) == (4, 4, 8)
    ts = [torch.arange(6.).reshape(2, 3), torch.arange(5.)]
    flat = mdist.flatten(ts, pad_to=4)
    assert flat.numel() == 12 and flat[11] == 0
    back = mdist.unflatten(flat, ts)
    assert all(torch.equal(a, b) for a, b in zip(back, ts))
    assert not mdist.active() and mdist.world() == 1 and mdist.rank() == 0
