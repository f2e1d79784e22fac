"""CPU, world size 2, gloo: the host-side logic of the multi-GPU path (sharding of the pairs,
the single learning-mode all-reduce, result gathering).  No ICP compute: there is no GPU here."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dicp_b200.parallel import allreduce_loss_and_grads, gather_results, shard_pairs, shard_range


def test_shard_ranges_cover_and_balance():
    for n in (1, 2, 7, 256, 2048, 2049):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_pairs([10, 11, 12, 13, 14], 1, 2) == [13, 14]
    assert shard_pairs(torch.arange(6), 0, 4).tolist() == [0, 1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N = 5
        a, b = shard_range(N, rank, world)
        # a toy "learning mode" step: each rank owns its pairs' transforms, theta is replicated
        T = torch.arange(N * 16, dtype=torch.float64).reshape(N, 4, 4)[a:b]
        theta = torch.ones(12, dtype=torch.float64, requires_grad=True)
        bias = torch.zeros(3, dtype=torch.float64, requires_grad=True)
        loss = (T[:, :3, :].reshape(-1, 12) * theta).sum() + (bias * float(rank + 1)).sum()
        loss.backward()
        total = allreduce_loss_and_grads(loss, [theta, bias])
        full = gather_results(T, N)
        ret[rank] = (float(total), theta.grad.clone(), bias.grad.clone(), full.clone())
    finally:
        dist.destroy_process_group()


def test_allreduce_and_gather_world2():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        Tall = torch.arange(5 * 16, dtype=torch.float64).reshape(5, 4, 4)
        want_loss = float(Tall[:, :3, :].sum())
        want_theta = Tall[:, :3, :].reshape(-1, 12).sum(0)
        for r in range(world):
            total, g_theta, g_bias, full = ret[r]
            assert abs(total - want_loss) < 1e-9
            assert torch.allclose(g_theta, want_theta)
            assert torch.allclose(g_bias, torch.full((3,), 3.0, dtype=torch.float64))   # 1 + 2
            assert torch.equal(full, Tall)
