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


# ---------------------------------------------------------------------------------------
# one pair with its targets split over the ranks: the merge of the slice results
# ---------------------------------------------------------------------------------------
def _split_case(dtype, seed=5):
    g = torch.Generator().manual_seed(seed)
    N, n, m = 2, 300, 257                              # m not divisible by the world size
    x = torch.randn((N, n, 3), generator=g, dtype=dtype)
    y = torch.randn((N, m, 3), generator=g, dtype=dtype)
    y[:, 200] = y[:, 17]                               # an exact duplicate: the tie must go to the lower index
    x[:, 0] = y[:, 17]
    return x, y


def _direct_d2(x, y):
    d = x[:, :, None, :] - y[:, None, :, :]
    return (d * d).sum(-1)


def test_pack_keys_order_and_roundtrip():
    from dicp_b200.parallel import pack_keys, unpack_keys
    d2 = torch.tensor([0.0, 1e-30, 1.0, 1.0, 3.5e38], dtype=torch.float32)
    idx = torch.tensor([7, 0, 9, 3, 2 ** 31 - 1], dtype=torch.int32)
    k = pack_keys(d2, idx, 0)
    assert torch.equal(torch.argsort(k), torch.tensor([0, 1, 3, 2, 4]))     # by distance, then by index
    i2, d2b = unpack_keys(k)
    assert torch.equal(i2, idx) and torch.equal(d2b, d2)
    assert torch.equal(unpack_keys(pack_keys(d2[:2], idx[:2], 1000))[0], idx[:2] + 1000)


def _split_worker(rank, world, port, ret):
    from dicp_b200.parallel import merge_slices
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = {}
        for dtype in (torch.float32, torch.float64):
            x, y = _split_case(dtype)
            lo, hi = shard_range(y.shape[1], rank, world)
            d2 = _direct_d2(x, y[:, lo:hi])                              # what the slice scan returns ...
            best, idx = d2.min(-1)
            out[str(dtype)] = merge_slices(best, idx.to(torch.int32), lo).clone()
        ret[rank] = out
    finally:
        dist.destroy_process_group()


def test_split_target_merge_world2():
    world = 2
    port = 31500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_split_worker, args=(world, port, ret), nprocs=world, join=True)
        for dtype in (torch.float32, torch.float64):
            x, y = _split_case(dtype)
            want = _direct_d2(x, y).argmin(-1).to(torch.int32)           # first minimum over the whole cloud
            assert int(want[0, 0]) == 17
            for r in range(world):
                assert torch.equal(ret[r][str(dtype)], want)
