"""Multi-GPU plumbing: one process per GPU (`torch.distributed`, NCCL on GPUs, gloo in CPU tests).

Cloud pairs are independent, so the hot path shards the batch dimension contiguously over the
ranks and needs no collective.  Learning mode adds exactly one all-reduce per
step over `[loss ‖ flattened parameter gradients]`."""
from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_pairs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) of the pairs rank `rank` owns (the first
    `n_pairs % world_size` ranks get one extra pair)."""
    assert 0 <= rank < world_size
    base, extra = divmod(n_pairs, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_pairs(batch, rank: int, world_size: int):
    """Slice a batched input — a tensor (N,...) or a list of per-pair tensors — to this rank's pairs."""
    if batch is None:
        return None
    start, stop = shard_range(len(batch), rank, world_size)
    return batch[start:stop]


def allreduce_loss_and_grads(loss: torch.Tensor, params: Sequence[torch.Tensor], group=None) -> torch.Tensor:
    """SUM-all-reduce the scalar loss and every `p.grad` in ONE collective (a flat buffer of
    1 + sum(numel) elements), write the reduced gradients back, return the reduced loss."""
    grads: List[torch.Tensor] = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
    flat = torch.cat([loss.detach().reshape(1).to(grads[0].dtype if grads else loss.dtype)] + [g.reshape(-1) for g in grads])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 1
    for p, g in zip(params, grads):
        n = g.numel()
        p.grad = flat[off:off + n].reshape(g.shape).clone()
        off += n
    return flat[0]


def gather_results(local: torch.Tensor, n_pairs: int, group=None) -> torch.Tensor:
    """All-gather per-pair results (e.g. T (n_local,4,4)) back into batch order; shards may differ by one pair."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(n_pairs, r, world) for r in range(world)]
    cap = max(b - a for a, b in sizes)
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: b - a] for o, (a, b) in zip(out, sizes)], dim=0)
