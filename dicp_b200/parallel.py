"""Multi-GPU plumbing: one process per GPU (`torch.distributed`, NCCL on GPUs, gloo in CPU tests).

Cloud pairs are independent, so the hot path shards the batch dimension contiguously over the
ranks and needs no collective.  Learning mode adds exactly one all-reduce per
step over `[loss ‖ flattened parameter gradients]`.

A single huge pair does not shard over pairs; `SplitTargetSearch` splits its TARGETS instead
(SURVEY.md §8 f-3): every rank scans its slice of the target cloud for all queries and one
MIN all-reduce per iteration over packed (distance, index) keys merges the slices — the one
place on the path with a real exchange step."""
from typing import List, Sequence, Tuple

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


# ---------------------------------------------------------------------------------------
# one pair, targets split over the ranks
# ---------------------------------------------------------------------------------------
_IDX_MASK = (1 << 32) - 1


def pack_keys(d2: torch.Tensor, idx: torch.Tensor, offset: int) -> torch.Tensor:
    """(squared distance, global index) as ONE int64 whose integer order is the lexicographic
    order of the pair: the IEEE bit pattern of a non-negative fp32 is monotonic in its value, so
    MIN over the keys of all slices is the nearest neighbour with `torch.argmin`'s first-minimum
    tie-break (nn.py:35-36).  d2 float32 >= 0, idx int32 local to the slice that starts at `offset`."""
    assert d2.dtype == torch.float32
    hi = d2.contiguous().view(torch.int32).to(torch.int64) << 32
    return hi | (idx.to(torch.int64) + int(offset))


def unpack_keys(keys: torch.Tensor):
    """-> (idx int32, d2 float32) of pack_keys."""
    idx = (keys & _IDX_MASK).to(torch.int32)
    d2 = (keys >> 32).to(torch.int32).view(torch.float32)
    return idx, d2


def reduce_min(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return t


def merge_slices(d2: torch.Tensor, idx: torch.Tensor, offset: int, group=None) -> torch.Tensor:
    """Global nearest-neighbour indices from every rank's slice result.  fp32: one MIN all-reduce
    over packed keys; fp64 (96 bits do not fit a key): MIN over the distances, then MIN over the
    indices of the ranks that hold that distance."""
    if d2.dtype == torch.float32:
        return unpack_keys(reduce_min(pack_keys(d2, idx, offset), group))[0]
    best = reduce_min(d2.clone(), group)
    cand = torch.where(d2 == best, idx.to(torch.int64) + int(offset), torch.full_like(idx, 2 ** 31 - 1, dtype=torch.int64))
    return reduce_min(cand, group).to(torch.int32)


class SplitTargetSearch:
    """Correspondence search of `ICP.icp` with the target cloud split over the ranks of `group`.

        icp = ICP(...); icp.nn_search = SplitTargetSearch()        # every rank, same inputs
        res = icp.icp(source, target, T_init, ...)                 # identical results on every rank

    Every rank holds the whole pair (a 130k-point cloud is 3 MB) and runs the cheap per-point
    kernels redundantly; only the O(n m) scan is divided: rank r scans target rows
    [lo_r, hi_r) with the same tiled brute-force kernels (`dicp_nn_pack_*` once per call,
    `dicp_nn_scan_*` per iteration with the loop's own query transform), then `merge_slices`.
    The scan's distance is the exact direct-form fp32/fp64 value whatever the slice, so the
    merged indices — and with them T, pc and the gradients — are bit-identical to the
    single-GPU brute-force call.  Ragged batches (m_valid) are not supported here.

    transport: how the slices are merged.
      'nccl'  one MIN all-reduce per iteration (`merge_slices`);
      'peer'  (fp32) no collective library on the path: a kernel of ours folds every rank's keys into
              the key buffers of ALL ranks with 64-bit system-scope atomic MINs over NVLink peer memory
              (`dicp_nn_push_keys_f32`; buffers from torch's symmetric-memory allocator), one
              cross-GPU barrier, `dicp_nn_unpack_keys_f32`;
      'auto'  'peer' when fp32, more than one rank and the peer buffers can be set up, else 'nccl'."""

    def __init__(self, group=None, rank: int = None, world_size: int = None, transport: str = "auto"):
        assert transport in ("auto", "nccl", "peer")
        self.group = group
        self.transport = transport
        self.used_transport = None            # what the last bind() chose
        self._peer = {}                       # (count, device) -> (keys tensor, handle, per-buffer address arrays)
        on = dist.is_available() and dist.is_initialized()
        self.rank = rank if rank is not None else (dist.get_rank(group) if on else 0)
        self.world_size = world_size if world_size is not None else (dist.get_world_size(group) if on else 1)

    def _peer_buffers(self, count, dev):
        """two key buffers of `count` u64 per rank, mapped into every rank (double-buffered over the
        iterations: iteration k merges in buffer k & 1 and resets the other one)"""
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        key = (count, str(dev))
        if key not in self._peer:
            group = self.group if self.group is not None else dist.group.WORLD
            keys = symm_mem.empty(2 * count, dtype=torch.int64, device=dev)
            hdl = symm_mem.rendezvous(keys, group)
            addr = [(ctypes.c_uint64 * self.world_size)(*[int(hdl.buffer_ptrs[r]) + j * count * 8 for r in range(self.world_size)])
                    for j in (0, 1)]
            self._peer[key] = (keys, hdl, addr)
        keys, hdl, addr = self._peer[key]
        keys.fill_(-1)                        # all ones = +inf key; nobody pushes before the barrier below
        hdl.barrier(channel=0)
        return keys, hdl, addr

    def bind(self, source: torch.Tensor, target: torch.Tensor, m_valid=None):
        from . import _lib
        from ._functions import _workspace
        if m_valid is not None:
            raise ValueError("SplitTargetSearch does not take ragged (padded) target batches")
        lib = _lib.load_library()
        # the scan reads dense (N,n,3) rows: ICP.dICP hands over the `source[:, :, :3]` view of a 6-column cloud
        source = source.contiguous()
        N, n, _ = source.shape
        m = target.shape[1]
        lo, hi = shard_range(m, self.rank, self.world_size)
        if hi <= lo:
            raise ValueError("fewer target points (%d) than ranks (%d)" % (m, self.world_size))
        dev, dt = source.device, source.dtype
        sfx = _lib.suffix(dt)
        st = _lib.stream_ptr(dev)
        part = target[:, lo:hi, :3].contiguous()
        ml = hi - lo
        packed = torch.empty(lib.dicp_nn_packed_bytes(N, ml, source.element_size()), dtype=torch.uint8, device=dev)
        _lib.check(getattr(lib, "dicp_nn_pack_" + sfx)(_lib.ptr(part), N, ml, 3, None, _lib.ptr(packed), st), "dicp_nn_pack")
        ws = _workspace(lib.dicp_nn_workspace_bytes(N, n, ml, source.element_size()), dev)
        idx_l = torch.empty((N, n), dtype=torch.int32, device=dev)
        d2_l = torch.empty((N, n), dtype=dt, device=dev)
        scan = getattr(lib, "dicp_nn_scan_" + sfx)

        def local(state_k):
            """this rank's slice: (squared distance, slice-local index, slice offset) of every query"""
            _lib.check(scan(_lib.ptr(source), _lib.ptr(state_k), _lib.ptr(packed), N, n, ml, _lib.ptr(idx_l), _lib.ptr(d2_l),
                            _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "dicp_nn_scan")
            return d2_l, idx_l, lo

        transport = self.transport
        peer = None
        if transport != "nccl" and dt == torch.float32 and self.world_size > 1 and dist.is_initialized():
            try:
                peer = self._peer_buffers(N * n, dev)
            except Exception:                 # noqa: BLE001 - no peer access / allocator: the collective still works
                if transport == "peer":
                    raise
        elif transport == "peer":
            raise ValueError("transport='peer' needs fp32 clouds and an initialised process group with more than one rank")
        self.used_transport = "peer" if peer is not None else "nccl"

        if peer is None:
            def hook(k, state_k, idx_k):
                idx_k.copy_(merge_slices(*local(state_k), self.group))
        else:
            keys, hdl, addr = peer
            count = N * n

            def hook(k, state_k, idx_k):
                d2, idx, off = local(state_k)
                j = k & 1
                stream = _lib.stream_ptr(dev)
                _lib.check(lib.dicp_nn_push_keys_f32(_lib.ptr(d2), _lib.ptr(idx), off, count, addr[j], self.world_size, self.rank,
                                                     _lib.ptr(keys[(1 - j) * count:]), stream), "dicp_nn_push_keys")
                hdl.barrier(channel=0)        # every rank's keys have landed everywhere
                _lib.check(lib.dicp_nn_unpack_keys_f32(_lib.ptr(keys[j * count:]), count, _lib.ptr(idx_k), None, stream),
                           "dicp_nn_unpack_keys")
        hook.local = local
        hook.keepalive = (source, part, packed, ws)
        return hook
