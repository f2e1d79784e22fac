"""BASELINE.json config 4 (one 130k x 130k pair, pt2pl, forward + backward, 10 iterations) with the
TARGETS split over the GPUs of one box (dicp_b200.parallel.SplitTargetSearch):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/run_c4_split.py [auto|nccl|peer]

Rank 0 also runs the plain single-GPU brute-force call and checks that the split run reproduces
it bit for bit; timing: CUDA events, barrier + synchronize on both sides, max over ranks."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dICP.ICP import ICP  # noqa: E402
from dicp_b200.parallel import SplitTargetSearch  # noqa: E402
from dicp_b200.synthetic import surface_pairs as make_wavy_pair  # noqa: E402

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

src, tgt = make_wavy_pair(1, 130000, 130000, seed=4000, dtype=torch.float32, extent=60.0)
src, tgt = src.to(dev), tgt.to(dev)
T0 = torch.eye(4, device=dev)[None]
theta = torch.arange(1, 13, dtype=torch.float32, device=dev)


def step(search):
    icp = ICP(icp_type="pt2pl", differentiable=True, max_iterations=10, tolerance=1e-12)
    icp.const_iter = True
    icp.nn_search = search
    s, t = src.clone().requires_grad_(True), tgt.clone().requires_grad_(True)
    res = icp.icp(s, t, T0, trim_dist=2.0, loss_fn={"name": "cauchy", "metric": 0.5}, dim=3)
    (res["T"][:, :3, :].reshape(-1, 12) * theta).sum().backward()
    return res["T"].detach(), s.grad


def timed(search, reps=5, warm=2):
    for _ in range(warm):
        out = step(search)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = step(search)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / reps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms), out


transport = sys.argv[1] if len(sys.argv) > 1 else "auto"
search = SplitTargetSearch(transport=transport)
ms_split, (T_split, g_split) = timed(search)
line = {"transport": search.used_transport, "config": "C4: 1 pair x 130000 x 130000, pt2pl, cauchy(0.5)+trim(2.0), fp32, 10 it, forward + backward",
        "n_gpus": world, "split_targets_ms": ms_split, "pairs_per_s": 1e3 / ms_split,
        "nn_tflops_algorithmic": 8.0 * 130000.0 ** 2 * 10 / (ms_split * 1e-3) / 1e12}
if rank == 0:
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step("brute")
    a.record()
    for _ in range(3):
        T_one, g_one = step("brute")
    b.record()
    torch.cuda.synchronize()
    line["single_gpu_brute_ms"] = a.elapsed_time(b) / 3
    line["bit_identical_T"] = bool(torch.equal(T_one, T_split))
    line["bit_identical_grad_source"] = bool(torch.equal(g_one, g_split))
    print(json.dumps(line))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
