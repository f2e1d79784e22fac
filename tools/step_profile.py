"""One C3 step (forward + backward) under torch.profiler: where does the step time go?"""
import os
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dICP.ICP import ICP  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
MODE = sys.argv[2] if len(sys.argv) > 2 else "brute"
dev = torch.device("cuda")
src, tgt = bench.make_batch(N, 0)
src = src.to(dev).requires_grad_(True)
tgt = tgt.to(dev).requires_grad_(True)
T0 = torch.eye(4, device=dev).repeat(N, 1, 1)
icp = ICP(icp_type="pt2pl", differentiable=True, max_iterations=10, tolerance=1e-12)
icp.const_iter = True
icp.nn_search = MODE
theta = torch.arange(1, 13, dtype=torch.float32, device=dev)


def step():
    src.grad = tgt.grad = None
    t0 = time.perf_counter()
    res = icp.icp(src, tgt, T0, trim_dist=2.0, loss_fn={"name": "cauchy", "metric": 0.5}, dim=3)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    (res["T"][:, :3, :].reshape(-1, 12) * theta).sum().backward()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3


for _ in range(3):
    print("forward %.2f ms, backward %.2f ms" % step())
if os.environ.get("NO_TORCH_PROFILER"):      # e.g. when the process runs under ncu
    sys.exit(0)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
