"""C3 step (forward + backward, device-resident inputs) with and without ICP.sync_free, for both searches:
what the host synchronisation at the end of the forward call costs.  CUDA events around `reps` steps."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dICP.ICP import ICP  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda")
src, tgt = bench.make_batch(N, 0)
src = src.to(dev).requires_grad_(True)
tgt = tgt.to(dev).requires_grad_(True)
T0 = torch.eye(4, device=dev).repeat(N, 1, 1)
theta = torch.arange(1, 13, dtype=torch.float32, device=dev)

for mode in ("grid", "brute"):
    for sf in (False, True, False, True):
        icp = ICP(icp_type="pt2pl", differentiable=True, max_iterations=10, tolerance=1e-12)
        icp.const_iter, icp.nn_search, icp.sync_free = True, mode, sf

        def step():
            src.grad = tgt.grad = None
            res = icp.icp(src, tgt, T0, trim_dist=2.0, loss_fn={"name": "cauchy", "metric": 0.5}, dim=3)
            (res["T"][:, :3, :].reshape(-1, 12) * theta).sum().backward()

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            step()
        b.record()
        torch.cuda.synchronize()
        print("%-5s sync_free=%-5s  %.3f ms per step  (%.0f pairs/s)" % (mode, sf, a.elapsed_time(b) / reps, N * reps / a.elapsed_time(b) * 1e3), flush=True)

# the same step captured once in a CUDA graph (possible because the sync-free call never waits for the GPU) and replayed
for mode in ("grid", "brute"):
    icp = ICP(icp_type="pt2pl", differentiable=True, max_iterations=10, tolerance=1e-12)
    icp.const_iter, icp.nn_search, icp.sync_free = True, mode, True

    def gstep():
        res = icp.icp(src, tgt, T0, trim_dist=2.0, loss_fn={"name": "cauchy", "metric": 0.5}, dim=3)
        return torch.autograd.grad((res["T"][:, :3, :].reshape(-1, 12) * theta).sum(), [src, tgt])

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            gstep()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        outs = gstep()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        graph.replay()
    b.record()
    torch.cuda.synchronize()
    print("%-5s CUDA graph replay  %.3f ms per step  (%.0f pairs/s)" % (mode, a.elapsed_time(b) / reps, N * reps / a.elapsed_time(b) * 1e3), flush=True)
