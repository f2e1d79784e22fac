"""Step times (forward + backward, 10 constant iterations, pt2pl + Cauchy + trim, fp32) for batch shapes smaller than the
benchmark's: where the loop is launch- / latency-bound rather than throughput-bound.  CUDA events around `reps` steps."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dICP.ICP import ICP  # noqa: E402
from dicp_b200.synthetic import surface_pairs  # noqa: E402

dev = torch.device("cuda")
theta = torch.arange(1, 13, dtype=torch.float32, device=dev)
for N, n in ((1, 4096), (1, 16384), (1, 65536), (8, 4096), (8, 16384), (32, 8192), (64, 4096)):
    src, tgt = surface_pairs(N, n, n, seed=7000 + N + n, dtype=torch.float32, with_normals=True)
    src, tgt = src.to(dev).requires_grad_(True), tgt.to(dev).requires_grad_(True)
    T0 = torch.eye(4, device=dev).repeat(N, 1, 1)
    row = []
    for mode, sf, const in (("brute", False, True), ("brute", True, True), ("grid", False, True), ("grid", True, True), ("grid", False, False)):
        icp = ICP(icp_type="pt2pl", differentiable=True, max_iterations=10, tolerance=1e-12 if const else 0.0)
        icp.const_iter, icp.nn_search, icp.sync_free = const, mode, sf

        def step():
            src.grad = tgt.grad = None
            res = icp.icp(src, tgt, T0, trim_dist=2.0, loss_fn={"name": "cauchy", "metric": 0.5}, dim=3)
            (res["T"][:, :3, :].reshape(-1, 12) * theta).sum().backward()

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        a.record()
        for _ in range(reps):
            step()
        b.record()
        torch.cuda.synchronize()
        row.append("%s%s%s %.3f ms" % (mode, " sync-free" if sf else "", "" if const else " (host check every iteration)", a.elapsed_time(b) / reps))
    # the sync-free grid step captured once in a CUDA graph and replayed (no launch overhead on the host)
    icp = ICP(icp_type="pt2pl", differentiable=True, max_iterations=10, tolerance=1e-12)
    icp.const_iter, icp.nn_search, icp.sync_free = True, "grid", True

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
    for _ in range(20):
        graph.replay()
    b.record()
    torch.cuda.synchronize()
    row.append("grid as a CUDA graph replay %.3f ms" % (a.elapsed_time(b) / 20))
    del graph, outs
    print("N=%d n=m=%d: " % (N, n) + " | ".join(row), flush=True)
