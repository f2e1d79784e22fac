"""BASELINE config 5 on ONE GPU: 2048 pairs x 16k x 16k, forward + backward, both search modes."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dICP.ICP import ICP  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
src, tgt = bench.make_batch(256, 0)
rep = N // 256
src = src.cuda().repeat(rep, 1, 1).requires_grad_(True)
tgt = tgt.cuda().repeat(rep, 1, 1).requires_grad_(True)
T0 = torch.eye(4, device="cuda").repeat(N, 1, 1)
out = {"pairs": N}
ref_T = None
for mode in ("grid", "brute"):
    icp = ICP(icp_type="pt2pl", max_iterations=10, tolerance=1e-12)
    icp.const_iter = True
    icp.nn_search = mode
    times = []
    for it in range(2 if mode == "brute" else 3):
        src.grad = tgt.grad = None
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = icp.icp(src, tgt, T0, trim_dist=2.0, loss_fn={"name": "cauchy", "metric": 0.5}, dim=3)
        res["T"].sum().backward()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    out[mode] = {"ms_per_step": min(times), "pairs_per_s": N / min(times) * 1e3}
    if ref_T is None:
        ref_T = res["T"].detach().clone()
    else:
        out["T_identical_across_modes"] = bool(torch.equal(ref_T, res["T"].detach()))
    # every replica of the 256-pair block must give the same transforms (determinism across the batch)
    out[mode]["replicas_identical"] = bool(torch.equal(res["T"][:256], res["T"][-256:]))
out["max_memory_gb"] = torch.cuda.max_memory_allocated() / 1e9
print(json.dumps(out))
