"""Time BASELINE.json configs 2 (64 x 4096 pt2pt forward) and 4 (130k x 130k single pair fwd+bwd)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dICP.ICP import ICP  # noqa: E402
from dicp_b200.synthetic import surface_pairs as make_wavy_pair  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
# config 1: the reference's pytest fixture, fp64, pt2pl + Cauchy, dim 2, tol 1e-10 (7 iterations), forward + backward
import numpy as np
fx = np.load(os.path.join(ROOT, "tests", "golden", "fixture_clouds.npz"))
s1 = torch.tensor(fx["scan"][:, :3]).cuda().requires_grad_(True)
t1 = torch.tensor(fx["map"]).cuda().requires_grad_(True)
icp1 = ICP(icp_type="pt2pl", max_iterations=100, tolerance=1e-10)


def step1():
    s1.grad = t1.grad = None
    r = icp1.icp(s1, t1, torch.eye(4, dtype=torch.float64, device="cuda"), trim_dist=5.0, loss_fn={"name": "cauchy", "metric": 0.5}, dim=2)
    r["T"].sum().backward()
    return r


ms = timed(step1, reps=20, warm=3)
out["config1_fixture_65pts_fp64_fwd_bwd"] = {"ms": ms, "iterations": int(step1()["stats"]["iterations"][0])}

src, tgt = make_wavy_pair(64, 4096, 4096, seed=2000, dtype=torch.float32, with_normals=False)
src, tgt, T0 = src.cuda(), tgt.cuda(), torch.eye(4, device="cuda").repeat(64, 1, 1)
icp = ICP(icp_type="pt2pt", max_iterations=10, tolerance=1e-12)
icp.const_iter = True
with torch.no_grad():
    ms = timed(lambda: icp.icp(src, tgt, T0, dim=3))
out["config2_pt2pt_64x4096_fwd"] = {"ms": ms, "pairs_per_s": 64 / ms * 1e3}

# the same clouds through the Kabsch / SVD loop (functionality.svd, ICP.py:533-591): 10 iterations, forward and forward + backward
icp_svd = ICP(icp_type="pt2pt", max_iterations=10, tolerance=0.0)
icp_svd.use_svd = True
icp_svd.check_every = 10
with torch.no_grad():
    ms = timed(lambda: icp_svd.icp(src, tgt, T0))
out["config2_shape_kabsch_svd_64x4096_fwd"] = {"ms": ms, "pairs_per_s": 64 / ms * 1e3}
src_g = src.clone().requires_grad_(True)


def step_svd():
    src_g.grad = None
    icp_svd.icp(src_g, tgt, T0)["T"].sum().backward()


ms = timed(step_svd)
out["config2_shape_kabsch_svd_64x4096_fwd_bwd"] = {"ms": ms, "pairs_per_s": 64 / ms * 1e3}

src, tgt = make_wavy_pair(1, 130000, 130000, seed=4000, dtype=torch.float32, extent=60.0)
src, tgt, T0 = src.cuda().requires_grad_(True), tgt.cuda().requires_grad_(True), torch.eye(4, device="cuda").unsqueeze(0)
icp = ICP(icp_type="pt2pl", max_iterations=10, tolerance=1e-12)
icp.const_iter = True


def step4():
    src.grad = tgt.grad = None
    res = icp.icp(src, tgt, T0, trim_dist=2.0, loss_fn={"name": "cauchy", "metric": 0.5}, dim=3)
    res["T"].sum().backward()


ms = timed(step4, reps=3, warm=1)
out["config4_pt2pl_1x130k_fwd_bwd"] = {"ms": ms, "pairs_per_s": 1 / ms * 1e3,
                                       "nn_tflops_algorithmic": 8.0 * 130000.0 ** 2 * 10 / (ms * 1e-3) / 1e12}
print(json.dumps(out))
