"""CPU: the hand-derived reverse pass (oracle/adjoint.py — the executable spec of the CUDA
backward kernels) against torch autograd through the oracle loop, correspondences pinned."""
import itertools

import numpy as np
import pytest
import torch

from oracle.dicp_oracle import IcpSettings, icp_forward, make_wavy_pair
from oracle.adjoint import icp_forward_saved, icp_backward
from helpers import rel_err

CASES = list(itertools.product(["pt2pl", "pt2pt"], [2, 3], [True, False],
                               [(None, None), ("cauchy", 2.0), ("huber", 2.0), ("huber", None)]))


@pytest.mark.parametrize("icp_type,dim,diff,lt", CASES)
def test_adjoint_matches_autograd(icp_type, dim, diff, lt):
    lname, trim = lt
    N, n, m, K = 2, 60, 70, 4
    src, tgt = make_wavy_pair(N, n, m, seed=21, dtype=torch.float64, noise=0.05)
    if icp_type == "pt2pt":
        tgt = tgt[:, :, :3].contiguous()
    gen = torch.Generator().manual_seed(3)
    w0 = 0.3 + torch.rand((N, n), generator=gen, dtype=torch.float64)
    T0 = torch.eye(4, dtype=torch.float64).repeat(N, 1, 1)
    T0[:, :3, 3] = torch.tensor([0.02, -0.03, 0.01], dtype=torch.float64)
    cfg = IcpSettings(icp_type=icp_type, max_iterations=K, tolerance=1e-14, differentiable=diff, const_iter=True,
                      trim_dist=trim, loss_name=lname, loss_metric=0.15, dim=dim)
    # autograd through the oracle
    s, t, T, w = (v.clone().requires_grad_(True) for v in (src, tgt, T0, w0))
    idx_log = []
    wfull = w.repeat_interleave(3, dim=1) if icp_type == "pt2pt" else w
    res = icp_forward(s, t, T, wfull, cfg, record_idx=idx_log)
    gT = torch.randn(res["T"].shape, generator=gen, dtype=torch.float64)
    gT[:, 3, :] = 0
    gpc = torch.randn(res["pc"].shape, generator=gen, dtype=torch.float64) * 0.1
    ((res["T"] * gT).sum() + (res["pc"] * gpc).sum()).backward()
    # explicit path (inputs pre-processed like ICP.py:107-116 for dim 2)
    s2, t2 = src.clone(), tgt.clone()
    if dim == 2:
        s2[:, :, 2] = 0
        t2[:, :, 2] = 0
        if icp_type == "pt2pl":
            t2[:, :, 5] = 0
    fw = icp_forward_saved(s2, t2, T0, w0, cfg, idx_log)
    assert rel_err(fw["T"], res["T"].detach()) < 1e-12
    assert rel_err(fw["pc"], res["pc"].detach()) < 1e-12
    g = icp_backward(s2, t2, w0, cfg, idx_log, fw["saved"], gT, gpc)
    gs, gt = g["source"].clone(), g["target"].clone()
    if dim == 2:
        gs[:, :, 2] = 0
        gt[:, :, 2] = 0
        if icp_type == "pt2pl":
            gt[:, :, 5] = 0
    assert rel_err(gs, s.grad) < 1e-9
    assert rel_err(gt, t.grad) < 1e-9
    assert rel_err(g["w0"], w.grad) < 1e-9
    assert rel_err(g["T_init"], T.grad) < 1e-9
