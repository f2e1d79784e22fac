"""Live differential test (`-m gpu`): the CUDA path against the UNMODIFIED reference itself (utiasASRL/dICP, installed by
baseline/install_ref.py into git-ignored baseline/_ref, which travels to the GPU box) on seeded random problems -- both
ICP types, dim 2 / 3, every weight option, dense and ragged list batches, user weights, non-identity initial guesses --
in fp64, where nearest neighbours are tie-free and everything must agree to rounding: T, pc, deltas, weights, costs, stats
and the autograd gradients of the reference (dICP/ICP.py:46-303, dICP/nn.py:11-40, dICP/loss.py).  Skipped when
baseline/_ref is absent."""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def ref():
    from baseline import ref_arm
    r = ref_arm.load_reference()
    if r is None:
        pytest.skip("baseline/_ref is absent (python baseline/install_ref.py needs /root/reference)")
    return r, ref_arm


def _clouds(gen, N, n, m, dim, ragged):
    """target: points on a bumpy surface with unit normals; source: a subset moved by a small rigid motion + noise"""
    out_s, out_t = [], []
    for b in range(N):
        mb = m - (int(torch.randint(0, m // 3, (1,), generator=gen)) if ragged else 0)
        nb = n - (int(torch.randint(0, n // 3, (1,), generator=gen)) if ragged else 0)
        xy = (torch.rand((mb, 2), generator=gen, dtype=torch.float64) - 0.5) * 8.0
        z = 0.4 * torch.sin(0.9 * xy[:, 0]) + 0.3 * torch.cos(1.3 * xy[:, 1])
        nrm = torch.stack((-0.36 * torch.cos(0.9 * xy[:, 0]), 0.39 * torch.sin(1.3 * xy[:, 1]), torch.ones(mb, dtype=torch.float64)), dim=1)
        nrm = nrm / nrm.norm(dim=1, keepdim=True)
        pts = torch.cat((xy, z[:, None]), dim=1)
        pick = torch.randperm(mb, generator=gen)[:min(nb, mb)]
        ang = 0.06 * (torch.rand(1, generator=gen, dtype=torch.float64) - 0.5)
        c, s_ = torch.cos(ang), torch.sin(ang)
        Rz = torch.tensor([[float(c), -float(s_), 0.0], [float(s_), float(c), 0.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
        shift = 0.1 * (torch.rand(3, generator=gen, dtype=torch.float64) - 0.5)
        if dim == 2:
            shift[2] = 0.0
        src = (pts[pick] + 0.005 * torch.randn((len(pick), 3), generator=gen, dtype=torch.float64)) @ Rz.T + shift
        out_s.append(src)
        out_t.append(torch.cat((pts, nrm), dim=1))
    return out_s, out_t


CASES = []
for seed, (icp_type, diff, loss_name, metric, trim, dim, ragged, use_w, const_iter) in enumerate([
        ("pt2pl", True, "cauchy", 0.5, 2.0, 3, False, False, True),
        ("pt2pl", True, "huber", 0.3, 1.0, 3, False, True, True),
        ("pt2pl", False, "huber", 0.2, 1.5, 3, False, False, False),
        ("pt2pl", False, "cauchy", 0.4, None, 3, True, False, False),
        ("pt2pl", True, None, 1.0, 1.0, 3, True, True, True),
        ("pt2pl", True, "cauchy", 0.5, 3.0, 2, False, False, False),
        ("pt2pt", True, "cauchy", 0.5, 2.0, 3, False, False, True),
        ("pt2pt", True, "huber", 0.3, None, 3, True, True, True),
        ("pt2pt", False, "huber", 0.25, 1.0, 3, False, False, False),
        ("pt2pt", False, None, 1.0, None, 3, True, False, False),
        ("pt2pt", True, "cauchy", 0.6, 2.5, 2, True, False, True),
        ("pt2pt", True, None, 1.0, None, 2, False, True, False)]):
    CASES.append((seed, icp_type, diff, loss_name, metric, trim, dim, ragged, use_w, const_iter))


@pytest.mark.parametrize("case", CASES, ids=["%02d_%s_%s_%s_dim%d%s%s" % (c[0], c[1], "soft" if c[2] else "hard", c[3], c[6],
                                                                          "_ragged" if c[7] else "", "_w" if c[8] else "") for c in CASES])
def test_cuda_path_equals_the_live_reference(ref, case):
    from dICP.ICP import ICP
    r, ref_arm = ref
    seed, icp_type, diff, loss_name, metric, trim, dim, ragged, use_w, const_iter = case
    gen = torch.Generator().manual_seed(9000 + seed)
    N, n, m, K = 3, 150, 220, 8
    src_l, tgt_l = _clouds(gen, N, n, m, dim, ragged)
    if icp_type == "pt2pt" and seed % 2 == 0:
        tgt_l = [t[:, :3].contiguous() for t in tgt_l]                 # 3-column targets are allowed for pt2pt
    T0 = torch.eye(4, dtype=torch.float64).repeat(N, 1, 1)
    T0[:, :2, 3] += 0.02 * (torch.rand((N, 2), generator=gen, dtype=torch.float64) - 0.5)
    w_l = [0.3 + 0.7 * torch.rand(s.shape[0], generator=gen, dtype=torch.float64) for s in src_l] if use_w else None
    loss_fn = {"name": loss_name, "metric": metric} if loss_name else None
    tol = 1e-9 if const_iter else 1e-8

    def run(make_icp, dev):
        leaves_s = [s.clone().to(dev).requires_grad_(True) for s in src_l]
        leaves_t = [t.clone().to(dev).requires_grad_(True) for t in tgt_l]
        leaves_w = [w.clone().to(dev).requires_grad_(True) for w in w_l] if use_w else None
        T_in = T0.clone().to(dev).requires_grad_(True)
        icp = make_icp()
        if ragged:
            res = icp.icp(leaves_s, leaves_t, T_in, weight=leaves_w, trim_dist=trim, loss_fn=loss_fn, dim=dim)
        else:
            res = icp.icp(torch.stack(leaves_s), torch.stack(leaves_t), T_in, weight=torch.stack(leaves_w) if use_w else None,
                          trim_dist=trim, loss_fn=loss_fn, dim=dim)
        wT = torch.arange(1, 13, dtype=torch.float64, device=dev).reshape(3, 4)
        ((res["T"][:, :3, :] * wT).sum() + 0.01 * res["pc"].sum()).backward()
        grads = {"T_init": T_in.grad.cpu()}
        for i in range(N):
            grads["s%d" % i] = leaves_s[i].grad.cpu()
            grads["t%d" % i] = leaves_t[i].grad.cpu()
            if use_w:
                grads["w%d" % i] = leaves_w[i].grad.cpu()
        return {k: (v.detach().cpu() if torch.is_tensor(v) else {kk: vv.cpu() for kk, vv in v.items()}) for k, v in res.items()}, grads

    cfg = ref_arm.reference_config(const_iter=const_iter)

    def make_ref():
        return r.ICP(config_path=cfg, icp_type=icp_type, max_iterations=K, tolerance=1e-9, differentiable=diff)

    def make_ours():
        icp = ICP(icp_type=icp_type, max_iterations=K, tolerance=1e-9, differentiable=diff)
        icp.const_iter = const_iter
        return icp

    want, gw = run(make_ref, "cpu")
    got, gg = run(make_ours, "cuda")
    assert got["deltas"].shape == want["deltas"].shape, "iteration count differs"
    for key in ("T", "pc", "weights", "costs"):
        assert rel_err(got[key], want[key]) < tol * (10 if key == "costs" else 1), key
    assert float((got["deltas"] - want["deltas"]).abs().max()) < tol * max(1.0, float(want["deltas"].abs().max()))
    assert torch.equal(got["stats"]["converged"], want["stats"]["converged"])
    assert torch.equal(got["stats"]["iterations"], want["stats"]["iterations"])
    assert torch.allclose(got["stats"]["matched_ratio"], want["stats"]["matched_ratio"], rtol=0, atol=1e-12)
    for key in gw:
        g_ref = torch.nan_to_num(gw[key], nan=0.0)         # hard Huber at an exactly-zero residual is NaN upstream (c / 0 under where)
        assert rel_err(gg[key], g_ref) < 1e-6 or float((gg[key] - g_ref).abs().max()) < 1e-9, key
