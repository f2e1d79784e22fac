"""CPU: the kernels' arithmetic (dicp_math.cuh compiled for the host, tests/host_emul) against
the golden vectors of the unmodified reference.  Covers eval_point / accum_point /
pair_update / final_stats / pair_backward / point_backward — everything but the CUDA launch
geometry and reductions, which the -m gpu tests cover."""
import numpy as np
import pytest
import torch

import emul
from cases import DENSE_CASES, FIXTURE_CASES, golden_loss_grads, planar
from dicp_b200._functions import make_params
from helpers import load_golden, rel_err, tt


def compare_outputs(out, g, tol):
    assert out["deltas"].shape[1] == g["deltas"].shape[1], "iteration count differs"
    assert rel_err(out["T"], g["T"]) < tol
    assert rel_err(out["pc"], g["pc"]) < tol
    assert np.abs(out["deltas"].numpy() - g["deltas"][..., 0]).max() < tol * max(1.0, np.abs(g["deltas"]).max())
    assert rel_err(out["weights"], g["weights"][..., 0]) < tol
    assert rel_err(out["costs"], g["costs"][..., 0]) < 10 * tol
    assert np.array_equal(out["converged"].numpy(), g["converged"])
    assert np.array_equal(out["iterations"].numpy(), g["iterations"])
    assert np.allclose(out["matched_ratio"].numpy(), g["matched_ratio"], rtol=0, atol=1e-7)
    assert np.array_equal(out["saved_idx"][: out["k_run"]].numpy(), g["nn_idx"])


@pytest.mark.parametrize("case", FIXTURE_CASES, ids=[c[0] for c in FIXTURE_CASES])
def test_fixture_cases(case):
    name, icp_type, diff, lname, metric, trim, max_it, tol, szap = case
    g = load_golden(name)
    pl = icp_type == "pt2pl"
    src, tgt = planar(tt(g["source"])[None], tt(g["target"])[None], pl)
    T0 = tt(g["T_init"])[None]
    w = torch.ones((1, src.shape[1]), dtype=src.dtype)
    if szap:
        w = w * (torch.linalg.norm(tt(g["source"])[None], dim=2) != 0).to(src.dtype)
    p = make_params(icp_type, 2, diff, lname, metric, trim, 5.0, tol, max_it, False, 0.0)
    out = emul.forward(src, tgt, T0, w, p)
    compare_outputs(out, g, 1e-9)
    if np.isnan(g["g_source"]).any():
        return          # reference autograd is NaN for hard Huber at zero residual; ours is finite by design
    gT, gpc = golden_loss_grads(out["T"], out["pc"])
    gr = emul.backward(src, tgt, w, p, out, gT, gpc)
    ns = g["g_source"].shape[0]
    gs, gt = gr["source"][0, :ns].clone(), gr["target"][0].clone()
    gs[:, 2] = 0
    gt[:, 2] = 0
    if pl:
        gt[:, 5] = 0
    assert rel_err(gs, g["g_source"]) < 1e-7
    assert rel_err(gt, g["g_target"]) < 1e-7
    assert rel_err(gr["T_init"][0], g["g_T_init"]) < 1e-7


@pytest.mark.parametrize("case", DENSE_CASES, ids=[c[0] for c in DENSE_CASES])
def test_dense_cases(case):
    name, icp_type, diff, lname, metric, trim, K, const_iter, tol, out_tol, grad_tol = case
    g = load_golden(name)
    src, tgt, T0 = tt(g["source"]), tt(g["target"]), tt(g["T_init"])
    w = tt(g["weight"]) if "weight" in g else None
    p = make_params(icp_type, 3, diff, lname, metric, trim, 5.0, tol, K, const_iter, 0.0)
    out = emul.forward(src, tgt, T0, w, p)
    compare_outputs(out, g, out_tol)
    gT, gpc = golden_loss_grads(out["T"], out["pc"])
    gr = emul.backward(src, tgt, w, p, out, gT, gpc)
    assert rel_err(gr["source"], g["g_source"]) < grad_tol
    assert rel_err(gr["target"], g["g_target"]) < grad_tol
    assert rel_err(gr["T_init"], g["g_T_init"]) < grad_tol
    if w is not None:
        assert rel_err(gr["w"], g["g_weight"]) < grad_tol


def test_ragged_padded():
    g = load_golden("ragged_pt2pl")
    src, tgt = planar(tt(g["pad_source"]), tt(g["pad_target"]), True)
    T0 = torch.eye(4, dtype=src.dtype).repeat(3, 1, 1)
    p = make_params("pt2pl", 2, True, "huber", 1.0, 5.0, 5.0, 1e-8, 25, False, 0.0)
    out = emul.forward(src, tgt, T0, tt(g["pad_w"]), p)
    assert out["k_run"] == g["deltas"].shape[1]
    assert rel_err(out["T"], g["T"]) < 1e-9
    assert rel_err(out["weights"], g["weights"][..., 0]) < 1e-9      # incl. carry-forward rows of frozen pairs
    assert rel_err(out["costs"], g["costs"][..., 0]) < 1e-8
    assert np.array_equal(out["iterations"].numpy(), g["iterations"])
    assert np.allclose(out["matched_ratio"].numpy(), g["matched_ratio"], rtol=0, atol=1e-7)
    # gradient through frozen iterations: T_init.grad of the reference is not stored for this case,
    # but source/target grads are
    gT = torch.zeros((3, 4, 4), dtype=src.dtype)
    gT[:, :3, :] = torch.arange(1, 13, dtype=src.dtype).reshape(3, 4)
    gr = emul.backward(src, tgt, tt(g["pad_w"]), p, out, gT, None)
    for b, key, rows in ((1, "g_s2", 65), (2, "g_s3", 55)):
        gs = gr["source"][b, :rows].clone()
        gs[:, 2] = 0
        assert rel_err(gs, g[key]) < 1e-7
    for b, key, rows in ((0, "g_t1", 55), (1, "g_t2", 65), (2, "g_t3", 60)):
        gt = gr["target"][b, :rows].clone()
        gt[:, 2] = 0
        gt[:, 5] = 0
        assert rel_err(gt, g[key]) < 1e-7
