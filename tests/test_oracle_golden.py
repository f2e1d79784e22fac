"""CPU: the oracle restatement vs golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle (prompt ③)."""
import numpy as np
import pytest
import torch

from oracle.dicp_oracle import IcpSettings, icp_forward, find_nn, robust_weight, trim_weight
from helpers import load_golden, rel_err, se3_log, T_ts_true, tt


def golden_loss(res):
    T, pc = res["T"], res["pc"]
    wT = torch.arange(1, 13, dtype=T.dtype).reshape(3, 4)
    wp = torch.linspace(-1.0, 1.0, pc.shape[1] * 3, dtype=pc.dtype).reshape(pc.shape[1], 3)
    return (T[:, :3, :] * wT).sum() + 0.01 * (pc * wp).sum()


def grad_close(mine, ref, tol):
    """Gradients agree; where the reference's autograd yields NaN (non-differentiable
    weight branches with an exactly-zero residual) the oracle must yield NaN too."""
    mine = np.asarray(mine.detach() if hasattr(mine, "detach") else mine, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    nan = np.isnan(ref)
    if not np.array_equal(np.isnan(mine), nan):
        return False
    if nan.all():
        return True
    return rel_err(mine[~nan], ref[~nan]) < tol


def check_outputs(res, g, tol):
    assert res["T"].shape == g["T"].shape
    assert rel_err(res["T"].detach(), g["T"]) < tol
    assert rel_err(res["pc"].detach(), g["pc"]) < tol
    assert res["deltas"].shape == g["deltas"].shape, "iteration count differs"
    assert res["weights"].shape == g["weights"].shape
    assert res["costs"].shape == g["costs"].shape
    assert np.abs(res["deltas"].numpy() - g["deltas"]).max() < tol * max(1.0, np.abs(g["deltas"]).max())
    assert rel_err(res["weights"], g["weights"]) < tol
    assert rel_err(res["costs"], g["costs"]) < 10 * tol
    assert np.array_equal(res["stats"]["converged"].numpy(), g["converged"])
    assert np.array_equal(res["stats"]["iterations"].numpy(), g["iterations"])
    assert np.allclose(res["stats"]["matched_ratio"].numpy(), g["matched_ratio"], rtol=0, atol=1e-7)


FIXTURE_CASES = [
    # name, icp_type, diff, loss, metric, trim, max_it, tol, szap
    ("fix_pt2pt_diff_huber", "pt2pt", True, "huber", 1.0, 5.0, 100, 1e-10, False),
    ("fix_pt2pl_diff_huber", "pt2pl", True, "huber", 10.0, 5.0, 100, 1e-10, False),
    ("fix_pt2pt_nondiff_huber", "pt2pt", False, "huber", 10.0, 5.0, 100, 1e-10, False),
    ("fix_pt2pl_diff_cauchy", "pt2pl", True, "cauchy", 0.5, 5.0, 100, 1e-10, False),
    ("fix_pt2pl_nondiff_huber_sub", "pt2pl", False, "huber", 1.0, 5.0, 25, 1e-8, False),
    ("fix_pt2pl_diff_huber_sub", "pt2pl", True, "huber", 1.0, 5.0, 25, 1e-8, False),
    ("fix_pt2pl_nondiff_cauchy_sub", "pt2pl", False, "cauchy", 0.5, 5.0, 25, 1e-8, False),
    ("fix_pt2pt_nondiff_notrim_sub", "pt2pt", False, None, 1.0, None, 25, 1e-8, True),
    ("fix_pt2pt_nondiff_notrim_pad", "pt2pt", False, None, 1.0, None, 25, 1e-8, True),
]


@pytest.mark.parametrize("case", FIXTURE_CASES, ids=[c[0] for c in FIXTURE_CASES])
def test_fixture_cases(case):
    name, icp_type, diff, lname, metric, trim, max_it, tol, szap = case
    g = load_golden(name)
    src = tt(g["source"], grad=True)
    tgt = tt(g["target"], grad=True)
    T0 = tt(g["T_init"], grad=True)
    w = torch.ones((1, src.shape[0]), dtype=src.dtype)
    if szap:
        w = w * (torch.linalg.norm(src.detach(), dim=1) != 0).to(src.dtype)
    if icp_type == "pt2pt":
        w = w.repeat_interleave(3, dim=1)
    cfg = IcpSettings(icp_type=icp_type, max_iterations=max_it, tolerance=tol, differentiable=diff,
                      trim_dist=trim, loss_name=lname, loss_metric=metric, dim=2)
    idx_log = []
    res = icp_forward(src.unsqueeze(0), tgt.unsqueeze(0), T0.unsqueeze(0), w, cfg, record_idx=idx_log)
    check_outputs(res, g, 1e-9)
    assert np.array_equal(torch.stack(idx_log, 0).numpy(), g["nn_idx"])
    golden_loss(res).backward()
    ns = g["g_source"].shape[0]
    assert grad_close(src.grad[:ns], g["g_source"], 1e-7)
    assert grad_close(tgt.grad, g["g_target"], 1e-7)
    assert grad_close(T0.grad, g["g_T_init"], 1e-7)


def test_fixture_known_answer():
    """Reference tests/test_ICP.py:65-69 on the oracle: ||log(T_true T_pred^-1)|| < 1e-10."""
    for name, icp_type, lname, metric in (("fix_pt2pt_diff_huber", "pt2pt", "huber", 1.0),
                                          ("fix_pt2pl_diff_huber", "pt2pl", "huber", 10.0)):
        g = load_golden(name)
        src, tgt = tt(g["source"]), tt(g["target"])
        w = torch.ones((1, src.shape[0] * (3 if icp_type == "pt2pt" else 1)), dtype=src.dtype)
        cfg = IcpSettings(icp_type=icp_type, max_iterations=100, tolerance=1e-10, trim_dist=5.0,
                          loss_name=lname, loss_metric=metric, dim=2)
        res = icp_forward(src[None], tgt[None], torch.eye(4, dtype=src.dtype)[None], w, cfg)
        xi = se3_log(T_ts_true() @ np.linalg.inv(res["T"][0].numpy()))
        assert np.linalg.norm(xi) < 1e-10
        assert np.allclose(res["pc"][0].numpy(), g["target"][:, :3], atol=1e-5)


DENSE_CASES = [
    # name, icp_type, diff, loss, metric, trim, K, const_iter, tol, out_tol, grad_tol
    ("dense_pt2pl_cauchy_trim_f64", "pt2pl", True, "cauchy", 0.5, 2.0, 6, True, 1e-12, 1e-10, 1e-8),
    ("dense_pt2pl_cauchy_trim_f32", "pt2pl", True, "cauchy", 0.5, 2.0, 6, True, 1e-12, 2e-5, 2e-3),
    ("dense_pt2pt_huber_trim_f64", "pt2pt", True, "huber", 0.3, 2.0, 5, True, 1e-12, 1e-10, 1e-8),
    ("dense_pt2pt_huber_trim_f32", "pt2pt", True, "huber", 0.3, 2.0, 5, True, 1e-12, 2e-5, 2e-3),
    ("dense_pt2pl_nondiff_huber_f64", "pt2pl", False, "huber", 0.3, 1.0, 30, False, 1e-9, 1e-9, 1e-7),
    ("dense_pt2pt_plain_earlystop_f64", "pt2pt", True, None, 1.0, None, 40, False, 1e-9, 1e-9, 1e-7),
    ("dense_pt2pl_plain_f64", "pt2pl", True, None, 1.0, None, 4, True, 1e-12, 1e-10, 1e-8),
]


@pytest.mark.parametrize("case", DENSE_CASES, ids=[c[0] for c in DENSE_CASES])
def test_dense_cases(case):
    name, icp_type, diff, lname, metric, trim, K, const_iter, tol, out_tol, grad_tol = case
    g = load_golden(name)
    src = tt(g["source"], grad=True)
    tgt = tt(g["target"], grad=True)
    T0 = tt(g["T_init"], grad=True)
    N, n = src.shape[0], src.shape[1]
    if "weight" in g:
        wu = tt(g["weight"], grad=True)
    else:
        wu = None
    w = wu if wu is not None else torch.ones((N, n), dtype=src.dtype)
    if icp_type == "pt2pt":
        w = w.repeat_interleave(3, dim=1)
    cfg = IcpSettings(icp_type=icp_type, max_iterations=K, tolerance=tol, differentiable=diff,
                      const_iter=const_iter, trim_dist=trim, loss_name=lname, loss_metric=metric, dim=3)
    idx_log = []
    res = icp_forward(src, tgt, T0, w, cfg, record_idx=idx_log)
    assert np.array_equal(torch.stack(idx_log, 0).numpy(), g["nn_idx"])
    check_outputs(res, g, out_tol)
    golden_loss(res).backward()
    assert grad_close(src.grad, g["g_source"], grad_tol)
    assert grad_close(tgt.grad, g["g_target"], grad_tol)
    assert grad_close(T0.grad, g["g_T_init"], grad_tol)
    if wu is not None:
        assert grad_close(wu.grad, g["g_weight"], grad_tol)


def test_ragged_padded_dense():
    """The reference's padded tensors for the ragged list batch (ICP.py:361-482) fed to the
    oracle loop reproduce the reference's batched result (tests/test_ICP_inputs.py:36-110)."""
    g = load_golden("ragged_pt2pl")
    src, tgt, w = tt(g["pad_source"]), tt(g["pad_target"]), tt(g["pad_w"])
    T0 = torch.eye(4, dtype=src.dtype).repeat(3, 1, 1)
    cfg = IcpSettings(icp_type="pt2pl", max_iterations=25, tolerance=1e-8, trim_dist=5.0,
                      loss_name="huber", loss_metric=1.0, dim=2)
    res = icp_forward(src, tgt, T0, w, cfg)
    check_outputs(res, g, 1e-9)


def test_nn_known_answers():
    g = load_golden("nn_known_answers")
    out, _ = find_nn(tt(g["q"])[None], tt(g["pts"])[None])
    assert np.array_equal(out[0].numpy(), g["a1"][0]) and np.array_equal(g["a1"][0, 0], [8.0, 7.0, 0.0])
    out, _ = find_nn(tt(g["q"])[None], tt(g["pts2"])[None])
    assert np.array_equal(out[0].numpy(), g["a2"][0]) and np.array_equal(g["a2"][0, 0], [10.0, 2.0, 0.0])
    y = tt(g["y"], grad=True)
    out, _ = find_nn(tt(g["x"]), y)
    assert np.array_equal(out.detach().numpy(), g["out"])
    (out * tt(g["gw"])).sum().backward()
    assert np.array_equal(y.grad.numpy(), g["g_y"])
    assert np.array_equal(g["out_cf"], g["out"])


def test_loss_known_answers():
    g = load_golden("loss_known_answers")
    for key in ("e3", "e1"):
        e = tt(g[key])
        for diff in (True, False):
            assert np.allclose(robust_weight("huber", e, 1.3, diff).numpy(), g["huber_%d_%s" % (diff, key)], rtol=1e-14)
            assert np.allclose(robust_weight("cauchy", e, 0.7, diff).numpy(), g["cauchy_%d_%s" % (diff, key)], rtol=1e-14)
            assert np.allclose(trim_weight(e, 1.9, diff, 5.0).numpy(), g["trim_%d_%s" % (diff, key)], rtol=1e-14)


def test_kabsch_fixture_golden_and_known_answer():
    """ICP.py:533-591: the literal restatement reproduces the reference on its planar fixture, and
    the Kabsch-correct variant reaches the same known answer (tests/test_ICP.py:45-47)."""
    from oracle.dicp_oracle import kabsch_icp
    g = load_golden("svd_fixture")
    src, tgt = tt(g["source"]), tt(g["target"])
    ps, T, _ = kabsch_icp(src, tgt, torch.eye(4, dtype=src.dtype), 50, 1e-20, literal=True)
    assert rel_err(T, g["T"]) < 1e-12 and rel_err(ps, g["pc"]) < 1e-12
    ps2, T2, _ = kabsch_icp(src, tgt, torch.eye(4, dtype=src.dtype), 50, 1e-20)
    assert np.linalg.norm(se3_log(T_ts_true() @ np.linalg.inv(T2.numpy()))) < 1e-10
    assert rel_err(T2, g["T"]) < 1e-10


def test_kabsch_oracle_recovers_3d_motion():
    """On a general 3-D cloud only the Kabsch rotation U diag V^T recovers the motion."""
    from oracle.dicp_oracle import kabsch_icp
    gen = torch.Generator().manual_seed(0)
    P = torch.randn((300, 3), generator=gen, dtype=torch.float64) * 3
    a, b = 0.03, 0.02
    Rz = torch.tensor([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], dtype=torch.float64)
    Rx = torch.tensor([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]], dtype=torch.float64)
    R, t = Rz @ Rx, torch.tensor([0.05, -0.02, 0.03], dtype=torch.float64)
    Q = P @ R.T + t
    ps, T, it = kabsch_icp(P, Q, torch.eye(4, dtype=torch.float64), 30, 1e-24)
    assert (T[:3, :3] - R).abs().max() < 1e-10 and (T[:3, 3] - t).abs().max() < 1e-10
    _, T_lit, _ = kabsch_icp(P, Q, torch.eye(4, dtype=torch.float64), 30, 1e-24, literal=True)
    assert (T_lit[:3, :3] - R).abs().max() > 1e-3


def test_product_side_input_generator_equals_the_oracle_side_one():
    """bench.py and the tools draw their synthetic clouds from dicp_b200.synthetic (the product path must not
    import oracle/); the parity tests draw theirs from the oracle module.  Same numbers, bit for bit."""
    from dicp_b200.synthetic import surface_pairs
    from oracle.dicp_oracle import make_wavy_pair
    for n_pairs, n, m, seed, dtype, normals in ((3, 100, 80, 5, torch.float32, True), (2, 50, 70, 9, torch.float64, False)):
        a = surface_pairs(n_pairs, n, m, seed, dtype=dtype, with_normals=normals, extent=33.0)
        b = make_wavy_pair(n_pairs, n, m, seed, dtype=dtype, with_normals=normals, extent=33.0)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
