"""CPU: the Kabsch / SVD step of the CUDA path (dicp_math.cuh: svd3, kabsch_step, kabsch_step_backward, compiled for the
host by tests/host_emul) against torch.svd and torch autograd (ICP.py:557-573)."""
import ctypes

import numpy as np
import pytest
import torch

import emul


def _step_torch(sums, n, literal):
    xbar, qbar = sums[0:3] / n, sums[3:6] / n
    W = sums[6:15].reshape(3, 3) / n - torch.outer(qbar, xbar)            # ICP.py:562-563
    U, S, V = torch.svd(W)                                                 # ICP.py:566
    d = (torch.det(U) * torch.det(V)).detach()
    C = U @ torch.diag(torch.stack((torch.ones_like(d), torch.ones_like(d), d))) @ (V if literal else V.T)   # ICP.py:570
    r = qbar - C @ xbar                                                    # ICP.py:573
    return C, r, S


def _sums(seed, n=200, planar=False, reflect=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((n, 3), generator=g, dtype=torch.float64) * torch.tensor([3.0, 2.0, 0.5], dtype=torch.float64)
    if planar:
        x[:, 2] = 0.0
    ang = 0.3
    R = torch.tensor([[np.cos(ang), -np.sin(ang), 0.0], [np.sin(ang), np.cos(ang), 0.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
    q = x @ R.T + 0.05 * torch.randn((n, 3), generator=g, dtype=torch.float64) + torch.tensor([0.3, -0.2, 0.1], dtype=torch.float64)
    if planar:
        q[:, 2] = 0.0
    if reflect:
        q[:, 2] = -q[:, 2] * 3.0
    return torch.cat((x.sum(0), q.sum(0), (q.T @ x).reshape(9))), float(n)


def _run(sums, n, literal, Cbar, rbar):
    lib = emul.load()
    lib.emul_kabsch_step.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_int, ctypes.c_void_p]
    lib.emul_kabsch_step_backward.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    rec = np.zeros(64)
    s_np = np.ascontiguousarray(sums.detach().numpy())
    lib.emul_kabsch_step(s_np.ctypes.data, n, int(literal), rec.ctypes.data)
    g = np.zeros(15)
    cb, rb = np.ascontiguousarray(Cbar.numpy()), np.ascontiguousarray(rbar.numpy())
    lib.emul_kabsch_step_backward(rec.ctypes.data, n, int(literal), cb.ctypes.data, rb.ctypes.data, g.ctypes.data)
    return rec, g


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("kind", ["general", "planar", "reflect"])
def test_kabsch_step_forward_and_adjoint_match_torch(seed, kind):
    sums, n = _sums(seed, planar=kind == "planar", reflect=kind == "reflect")
    sums = sums.clone().requires_grad_(True)
    C, r, S = _step_torch(sums, n, False)
    g = torch.Generator().manual_seed(100 + seed)
    Cbar, rbar = torch.randn((3, 3), generator=g, dtype=torch.float64), torch.randn(3, generator=g, dtype=torch.float64)
    rec, grad = _run(sums, n, False, Cbar, rbar)
    assert np.allclose(rec[21:24], S.detach().numpy(), rtol=1e-12, atol=1e-14)           # singular values, descending
    assert np.allclose(rec[40:49].reshape(3, 3), C.detach().numpy(), atol=1e-12)
    assert np.allclose(rec[49:52], r.detach().numpy(), atol=1e-12)
    if kind == "reflect":
        assert abs(rec[33] + 1.0) < 1e-12
    if kind != "planar":            # torch's own SVD gradient is singular for a rank-deficient W; ours is defined (checked below)
        ((C * Cbar).sum() + (r * rbar).sum()).backward()
        assert np.allclose(grad, sums.grad.numpy(), rtol=1e-8, atol=1e-10)


def test_kabsch_step_adjoint_matches_finite_differences_on_planar_clouds():
    sums, n = _sums(7, planar=True)
    g = torch.Generator().manual_seed(5)
    Cbar, rbar = torch.randn((3, 3), generator=g, dtype=torch.float64), torch.randn(3, generator=g, dtype=torch.float64)
    Cbar[2, :] = 0.0
    Cbar[:, 2] = 0.0               # in-plane loss: the out-of-plane part of C is not determined by a planar cloud
    rbar[2] = 0.0
    rec, grad = _run(sums, n, False, Cbar, rbar)

    def loss(s):
        rc, _ = _run(s, n, False, Cbar, rbar)
        return float((rc[40:49].reshape(3, 3) * Cbar.numpy()).sum() + (rc[49:52] * rbar.numpy()).sum())
    for k in (0, 1, 3, 4, 6, 7, 9, 10):          # in-plane sums only
        e = torch.zeros(15, dtype=torch.float64)
        e[k] = 1e-5
        fd = (loss(sums + e) - loss(sums - e)) / 2e-5
        assert abs(fd - grad[k]) < 1e-6 * max(1.0, abs(fd)), (k, fd, grad[k])


def test_kabsch_step_literal_product_and_adjoint():
    """C = U D V (what ICP.py:570 computes, torch.svd returning V): same numbers as torch when the signs of the singular
    vectors agree, which the test arranges by comparing sign-invariant quantities: C_literal V^T ... = U D V V^T."""
    sums, n = _sums(11)
    rec, _ = _run(sums, n, True, torch.zeros((3, 3), dtype=torch.float64), torch.zeros(3, dtype=torch.float64))
    U, V, D = rec[12:21].reshape(3, 3), rec[24:33].reshape(3, 3), np.diag([1.0, 1.0, rec[33]])
    assert np.allclose(rec[40:49].reshape(3, 3), U @ D @ V, atol=1e-13)
    # adjoint of the literal product vs finite differences
    g = torch.Generator().manual_seed(9)
    Cbar, rbar = torch.randn((3, 3), generator=g, dtype=torch.float64), torch.randn(3, generator=g, dtype=torch.float64)
    _, grad = _run(sums, n, True, Cbar, rbar)

    def loss(s):
        rc, _ = _run(s, n, True, Cbar, rbar)
        return float((rc[40:49].reshape(3, 3) * Cbar.numpy()).sum() + (rc[49:52] * rbar.numpy()).sum())
    for k in range(15):
        e = torch.zeros(15, dtype=torch.float64)
        e[k] = 1e-4
        fd = (loss(sums + e) - loss(sums - e)) / 2e-4
        assert abs(fd - grad[k]) < 1e-6 * max(1.0, abs(fd)), (k, fd, grad[k])
