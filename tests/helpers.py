"""Shared test helpers: golden loading, SE(3) log/exp restatements of the two pylgmath
calls the reference tests use (`Transformation(xi_ab=xi).matrix()` = expm(hat(xi)) with
xi = [rho; phi], and `se3op.tran2vec` = vee(logm(T)); reference tests/test_ICP.py:45-47,65)."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def hat3(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)


def se3_exp(xi):
    """xi = [rho(3); phi(3)] -> 4x4, closed form."""
    xi = np.asarray(xi, dtype=np.float64).reshape(6)
    rho, phi = xi[:3], xi[3:]
    th = np.linalg.norm(phi)
    K = hat3(phi)
    if th < 1e-12:
        R = np.eye(3) + K
        V = np.eye(3) + 0.5 * K
    else:
        R = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * K + (th - np.sin(th)) / th ** 3 * K @ K
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = V @ rho
    return T


def se3_log(T):
    """4x4 (or batch) -> xi = [rho; phi] (…,6,1) like pylgmath.se3op.tran2vec."""
    T = np.asarray(T, dtype=np.float64)
    if T.ndim == 3:
        return np.stack([se3_log(t) for t in T], 0)
    R, t = T[:3, :3], T[:3, 3]
    c = np.clip((np.trace(R) - 1) / 2, -1, 1)
    th = np.arccos(c)
    if th < 1e-10:
        phi = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
        Vinv = np.eye(3) - 0.5 * hat3(phi)
    else:
        phi = th / (2 * np.sin(th)) * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
        K = hat3(phi)
        Vinv = np.eye(3) - 0.5 * K + (1 / th ** 2) * (1 - th * np.sin(th) / (2 * (1 - np.cos(th)))) * K @ K
    return np.concatenate([Vinv @ t, phi]).reshape(6, 1)


def T_ts_true():
    """Known answer of the reference fixture (tests/test_ICP.py:45-47)."""
    return np.linalg.inv(se3_exp([1.0, 1.0, 0.0, 0.0, 0.0, 0.1]))


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def tt(x, dtype=None, grad=False):
    t = torch.tensor(np.asarray(x))
    if dtype is not None:
        t = t.to(dtype)
    if grad:
        t.requires_grad_(True)
    return t
