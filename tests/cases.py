"""Golden-case tables shared by the CPU (oracle, host emulation) and GPU parity tests."""
import numpy as np
import torch

from helpers import load_golden, tt

# name, icp_type, diff, loss, metric, trim, max_it, tol, source_zeroes_are_pad
FIXTURE_CASES = [
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

# name, icp_type, diff, loss, metric, trim, K, const_iter, tol, out_tol, grad_tol
DENSE_CASES = [
    ("dense_pt2pl_cauchy_trim_f64", "pt2pl", True, "cauchy", 0.5, 2.0, 6, True, 1e-12, 1e-10, 1e-8),
    ("dense_pt2pl_cauchy_trim_f32", "pt2pl", True, "cauchy", 0.5, 2.0, 6, True, 1e-12, 2e-5, 2e-3),
    ("dense_pt2pt_huber_trim_f64", "pt2pt", True, "huber", 0.3, 2.0, 5, True, 1e-12, 1e-10, 1e-8),
    ("dense_pt2pt_huber_trim_f32", "pt2pt", True, "huber", 0.3, 2.0, 5, True, 1e-12, 2e-5, 2e-3),
    ("dense_pt2pl_nondiff_huber_f64", "pt2pl", False, "huber", 0.3, 1.0, 30, False, 1e-9, 1e-9, 1e-7),
    ("dense_pt2pt_plain_earlystop_f64", "pt2pt", True, None, 1.0, None, 40, False, 1e-9, 1e-9, 1e-7),
    ("dense_pt2pl_plain_f64", "pt2pl", True, None, 1.0, None, 4, True, 1e-12, 1e-10, 1e-8),
]


def golden_loss_grads(T, pc):
    """d/dT and d/dpc of the loss make_golden.py back-propagated (loss_of)."""
    gT = torch.zeros_like(T)
    gT[:, :3, :] = torch.arange(1, 13, dtype=T.dtype, device=T.device).reshape(3, 4)
    wp = torch.linspace(-1.0, 1.0, pc.shape[1] * 3, dtype=pc.dtype).reshape(pc.shape[1], 3).to(pc.device)
    gpc = (0.01 * wp).expand_as(pc).contiguous()
    return gT, gpc


def golden_loss(res):
    gT, gpc = golden_loss_grads(res["T"].detach(), res["pc"].detach())
    return (res["T"] * gT).sum() + (res["pc"] * gpc).sum()


def planar(source, target, pt2pl):
    """ICP.py:107-116 on detached tensors."""
    s, t = source.clone(), target.clone()
    s[..., 2] = 0
    t[..., 2] = 0
    if pt2pl:
        t[..., 5] = 0
    return s, t
