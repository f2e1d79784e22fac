"""CPU oracle for the dICP differentiable-ICP inner loop.  TEST INFRASTRUCTURE ONLY.

This file is a torch-CPU restatement of the reference's algorithm for the hot path
(`dICP/ICP.py:132-303`, `dICP/nn.py:23-40`, `dICP/loss.py:11-58` of utiasASRL/dICP).
It is the checker for the CUDA path: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  Nothing under
`dicp_b200/` imports it, and it is never the thing measured as the product or shipped.

Parity status: PINNED.  `tests/golden/make_golden.py` imports the unmodified
reference from `/root/reference` in the build container, runs it on the reference's
own fixture (`tests/data/points_{scan,map}.npy`) and on seeded synthetic clouds, and
commits inputs + outputs (T, pc, costs, deltas, weights, stats, autograd gradients,
per-iteration NN indices) under `tests/golden/`.  `tests/test_oracle_golden.py` checks
this restatement against every one of those vectors and against the reference tests'
known answer `T_ts_true` (`tests/test_ICP.py:45-47`).

All arithmetic is float torch on CPU in the dtype of the inputs; gradients of the
oracle come from torch autograd through this unrolled loop, which is exactly how the
reference obtains them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

__all__ = [
    "IcpSettings",
    "nn_index_reference",
    "nn_index_exact",
    "gather_rows",
    "find_nn",
    "trim_weight",
    "robust_weight",
    "icp_forward",
    "kabsch_icp",
    "make_wavy_pair",
]


# --------------------------------------------------------------------------------------
# settings
# --------------------------------------------------------------------------------------
@dataclass
class IcpSettings:
    """Everything `ICP.__init__` + `ICP.dICP`'s keyword arguments pin down
    (`ICP.py:15-44`, `ICP.py:49`, `config/dICP_config.yaml:1-20`)."""

    icp_type: str = "pt2pl"          # 'pt2pl' | 'pt2pt'           ICP.py:30
    max_iterations: int = 100        #                              ICP.py:31
    tolerance: float = 1e-12         #                              ICP.py:32
    differentiable: bool = True      #                              ICP.py:38
    const_iter: bool = False         # yaml parameters.const_iter   ICP.py:33
    tanh_steepness: float = 5.0      # yaml parameters.tanh_steepness ICP.py:119
    match_ratio_thresh: float = 0.0  # yaml logging.matched_ratio_thresh ICP.py:37
    trim_dist: Optional[float] = None  # None or <0 => off          ICP.py:153
    loss_name: Optional[str] = None  # None | 'huber' | 'cauchy'    ICP.py:158
    loss_metric: float = 1.0
    dim: int = 3                     # 2 | 3                        ICP.py:79


# --------------------------------------------------------------------------------------
# nearest neighbour  (nn.py:23-40 == nn.py:72-92)
# --------------------------------------------------------------------------------------
def nn_index_reference(x: torch.Tensor, y_xyz: torch.Tensor) -> torch.Tensor:
    """Index of the nearest target row for every query, exactly as the reference
    computes it: dense Euclidean `cdist` then first-minimum `argmin` (nn.py:32-35).
    x (N,n,3), y_xyz (N,m,3) -> (N,n) int64.  O(N*n*m) memory."""
    return torch.argmin(torch.cdist(x, y_xyz, p=2), dim=2)


def nn_index_exact(x: torch.Tensor, y_xyz: torch.Tensor, chunk: int = 2048):
    """fp64 direct-difference NN: the parity target for the CUDA search.

    Returns (idx (N,n) int64, d2_best (N,n) fp64, d2_second (N,n) fp64).  `d2_second`
    is the smallest squared distance among the *other* targets, so a test can decide
    whether a query is tie-free (relative gap rule).  First-minimum
    tie-break like `argmin` (nn.py:35).  Row-chunked so n=m=130k fits in host RAM."""
    X = x.detach().to(torch.float64)
    Y = y_xyz.detach().to(torch.float64)
    N, n, _ = X.shape
    m = Y.shape[1]
    idx = torch.empty((N, n), dtype=torch.int64)
    best = torch.empty((N, n), dtype=torch.float64)
    second = torch.empty((N, n), dtype=torch.float64)
    for b in range(N):
        for s in range(0, n, chunk):
            q = X[b, s:s + chunk]                                   # (c,3)
            d2 = ((q[:, None, :] - Y[b][None, :, :]) ** 2).sum(-1)   # (c,m)
            j = torch.argmin(d2, dim=1)
            idx[b, s:s + chunk] = j
            best[b, s:s + chunk] = d2.gather(1, j[:, None]).squeeze(1)
            if m > 1:
                d2.scatter_(1, j[:, None], float("inf"))
                second[b, s:s + chunk] = d2.min(dim=1).values
            else:
                second[b, s:s + chunk] = float("inf")
    return idx, best, second


def gather_rows(y: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """Rows of y (N,m,S) picked by idx (N,n) -> (N,n,S); differentiable w.r.t. y
    (backward is a scatter-add, the only gradient path into the target; nn.py:37-38)."""
    return torch.gather(y, 1, idx.unsqueeze(-1).expand(-1, -1, y.shape[-1]))


def find_nn(x: torch.Tensor, y: torch.Tensor, idx: Optional[torch.Tensor] = None):
    """Hard nearest-neighbour lookup `nn.find_nn` with `use_gumbel=False`
    (nn.py:11-40).  x (N,n,3), y (N,m,3|6) -> neighbours (N,n,3|6), idx (N,n)."""
    if idx is None:
        idx = nn_index_reference(x, y[:, :, :3])
    return gather_rows(y, idx), idx


# --------------------------------------------------------------------------------------
# weights  (loss.py)
# --------------------------------------------------------------------------------------
def trim_weight(err3: torch.Tensor, tau: float, differentiable: bool, steep: float):
    """loss.py:43-58 on the 3-vector residual (N,n,3) -> (N,n)."""
    d = torch.linalg.norm(err3, dim=-1)
    if differentiable:
        return 0.5 * torch.tanh(steep * (tau - d) - 3.0) + 0.5
    return torch.where(d < tau, torch.ones_like(d), torch.zeros_like(d))


def robust_weight(name: str, err: torch.Tensor, c: float, differentiable: bool):
    """loss.py:21-41 on err (N,n,k) -> (N,n); k=1 for pt2pl (|rho|), 3 for pt2pt."""
    d = torch.linalg.norm(err, dim=-1)
    if name == "cauchy":
        return 1.0 / (1.0 + (d / c) ** 2)
    if name == "huber":
        if differentiable:
            return c ** 2 / (c ** 2 + d ** 2)
        return torch.where(d > c, c / d, torch.ones_like(d))
    raise ValueError("Invalid loss name: {}".format(name))      # loss.py:19


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def _hat(v: torch.Tensor) -> torch.Tensor:
    """(..., 3) -> (..., 3, 3) cross-product matrix (ICP.py:513-531)."""
    z = torch.zeros_like(v[..., 0])
    rows = [z, -v[..., 2], v[..., 1],
            v[..., 2], z, -v[..., 0],
            -v[..., 1], v[..., 0], z]
    return torch.stack(rows, dim=-1).reshape(*v.shape[:-1], 3, 3)


def _count_gt(w: torch.Tensor, thresh: float) -> torch.Tensor:
    return torch.sum(w > thresh, dim=1)


# --------------------------------------------------------------------------------------
# the loop  (ICP.py:86-303)
# --------------------------------------------------------------------------------------
def icp_forward(source: torch.Tensor, target: torch.Tensor, T_init: torch.Tensor,
                w_init: torch.Tensor, cfg: IcpSettings,
                replay_idx: Optional[List[torch.Tensor]] = None,
                record_idx: Optional[List[torch.Tensor]] = None) -> Dict:
    """Dense-tensor ICP loop: source (N,n,3), target (N,m,3|6), T_init (N,4,4),
    w_init (N,n) for pt2pl or (N,3n) for pt2pt (already x3-repeated, ICP.py:508-509).

    `replay_idx[k]` (N,n) pins the correspondences of iteration k (gradient parity is
    only meaningful with pinned correspondences); `record_idx`, if a
    list, receives the indices used.  Returns the reference's result dict
    (ICP.py:294-303) plus 'iters_run'."""
    assert cfg.dim in (2, 3), "dim must be 2 or 3"                       # ICP.py:79
    assert source.dtype == target.dtype == T_init.dtype                  # ICP.py:96
    N, n = source.shape[0], source.shape[1]
    dt, dev = source.dtype, source.device
    pt2pl = cfg.icp_type == "pt2pl"
    if pt2pl:
        assert target.shape[2] == 6                                       # ICP.py:103
    else:
        target = target[:, :, :3]                                         # ICP.py:105

    if cfg.dim == 2:                                                      # ICP.py:107-116
        # planar problem: z of the points and of the normals is dropped (copied out, not
        # multiplied, so a NaN/inf z cannot leak and z receives exactly zero gradient)
        zs = torch.zeros((N, n, 1), dtype=dt, device=dev)
        source = torch.cat((source[:, :, :2], zs), dim=2)
        zt = torch.zeros((N, target.shape[1], 1), dtype=dt, device=dev)
        if pt2pl:
            target = torch.cat((target[:, :, :2], zt, target[:, :, 3:5], zt), dim=2)
        else:
            target = torch.cat((target[:, :, :2], zt), dim=2)

    C = T_init[:, 0:3, 0:3]                                               # ICP.py:125
    r = T_init[:, 0:3, 3:]                                                # ICP.py:126
    P = source.transpose(1, 2)                                            # (N,3,n)  ICP.py:129

    converged = torch.zeros(N, dtype=torch.bool, device=dev)
    num_iters = torch.zeros(N, dtype=dt, device=dev)
    match_ratio = torch.zeros(N, dtype=dt, device=dev)
    deltas: List[torch.Tensor] = []
    weights: List[torch.Tensor] = []
    costs: List[torch.Tensor] = []
    ncols = 3 if cfg.dim == 2 else 6
    eye_reg = 1e-12 * torch.eye(ncols, dtype=dt, device=dev)
    use_trim = cfg.trim_dist is not None and cfg.trim_dist >= 0.0        # ICP.py:153
    k_run = 0
    w = w_init

    for k in range(cfg.max_iterations):                                  # ICP.py:132
        k_run = k + 1
        A_pts = C @ P                                                     # rotated, not translated (ICP.py:173)
        X = A_pts + r                                                     # ICP.py:137
        Xr = X.transpose(1, 2)                                            # (N,n,3)
        forced = replay_idx[k] if replay_idx is not None else None
        nbr, idx = find_nn(Xr, target, forced)                            # ICP.py:140
        if record_idx is not None:
            record_idx.append(idx.clone())

        e3 = Xr - nbr[:, :, :3]                                           # ICP.py:144/148
        if pt2pl:
            nrm = nbr[:, :, 3:]
            err = (e3 * nrm).sum(dim=2, keepdim=True)                     # ICP.py:146
        else:
            err = e3

        t_w = trim_weight(e3, cfg.trim_dist, cfg.differentiable, cfg.tanh_steepness) if use_trim \
            else torch.ones((N, n), dtype=dt, device=dev)                 # ICP.py:152-155
        l_w = robust_weight(cfg.loss_name, err, cfg.loss_metric, cfg.differentiable) \
            if cfg.loss_name is not None else torch.ones((N, n), dtype=dt, device=dev)  # ICP.py:157-160
        if not pt2pl:                                                     # ICP.py:162-166
            t_w = t_w.repeat_interleave(3, dim=1)
            l_w = l_w.repeat_interleave(3, dim=1)
            err = err.reshape(N, -1, 1)
        w = w_init * t_w * l_w                                            # ICP.py:169

        a = A_pts.transpose(1, 2)                                         # (N,n,3)
        if pt2pl:                                                         # ICP.py:172-176
            J_rot = torch.cross(nrm, a, dim=2)                            # skew(a)^T n = n x a
            J_tr = -nrm
        else:                                                             # ICP.py:177-180
            J_rot = _hat(a).reshape(N, -1, 3)
            J_tr = -torch.eye(3, device=dev).repeat(N, n, 1)
        J = torch.cat((J_rot, J_tr), dim=2)                               # ICP.py:183
        if cfg.dim == 2:                                                  # ICP.py:186-189
            J = J[:, :, 2:5]

        g = torch.sqrt(w + 1.0e-10) - 1.0e-5                              # ICP.py:194
        err_w = g.unsqueeze(-1) * err                                     # ICP.py:195
        J_w = g.unsqueeze(-1) * J                                         # ICP.py:196
        J_wT = J_w.transpose(1, 2)
        A = J_wT @ J_w + eye_reg                                          # ICP.py:200
        step = -torch.linalg.inv(A) @ J_wT @ err_w                        # ICP.py:201 (left-assoc)
        if cfg.dim == 2:                                                  # ICP.py:204-207
            step = torch.cat((torch.zeros((N, 2, 1), dtype=dt, device=dev), step,
                              torch.zeros((N, 1, 1), dtype=dt, device=dev)), dim=1)

        dC = torch.matrix_exp(_hat(step[:, 0:3, 0]))                      # ICP.py:210
        C = dC.transpose(1, 2) @ C                                        # ICP.py:214
        r = r - step[:, 3:6]                                              # ICP.py:216

        deltas.append(step.detach())                                      # ICP.py:220
        w_keep = w.detach()
        if weights:                                                       # ICP.py:224-226
            dead = (w_keep.sum(dim=1) == 0.0).unsqueeze(-1)
            w_keep = torch.where(dead, weights[-1].squeeze(-1), w_keep)
        weights.append(w_keep.unsqueeze(-1))                              # ICP.py:227
        cost = (err_w.transpose(1, 2) @ err_w).detach()                   # ICP.py:229
        if costs:                                                         # ICP.py:231-232
            cost = torch.where(cost == 0.0, costs[-1], cost)
        costs.append(cost)

        step_norm = torch.linalg.norm(step, dim=1).detach().squeeze(-1)   # ICP.py:237
        below = step_norm < cfg.tolerance
        converged = converged | below                                     # ICP.py:239
        if bool(below.any()) and not cfg.const_iter:                      # ICP.py:240
            first = below & (num_iters == 0)
            num_iters = torch.where(first, torch.full_like(num_iters, k + 1), num_iters)   # ICP.py:245
            matched = _count_gt(w, cfg.match_ratio_thresh)                # ICP.py:247
            at_start = _count_gt(w_init, cfg.match_ratio_thresh)          # ICP.py:248
            at_start = torch.where(at_start == 0, torch.ones_like(at_start), at_start)   # ICP.py:249
            match_ratio = torch.where(below & (match_ratio == 0), match_ratio + matched / at_start,
                                      match_ratio)                        # ICP.py:251
            w_init = w_init * (~below).to(dt).unsqueeze(-1)               # ICP.py:256-257
            if bool(below.all()):                                         # ICP.py:259
                break

    num_iters = torch.where(num_iters == 0, torch.full_like(num_iters, k_run), num_iters)  # ICP.py:267
    matched = _count_gt(w, cfg.match_ratio_thresh)                        # ICP.py:268
    at_start = _count_gt(w_init, cfg.match_ratio_thresh)                  # ICP.py:269
    at_start = torch.where(at_start == 0, torch.ones_like(at_start), at_start)
    match_ratio = torch.where(match_ratio == 0, (matched / at_start).to(dt), match_ratio)  # ICP.py:271

    pc = (C @ P + r).transpose(1, 2)                                      # ICP.py:274
    T = torch.diag_embed(torch.ones((N, 4), dtype=dt, device=dev))        # ICP.py:277-280
    T[:, 0:3, 0:3] = C
    T[:, 0:3, 3] = r.squeeze(-1)

    return {
        "pc": pc,
        "T": T,
        "costs": torch.stack(costs, dim=1).squeeze(-1),                   # (N,K,1)   ICP.py:285
        "deltas": torch.stack(deltas, dim=1),                             # (N,K,6,1) ICP.py:283
        "weights": torch.stack(weights, dim=1),                           # (N,K,n|3n,1) ICP.py:284
        "stats": {"converged": converged, "iterations": num_iters, "matched_ratio": match_ratio},
        "iters_run": k_run,
    }


# --------------------------------------------------------------------------------------
# Kabsch / SVD point-to-point ICP  (ICP.py:533-591, "not yet integrated" upstream)
# --------------------------------------------------------------------------------------
def kabsch_icp(source: torch.Tensor, target: torch.Tensor, T_init: torch.Tensor, max_iterations: int,
               tolerance: float, literal: bool = False):
    """Single pair: source (n,>=3), target (m,>=3), T_init (4,4) -> (moved source (n,3), T (4,4), iterations).

    Restates ICP.py:545-591 step by step.  Two quirks of the reference are kept: T_init is only
    composed into the returned transform, it is NOT applied to the source before the first search
    (ICP.py:545-547), and convergence compares the moved points with the neighbours found before
    the move (ICP.py:585).  One is NOT kept unless `literal=True`: ICP.py:566-570 names the third
    output of `torch.svd` "Vt" although it is V, and forms U diag(1,1,det) V; that equals the
    Kabsch rotation U diag(1,1,det) V^T only when V happens to be symmetric (it is on the planar
    test fixture, where the reference reaches its known answer; on a general 3-D cloud it does
    not).  The default here is the Kabsch rotation."""
    ps = source[:, 0:3]
    tgt = target[:, 0:3]
    T = T_init
    n = ps.shape[0]
    it = 0
    for it in range(1, max_iterations + 1):
        idx = nn_index_reference(ps.detach()[None], tgt.detach()[None])[0]            # ICP.py:553-555 (per-point loop)
        nn_t = tgt[idx]
        p_mean = ps.mean(dim=0).reshape(3, 1)                                          # ICP.py:558
        q_mean = nn_t.mean(dim=0).reshape(3, 1)                                        # ICP.py:559
        W = (nn_t.T - q_mean) @ (ps.T - p_mean).T / n                                  # ICP.py:562-563
        U, _, V = torch.svd(W)                                                         # ICP.py:566
        d = torch.det(U) * torch.det(V)
        D = torch.diag(torch.stack((torch.ones_like(d), torch.ones_like(d), d)))
        C = U @ D @ (V if literal else V.T)                                            # ICP.py:570
        r = q_mean - C @ p_mean                                                        # ICP.py:573
        T_up = torch.eye(4, dtype=ps.dtype)
        T_up = torch.cat((torch.cat((C, r), dim=1), T_up[3:4]), dim=0)                 # ICP.py:577
        T = T_up @ T                                                                   # ICP.py:578
        ps = (C @ ps.T + r).T                                                          # ICP.py:581
        if float(torch.sum((ps - nn_t).detach() ** 2)) < tolerance:                           # ICP.py:585
            break
    return ps, T, it


# --------------------------------------------------------------------------------------
# synthetic workload generator  (used by tests and bench)
# --------------------------------------------------------------------------------------
def make_wavy_pair(N: int, n: int, m: int, seed: int, dtype=torch.float32, with_normals: bool = True,
                   noise: float = 0.02, yaw: float = 0.02, shift=(0.3, -0.2, 0.05), extent: float = 20.0):
    """Seeded wavy-surface cloud pairs.  target (N,m,6|3): xy~U(-extent,extent)^2,
    z = 0.5 sin(0.3x) + 0.5 cos(0.2y), analytic unit normals; source (N,n,3): a random
    subset/permutation of the target points + N(0,noise^2), moved by Rz(yaw) and `shift`.
    Generated in fp64 then cast."""
    gen = torch.Generator().manual_seed(seed)
    xy = (torch.rand((N, m, 2), generator=gen, dtype=torch.float64) * 2.0 - 1.0) * extent
    x, y = xy[..., 0], xy[..., 1]
    z = 0.5 * torch.sin(0.3 * x) + 0.5 * torch.cos(0.2 * y)
    nrm = torch.stack((-0.15 * torch.cos(0.3 * x), 0.1 * torch.sin(0.2 * y), torch.ones_like(x)), dim=-1)
    nrm = nrm / torch.linalg.norm(nrm, dim=-1, keepdim=True)
    pts = torch.stack((x, y, z), dim=-1)
    tgt = torch.cat((pts, nrm), dim=-1) if with_normals else pts
    src = torch.empty((N, n, 3), dtype=torch.float64)
    cy, sy = math.cos(yaw), math.sin(yaw)
    Rz = torch.tensor([[cy, -sy, 0.0], [sy, cy, 0.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
    t = torch.tensor(shift, dtype=torch.float64)
    for b in range(N):
        perm = torch.randperm(m, generator=gen)
        if n <= m:
            pick = perm[:n]
        else:
            pick = torch.cat([perm] * (n // m + 1))[:n]
        p = pts[b, pick] + noise * torch.randn((n, 3), generator=gen, dtype=torch.float64)
        src[b] = p @ Rz.T + t
    return src.to(dtype), tgt.to(dtype)
