"""Hand-derived reverse pass of the dICP loop.  TEST INFRASTRUCTURE ONLY (see
`oracle/dicp_oracle.py` for the import rules).

The reference gets its gradients from torch autograd through the unrolled loop
(`dICP/ICP.py:132-280`).  The CUDA backward kernels cannot use
autograd, so the adjoint of one Gauss-Newton iteration is written out here, in the same
per-point form the kernels use, and checked against autograd of `icp_forward`
(`tests/test_adjoint.py`).  It is the executable specification of
`dicp_b200/csrc/bwd_kernels.cu`.

Notation (one cloud pair, iteration k):
  a_i = C p_i, x_i = a_i + r, q_i = y[idx_i,:3], n_i = y[idx_i,3:6], e_i = x_i - q_i,
  rho_i = n_i.e_i (pt2pl residual; pt2pt residual is e_i), d_i = |e_i|,
  t_i = trim(d_i), l_i = loss(|rho_i| or d_i), w_i = w0_i t_i l_i,
  g_i = sqrt(w_i + 1e-10) - 1e-5, u_i = g_i^2,
  pt2pl J_i = [n_i x a_i ; -n_i] (6), pt2pt J_i = [hat(a_i) | -I] (3x6),
  A = sum u_i J_i^T J_i + 1e-12 I, b = sum u_i J_i^T res_i, delta = -A^-1 b = [phi; rho_t],
  R = exp(hat(phi)), C+ = R^T C, r+ = r - rho_t.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from .dicp_oracle import IcpSettings


def _hat(v):
    z = torch.zeros_like(v[..., 0])
    return torch.stack([z, -v[..., 2], v[..., 1], v[..., 2], z, -v[..., 0], -v[..., 1], v[..., 0], z],
                       dim=-1).reshape(*v.shape[:-1], 3, 3)


def _axial(G):
    """<G, hat(v)> = v . axial(G)."""
    return torch.stack([G[..., 2, 1] - G[..., 1, 2], G[..., 0, 2] - G[..., 2, 0], G[..., 1, 0] - G[..., 0, 1]], dim=-1)


def so3_exp(phi):
    """Rodrigues; phi (N,3) -> (N,3,3).  (`torch.matrix_exp`, ICP.py:210)"""
    th2 = (phi * phi).sum(-1)
    th = torch.sqrt(th2)
    small = th < 1e-4
    ths = torch.where(small, torch.ones_like(th), th)
    A = torch.where(small, 1 - th2 / 6 + th2 * th2 / 120, torch.sin(ths) / ths)
    B = torch.where(small, 0.5 - th2 / 24 + th2 * th2 / 720, (1 - torch.cos(ths)) / (ths * ths))
    K = _hat(phi)
    eye = torch.eye(3, dtype=phi.dtype).expand_as(K)
    return eye + A[:, None, None] * K + B[:, None, None] * (K @ K)


def so3_left_jacobian(phi):
    """J_l(phi) with d exp(hat(phi)) = hat(J_l dphi) exp(hat(phi)).  (N,3) -> (N,3,3)"""
    th2 = (phi * phi).sum(-1)
    th = torch.sqrt(th2)
    small = th < 1e-4
    ths = torch.where(small, torch.ones_like(th), th)
    B = torch.where(small, 0.5 - th2 / 24 + th2 * th2 / 720, (1 - torch.cos(ths)) / (ths * ths))
    Cc = torch.where(small, 1.0 / 6 - th2 / 120 + th2 * th2 / 5040, (ths - torch.sin(ths)) / (ths * ths * ths))
    K = _hat(phi)
    eye = torch.eye(3, dtype=phi.dtype).expand_as(K)
    return eye + B[:, None, None] * K + Cc[:, None, None] * (K @ K)


def _trim(d, cfg: IcpSettings):
    """value and derivative w.r.t. d (loss.py:43-58)."""
    if cfg.trim_dist is None or cfg.trim_dist < 0:
        return torch.ones_like(d), torch.zeros_like(d)
    if cfg.differentiable:
        th = torch.tanh(cfg.tanh_steepness * (cfg.trim_dist - d) - 3.0)
        return 0.5 * th + 0.5, -0.5 * cfg.tanh_steepness * (1 - th * th)
    return (d < cfg.trim_dist).to(d.dtype), torch.zeros_like(d)


def _loss(z, cfg: IcpSettings):
    """value and derivative w.r.t. z = |residual| >= 0 (loss.py:21-41)."""
    if cfg.loss_name is None:
        return torch.ones_like(z), torch.zeros_like(z)
    c = cfg.loss_metric
    if cfg.loss_name == "cauchy" or (cfg.loss_name == "huber" and cfg.differentiable):
        l = 1.0 / (1.0 + (z / c) ** 2) if cfg.loss_name == "cauchy" else c ** 2 / (c ** 2 + z ** 2)
        return l, -2.0 * z / (c * c) * l * l
    if cfg.loss_name == "huber":
        big = z > c
        zs = torch.where(big, z, torch.ones_like(z))
        return torch.where(big, c / zs, torch.ones_like(z)), torch.where(big, -c / (zs * zs), torch.zeros_like(z))
    raise ValueError(cfg.loss_name)


def icp_forward_saved(source, target, T_init, w0, cfg: IcpSettings, idx_list: List[torch.Tensor]) -> Dict:
    """Forward with pinned correspondences, const-iteration semantics, keeping exactly
    the state the CUDA forward saves: C_k, r_k, A_k, delta_k (per iteration) and idx_k.
    w0 is per point (N,n) here (pt2pt repeats it internally).  Plain tensor math, no autograd."""
    N, n, _ = source.shape
    pt2pl = cfg.icp_type == "pt2pl"
    cols = [2, 3, 4] if cfg.dim == 2 else [0, 1, 2, 3, 4, 5]
    C = T_init[:, :3, :3].clone()
    r = T_init[:, :3, 3].clone()
    saved = {"C": [], "r": [], "A": [], "delta": []}
    for k in range(len(idx_list)):
        idx = idx_list[k].long()
        a = source @ C.transpose(1, 2)
        x = a + r[:, None, :]
        nbr = torch.gather(target, 1, idx.unsqueeze(-1).expand(-1, -1, target.shape[-1]))
        e = x - nbr[:, :, :3]
        d = torch.linalg.norm(e, dim=-1)
        t, _ = _trim(d, cfg)
        if pt2pl:
            nrm = nbr[:, :, 3:6]
            rho = (e * nrm).sum(-1)
            l, _ = _loss(rho.abs(), cfg)
            J = torch.cat((torch.cross(nrm, a, dim=-1), -nrm), dim=-1)            # (N,n,6)
            res = rho.unsqueeze(-1)                                               # (N,n,1)
            Jm = J.unsqueeze(2)                                                   # (N,n,1,6)
        else:
            l, _ = _loss(d, cfg)
            Jm = torch.cat((_hat(a), -torch.eye(3, dtype=a.dtype).expand(N, n, 3, 3)), dim=-1)   # (N,n,3,6)
            res = e
        w = w0 * t * l
        g = torch.sqrt(w + 1e-10) - 1e-5
        u = g * g
        Jc = Jm[..., cols]
        A = torch.einsum("bn,bnki,bnkj->bij", u, Jc, Jc) + 1e-12 * torch.eye(len(cols), dtype=a.dtype)
        b = torch.einsum("bn,bnki,bnk->bi", u, Jc, res)
        dl = -torch.linalg.solve(A, b.unsqueeze(-1)).squeeze(-1)
        delta = torch.zeros((N, 6), dtype=a.dtype)
        delta[:, cols] = dl
        saved["C"].append(C.clone()); saved["r"].append(r.clone()); saved["A"].append(A); saved["delta"].append(delta)
        R = so3_exp(delta[:, :3])
        C = R.transpose(1, 2) @ C
        r = r - delta[:, 3:]
    saved["C_final"], saved["r_final"] = C, r
    T = torch.eye(4, dtype=source.dtype).repeat(N, 1, 1)
    T[:, :3, :3] = C
    T[:, :3, 3] = r
    pc = source @ C.transpose(1, 2) + r[:, None, :]
    return {"T": T, "pc": pc, "saved": saved}


def icp_backward(source, target, w0, cfg: IcpSettings, idx_list, saved, grad_T, grad_pc) -> Dict:
    """Explicit reverse pass.  Returns grads for source (N,n,3), target (N,m,S),
    w0 (N,n) and T_init (N,4,4)."""
    N, n, _ = source.shape
    pt2pl = cfg.icp_type == "pt2pl"
    cols = [2, 3, 4] if cfg.dim == 2 else [0, 1, 2, 3, 4, 5]
    g_src = torch.zeros_like(source)
    g_tgt = torch.zeros_like(target)
    g_w0 = torch.zeros_like(w0)

    # seeds from pc = C_K p + r_K and T = [C_K r_K; 0 1]              (ICP.py:274-280)
    Cbar = grad_T[:, :3, :3] + torch.einsum("bni,bnj->bij", grad_pc, source)
    rbar = grad_T[:, :3, 3] + grad_pc.sum(1)
    g_src += grad_pc @ saved["C_final"]                                # C^T pbar, row form

    for k in reversed(range(len(idx_list))):
        idx = idx_list[k].long()
        C, r, A, delta = saved["C"][k], saved["r"][k], saved["A"][k], saved["delta"][k]
        phi = delta[:, :3]
        R = so3_exp(phi)
        # update step: C+ = R^T C ; r+ = r - rho_t                      (ICP.py:214-216)
        Rbar = C @ Cbar.transpose(1, 2)
        Cbar = R @ Cbar
        phibar = (so3_left_jacobian(phi).transpose(1, 2) @ _axial(Rbar @ R.transpose(1, 2)).unsqueeze(-1)).squeeze(-1)
        dbar = torch.cat((phibar, -rbar), dim=-1)                       # rbar passes through unchanged
        lam_c = torch.linalg.solve(A, dbar[:, cols].unsqueeze(-1)).squeeze(-1)
        lam = torch.zeros((N, 6), dtype=source.dtype)
        lam[:, cols] = lam_c

        # per-point recomputation
        a = source @ C.transpose(1, 2)
        x = a + r[:, None, :]
        nbr = torch.gather(target, 1, idx.unsqueeze(-1).expand(-1, -1, target.shape[-1]))
        e = x - nbr[:, :, :3]
        d = torch.linalg.norm(e, dim=-1)
        t, dt_dd = _trim(d, cfg)
        lam_b, del_b = lam[:, None, :], delta[:, None, :]
        if pt2pl:
            nrm = nbr[:, :, 3:6]
            rho = (e * nrm).sum(-1)
            l, dl_dz = _loss(rho.abs(), cfg)
            J = torch.cat((torch.cross(nrm, a, dim=-1), -nrm), dim=-1)
            alpha = (J * lam_b).sum(-1)
            beta = (J * del_b).sum(-1)
        else:
            l, dl_dz = _loss(d, cfg)
            # J_i v = hat(a) v_c - v_r = a x v_c - v_r
            alpha = torch.cross(a, lam_b[..., :3].expand_as(a), dim=-1) - lam_b[..., 3:]
            beta = torch.cross(a, del_b[..., :3].expand_as(a), dim=-1) - del_b[..., 3:]
        w = w0 * t * l
        sq = torch.sqrt(w + 1e-10)
        g = sq - 1e-5
        u = g * g
        if pt2pl:
            ubar = -alpha * (beta + rho)
            rhobar = -u * alpha
            Jbar = -u.unsqueeze(-1) * ((beta + rho).unsqueeze(-1) * lam_b + alpha.unsqueeze(-1) * del_b)   # (N,n,6)
        else:
            ubar = -(alpha * (beta + e)).sum(-1)
            ebar = -u.unsqueeze(-1) * alpha
        wbar = ubar * g / sq
        g_w0 += wbar * t * l
        tbar = wbar * w0 * l
        lbar = wbar * w0 * t
        dbar_pt = tbar * dt_dd
        if pt2pl:
            rhobar = rhobar + lbar * dl_dz * torch.sign(rho)
            ebar = rhobar.unsqueeze(-1) * nrm
            nbar = rhobar.unsqueeze(-1) * e + torch.cross(a, Jbar[..., :3], dim=-1) - Jbar[..., 3:]
            abar = torch.cross(Jbar[..., :3], nrm, dim=-1)
        else:
            dbar_pt = dbar_pt + lbar * dl_dz
            be = beta + e
            abar = u.unsqueeze(-1) * (torch.cross(be, lam_b[..., :3].expand_as(a), dim=-1)
                                      + torch.cross(alpha, del_b[..., :3].expand_as(a), dim=-1))
        dsafe = torch.where(d > 0, d, torch.ones_like(d))
        ebar = ebar + torch.where(d > 0, dbar_pt / dsafe, torch.zeros_like(d)).unsqueeze(-1) * e
        abar = abar + ebar
        rbar = rbar + ebar.sum(1)
        Cbar = Cbar + torch.einsum("bni,bnj->bij", abar, source)
        g_src += abar @ C
        scat = -ebar if not pt2pl else torch.cat((-ebar, nbar), dim=-1)
        g_tgt.scatter_add_(1, idx.unsqueeze(-1).expand(-1, -1, scat.shape[-1]), scat)

    g_T = torch.zeros((N, 4, 4), dtype=source.dtype)
    g_T[:, :3, :3] = Cbar
    g_T[:, :3, 3] = rbar
    return {"source": g_src, "target": g_tgt, "w0": g_w0, "T_init": g_T}
