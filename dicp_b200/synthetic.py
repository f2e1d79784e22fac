"""Synthetic benchmark inputs: seeded cloud pairs on a wavy surface with analytic
normals.  Used by bench.py and the tools; no reference or oracle code is involved."""
import math

import torch


def surface_pairs(n_pairs: int, n: int, m: int, seed: int, dtype=torch.float32, with_normals: bool = True,
                  noise: float = 0.02, yaw: float = 0.02, shift=(0.3, -0.2, 0.05), extent: float = 20.0):
    """-> source (n_pairs, n, 3), target (n_pairs, m, 6 | 3), generated in float64 on the CPU and cast.

    Target: xy uniform in [-extent, extent]^2, z = 0.5 sin(0.3 x) + 0.5 cos(0.2 y), unit normals of that
    surface.  Source: n target points of a random permutation (wrapping around when n > m) with
    N(0, noise^2) added per coordinate, rotated by `yaw` about z and translated by `shift`."""
    g = torch.Generator().manual_seed(seed)
    f64 = torch.float64
    xy = extent * (2.0 * torch.rand((n_pairs, m, 2), generator=g, dtype=f64) - 1.0)
    px, py = xy.unbind(-1)
    pz = 0.5 * torch.sin(0.3 * px) + 0.5 * torch.cos(0.2 * py)
    points = torch.stack((px, py, pz), dim=-1)
    if with_normals:
        grad = torch.stack((-0.15 * torch.cos(0.3 * px), 0.1 * torch.sin(0.2 * py), torch.ones_like(px)), dim=-1)
        target = torch.cat((points, grad / grad.norm(dim=-1, keepdim=True)), dim=-1)
    else:
        target = points
    c, s = math.cos(yaw), math.sin(yaw)
    rot = torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], dtype=f64)
    move = torch.tensor(shift, dtype=f64)
    source = torch.empty((n_pairs, n, 3), dtype=f64)
    for b in range(n_pairs):
        order = torch.randperm(m, generator=g)
        if n > m:
            order = order.repeat(n // m + 1)
        noisy = points[b, order[:n]] + noise * torch.randn((n, 3), generator=g, dtype=f64)
        source[b] = noisy @ rot.T + move
    return source.to(dtype), target.to(dtype)
