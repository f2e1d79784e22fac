"""Robust / trim weights — host-side mirror of `dICP/loss.py:3-58`.

Inside `ICP.icp` these weights are evaluated by the fused Gauss-Newton kernel
(csrc/dicp_math.cuh: trim_weight / loss_weight); this class keeps the reference's public
`loss(name, metric, differentiable, tanh_steepness).get_weight(err)` for callers that use it
directly.  It is plain tensor math on whatever device `err` lives on."""
import torch


class loss:
    def __init__(self, name="huber", metric=1.0, differentiable=False, tanh_steepness=10.0):
        self.name = name                        # "huber" | "cauchy" | "trim"
        self.metric = metric
        self.differentiable = differentiable
        self.tanh_steepness = tanh_steepness

    def get_weight(self, err):
        """err (n,k) or (N,n,k) -> weights (n,) or (N,n); the norm is over the last axis."""
        if self.name not in ("huber", "cauchy", "trim"):
            raise ValueError("Invalid loss name: {}".format(self.name))
        mag = torch.linalg.norm(err, dim=-1)
        c = self.metric
        if self.name == "cauchy":
            return 1.0 / (1.0 + (mag / c) ** 2)
        if self.name == "huber":
            if self.differentiable:
                return c ** 2 / (c ** 2 + mag ** 2)
            return torch.where(mag > c, c / mag, torch.ones_like(mag))
        if self.differentiable:
            return 0.5 * torch.tanh(self.tanh_steepness * (c - mag) - 3.0) + 0.5
        # (for a 2-D `err` the reference's hard trim broadcasts an (n,) mask against (n,1)
        # constants and returns (n,n), loss.py:46-58; the (n,) answer is returned here)
        return torch.where(mag < c, torch.ones_like(mag), torch.zeros_like(mag))
