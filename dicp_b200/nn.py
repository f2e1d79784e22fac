"""Nearest-neighbour lookup — host-side mirror of `dICP/nn.py:4-125`.

`find_nn(x, y)` returns, for every query row of x, the nearest row of y (with its normals
if y has 6 columns).  The hard search (`differentiable=False`, or `use_gumbel=False`) runs
the brute-force CUDA kernel `dicp_nn_argmin_*` and a gather whose backward scatter-adds
into y — the same gradient path the reference has (nn.py:37-38).  Inputs on the CPU are
staged through the GPU; without a GPU this raises (no CPU fallback).

The Gumbel-softmax variant (nn.py:43-70, "not yet integrated" upstream, off in the shipped
config; the constructor default `use_gumbel=True` selects it) runs as a streaming online-softmax
kernel (`dicp_gumbel_nn_*`, O(n+m) memory instead of the reference's dense (N,n,m) tensor) with
a counter-based noise generator seeded from torch's default generator, so `torch.manual_seed`
makes it reproducible.  It is a different random stream than `torch.rand`, hence parity is
defined against the same formula evaluated on the kernel's own noise (tests)."""
import torch

from . import _functions


class nn:
    def __init__(self, differentiable=True, use_gumbel=True, eps=1e-20, tau=0.1):
        self.differentiable = differentiable
        self.use_gumbel = use_gumbel
        self.eps = eps
        self.tau = tau

    def find_nn(self, x, y):
        xq, yt = self._canonical(x, y)
        if self.differentiable and self.use_gumbel:
            return self._soft_nn(xq, yt)
        return self._hard_nn(xq, yt)

    # -- hard NN: CUDA kernels --------------------------------------------------------
    @staticmethod
    def _hard_nn(x, y):
        dev = x.device
        if not x.is_cuda:
            _functions.require_cuda()
            x, y = x.cuda(), y.cuda()
        if y.dtype != x.dtype:
            y = y.to(x.dtype)
        idx = _functions.nn_argmin(x, y)
        out = _functions.GatherRows.apply(y, idx)
        return out if out.device == dev else out.to(dev)

    # -- Gumbel-softmax NN (nn.py:43-70): streaming kernel --------------------------------
    def _soft_nn(self, x, y):
        dev = x.device
        if not x.is_cuda:
            _functions.require_cuda()
            x, y = x.cuda(), y.cuda()
        if y.dtype != x.dtype:
            y = y.to(x.dtype)
        seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())      # torch.manual_seed controls it
        self.last_seed = seed
        out = _functions.GumbelNN.apply(x, y, seed, self.eps, self.tau)
        return out if out.device == dev else out.to(dev)

    # -- layout normalisation (nn.py:94-125) --------------------------------------------
    @staticmethod
    def _canonical(x, y):
        """-> x (N,n,3), y (N,m,3|6).  Accepts (n,3), (N,n,3) and channel-first (N,3|6,n)."""
        xq = x.unsqueeze(0) if x.dim() == 2 else x
        if xq.shape[-2] == 3 or (xq.shape[-2] == 6 and xq.shape[-2] < xq.shape[-1]):
            xq = xq[:, :3, :].transpose(1, 2)
        assert xq.shape[2] == 3, "x must have 3 elements in the second dimension."
        yt = y.unsqueeze(0) if y.dim() == 2 else y
        if yt.shape[-2] == 3 or (yt.shape[-2] == 6 and yt.shape[-2] < yt.shape[-1]):
            yt = yt.transpose(1, 2)
        assert yt.shape[2] == 3 or yt.shape[2] == 6, "y must have 3 or 6 elements in the second dimension."
        return xq, yt
