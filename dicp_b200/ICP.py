"""`ICP` — host-side mirror of the reference's public class (`dICP/ICP.py:14-511`).

Same constructor, same `icp` / `dICP` / `batch_size_handling` signatures, same YAML keys,
same result dictionary.  What differs is where the work happens: after the (host-side,
once-per-call) batch normalisation, the whole Gauss-Newton loop of `ICP.py:125-303` —
transform, nearest neighbour, residual, trim / robust weights, Jacobian, normal equations,
solve, exp-map update, convergence bookkeeping — runs as hand-written sm_100a kernels
behind one `torch.autograd.Function` (`_functions.ICPFunction`), and so does its reverse
pass.  No CPU fallback: without a CUDA device `icp()` raises.
"""
import os.path as osp

import torch
import yaml

from . import _functions
from .nn import nn


def _is_empty(x):
    return x is None or len(x) == 0


class _Results(dict):
    """Opt-in (`ICP.lazy_host_results = True`) result dictionary of `ICP.icp` (ICP.py:294-303) for host inputs: the
    outputs are produced on the GPU and each entry is copied back to the host the first time it is read through
    `[]`, `get`, `items` or `values`, so a caller that only wants `T` does not pay for the (N,K,n,1) `weights`
    history.  (Bulk paths of the dict type -- `dict(res)`, `copy`, pickling -- bypass these hooks, hence opt-in.)"""

    def _force(self, key):
        v = dict.__getitem__(self, key)
        if isinstance(v, _Deferred):
            v = v()
            dict.__setitem__(self, key, v)
        return v

    def __getitem__(self, key):
        return self._force(key)

    def get(self, key, default=None):
        return self._force(key) if key in self else default

    def items(self):
        return [(k, self._force(k)) for k in self.keys()]

    def values(self):
        return [self._force(k) for k in self.keys()]


# pairs per kernel-side call (gridDim.y); larger batches are split by ICP.dICP
_MAX_PAIRS_PER_CALL = 65535


class _Deferred:
    def __init__(self, fn):
        self.fn = fn

    def __call__(self):
        return self.fn()


class ICP:
    def __init__(self, config_path=None, icp_type='pt2pl', max_iterations=100, tolerance=1e-12, differentiable=True):
        if config_path is None:
            # same default location as the reference: <package>/../config/dICP_config.yaml (ICP.py:17-19)
            config_path = osp.join(osp.dirname(osp.abspath(__file__)), '..', 'config', 'dICP_config.yaml')
        with open(config_path, 'r') as f:
            self.config = yaml.safe_load(f)

        params = self.config['dICP']['parameters']
        logging = self.config['dICP']['logging']
        func = self.config['dICP']['functionality']
        self.icp_type = icp_type
        self.max_iterations = max_iterations
        self.tolerance = tolerance
        self.const_iter = params['const_iter']
        self.verbose = logging['verbose']
        self.target_pad_val = params['target_pad_val']
        self.source_zeroes_are_pad = params['source_zeroes_are_pad']
        self.match_ratio_thresh = logging['matched_ratio_thresh']
        self.diff = differentiable
        self.nn = nn(self.diff, use_gumbel=func['gumbel'], eps=func['gumbel_eps'], tau=func['gumbel_tau'])
        # not in the reference: how often the host polls the device-side "all converged" flag
        # when const_iter is False (the reference synchronises every iteration, ICP.py:240)
        self.check_every = 1
        # not in the reference: CUDA device used when the inputs live on the CPU
        self.device = None
        # not in the reference: 'brute' = tiled brute-force scan (default), 'grid' = exact search through a
        # uniform grid built once per call (same correspondences, O(n) per iteration); optional YAML key
        # dICP.functionality.nn_search
        self.nn_search = func.get('nn_search', 'brute')
        # not in the reference: when True, every call leaves the correspondences it used in
        # `self.last_correspondences`, an int32 (K, N, n) CUDA tensor (iteration, pair, source point) -> target row.
        # Parity tests replay them through the fp64 oracle (gradients are only comparable on equal correspondences)
        self.keep_correspondences = False
        self.last_correspondences = None
        # optional YAML key dICP.functionality.svd (commented out upstream, config/dICP_config.yaml:15-16): point-to-point
        # problems are solved by the Kabsch / SVD loop `pt2pt_dICP_SVD` instead of Gauss-Newton
        # not in the reference: with host inputs, True returns a dict subclass whose large entries (pc, weights, deltas,
        # costs) are copied back from the GPU on first access instead of eagerly; the default is the plain dict upstream returns
        self.lazy_host_results = False
        self.use_svd = bool(func.get('svd', False))
        self.svd_literal = False
        # not in the reference (optional YAML key dICP.functionality.sync_free): with const_iter set, icp() only
        # enqueues its kernels and returns without waiting for the GPU -- the reference stops at every iteration
        # (ICP.py:240) and torch.linalg.inv raises at once on a singular matrix (ICP.py:201).  A singular normal matrix
        # then does not raise inside icp(); `check_status()` reports it for the most recent call.  Off by default.
        self.sync_free = bool(func.get('sync_free', False))
        self._last_status = None

    def icp(self, source, target, T_init, weight=None, trim_dist=None, loss_fn=None, dim=3):
        if self.use_svd and self.icp_type == 'pt2pt':
            return self._icp_svd(source, target, T_init)
        return self.dICP(source, target, T_init, weight, trim_dist, loss_fn, dim)

    def _icp_svd(self, source, target, T_init):
        """`functionality.svd`: the Kabsch / SVD solver behind the result dictionary of `icp()` (weights, trimming and
        robust losses do not exist on this path upstream either, ICP.py:533-591)."""
        source, target, T_init, _ = self.batch_size_handling(source, target, T_init, None)
        pc, T = self.pt2pt_dICP_SVD(source, target[:, :, :3], T_init)
        return {"pc": pc, "T": T}

    # ---------------------------------------------------------------------------------
    def dICP(self, source, target, T_init, weight=None, trim_dist=None, loss_fn=None, dim=3):
        """Point-to-point / point-to-plane ICP; argument and return conventions are those of
        the reference docstring (ICP.py:50-78): returns a dict with 'pc' (N,n,3), 'T' (N,4,4),
        'costs' (N,K,1), 'deltas' (N,K,6,1), 'weights' (N,K,n|3n,1) and 'stats'
        {'converged','iterations','matched_ratio'} (each (N,))."""
        assert dim == 2 or dim == 3, "dim must be 2 or 3"
        # ragged target lists: remember the real length of every cloud so that the padding rows
        # (ICP.py:471-479) can never be returned as a neighbour, whatever max(source) * target_pad_val is
        m_valid = None
        if isinstance(target, list) and len(target) > 0 and not _is_empty(source):
            m_valid = [max(len(t_i), 1) for t_i in target]
        # ragged source lists: the real length of every cloud, so that the kernels neither search nor evaluate the zero
        # padding (ICP.py:361-426 gives it weight 0, which is what skipping it amounts to; with a negative
        # matched_ratio_thresh zero-weight rows would count as matched, so then the padding is processed as upstream)
        n_valid = None
        if isinstance(source, list) and len(source) > 0 and not _is_empty(target) and self.match_ratio_thresh >= 0.0:
            n_valid = [max(len(s_i), 1) for s_i in source]
        default_prior = (weight is None and not self.source_zeroes_are_pad and not isinstance(source, list)
                         and not isinstance(target, list) and not _is_empty(source) and not _is_empty(target))
        source, target, T_init, w_init = self.batch_size_handling(source, target, T_init, weight)
        assert source.dtype == target.dtype == T_init.dtype
        N, n = source.shape[0], source.shape[1]
        pt2pl = self.icp_type == 'pt2pl'
        if pt2pl:
            assert target.shape[2] == 6          # normals are required (ICP.py:103)
        else:
            target = target[:, :, :3]
        if self.diff and self.nn.use_gumbel:
            # config functionality.gumbel ("not yet integrated" upstream, off in the shipped config): soft neighbours
            # from the streaming Gumbel-softmax kernel, gradients through the softmax; see _gumbel_loop
            return self._gumbel_loop(source, target, T_init, w_init, trim_dist, loss_fn, dim)

        if dim == 2:
            # planar problem: drop z of the points and of the normals (ICP.py:107-116)
            zs = source.new_zeros((N, n, 1))
            source = torch.cat((source[:, :, :2], zs), dim=2)
            zt = target.new_zeros((N, target.shape[1], 1))
            if pt2pl:
                target = torch.cat((target[:, :, :2], zt, target[:, :, 3:5], zt), dim=2)
            else:
                target = torch.cat((target[:, :, :2], zt), dim=2)

        # per-point prior weights; pt2pt carries them x3-repeated (ICP.py:508-509).  The all-ones default of a
        # dense call is not materialised on the device (the kernels treat a missing prior as 1)
        w_point = w_init if pt2pl else w_init[:, ::3]
        if default_prior:
            w_point = None

        out_dev = source.device
        dev = self._compute_device(out_dev)
        loss_name = loss_fn['name'] if loss_fn is not None else None
        loss_metric = loss_fn['metric'] if loss_fn is not None else 1.0
        params = _functions.make_params(
            self.icp_type, dim, self.diff, loss_name, loss_metric, trim_dist,
            self.config['dICP']['parameters']['tanh_steepness'], self.tolerance, self.max_iterations,
            self.const_iter, self.match_ratio_thresh, self.check_every,
            self.nn_search if isinstance(self.nn_search, str) else 'brute', self.sync_free)

        if m_valid is not None and (len(m_valid) != N or min(m_valid) == target.shape[1]):
            m_valid = None                                   # nothing is padded
        if m_valid is not None:
            m_valid = torch.tensor(m_valid, dtype=torch.int32, device=dev)
        if n_valid is not None and (len(n_valid) != N or min(n_valid) == n):
            n_valid = None                                   # nothing is padded
        if n_valid is not None:
            n_valid = torch.tensor(n_valid, dtype=torch.int32, device=dev)
        # host buffers are staged through the GPU; from pinned memory the copies are asynchronous and
        # stream-ordered in front of the kernels
        source_d, target_d = source.to(dev, non_blocking=True), target.to(dev, non_blocking=True)
        # nn_search may also be a search provider (e.g. parallel.SplitTargetSearch: the targets of one huge
        # pair split over the GPUs of a process group); it is bound to this call's clouds and then asked
        # for the correspondences of every iteration
        nn_hook = None if isinstance(self.nn_search, str) else self.nn_search.bind(source_d.detach(), target_d.detach(), m_valid)
        keep = [] if self.keep_correspondences else None
        status = []
        T_init_d = T_init.to(dev, non_blocking=True)
        w_d = w_point.to(dev, non_blocking=True) if w_point is not None else None
        if N <= _MAX_PAIRS_PER_CALL or nn_hook is not None:
            pc, T, deltas, weights, costs, converged, num_iters, match_ratio = _functions.ICPFunction.apply(
                source_d, target_d, T_init_d, w_d, params, m_valid, nn_hook, keep, n_valid, status)
        else:
            # the kernels index pairs with gridDim.y: a larger batch runs as several calls.  Pairs are independent, so
            # only the histories need care: the loop of the whole batch runs until its slowest pair has converged
            # (ICP.py:239-257), and a part that stopped earlier is frozen from then on -- zero steps, weights and costs
            # carried forward (ICP.py:224-232), which is what the padding below writes
            parts = []
            for lo in range(0, N, _MAX_PAIRS_PER_CALL):
                sl = slice(lo, min(N, lo + _MAX_PAIRS_PER_CALL))
                parts.append(_functions.ICPFunction.apply(
                    source_d[sl], target_d[sl], T_init_d[sl], w_d[sl] if w_d is not None else None, params,
                    m_valid[sl] if m_valid is not None else None, None, keep, n_valid[sl] if n_valid is not None else None, status))
            K = max(p_[2].shape[1] for p_ in parts)

            def hist(t, carry):
                if t.shape[1] == K:
                    return t
                tail = t[:, -1:].expand(-1, K - t.shape[1], *t.shape[2:]) if carry else t.new_zeros((t.shape[0], K - t.shape[1]) + t.shape[2:])
                return torch.cat((t, tail), dim=1)
            pc, T = torch.cat([p_[0] for p_ in parts]), torch.cat([p_[1] for p_ in parts])
            deltas = torch.cat([hist(p_[2], False) for p_ in parts])
            weights = torch.cat([hist(p_[3], True) for p_ in parts])
            costs = torch.cat([hist(p_[4], True) for p_ in parts])
            converged, num_iters, match_ratio = (torch.cat([p_[i] for p_ in parts]) for i in (5, 6, 7))
            if keep is not None:
                keep[:] = [torch.cat([torch.cat((c, c[-1:].expand(K - c.shape[0], -1, -1))) for c in keep], dim=1)]
        if keep is not None:
            self.last_correspondences = keep[0]
        self._last_status = status

        if self.verbose:
            print("ICP converged in {} iterations".format(deltas.shape[1]))
            print("Final del_T_ts: {}".format(torch.linalg.norm(deltas[:, -1])))

        if out_dev == dev:
            return {"pc": pc, "T": T, "costs": costs.unsqueeze(-1), "deltas": deltas.unsqueeze(-1),
                    "weights": weights.unsqueeze(-1),
                    "stats": {"converged": converged, "iterations": num_iters, "matched_ratio": match_ratio}}
        back = lambda t: t.to(out_dev)      # noqa: E731
        if not self.lazy_host_results:      # the reference's plain dict of tensors on the input device
            return {"pc": back(pc), "T": back(T), "costs": back(costs).unsqueeze(-1), "deltas": back(deltas).unsqueeze(-1),
                    "weights": back(weights).unsqueeze(-1),
                    "stats": {"converged": back(converged), "iterations": back(num_iters), "matched_ratio": back(match_ratio)}}
        return _Results({
            "pc": _Deferred(lambda: back(pc)),
            "T": back(T),
            "costs": _Deferred(lambda: back(costs).unsqueeze(-1)),
            "deltas": _Deferred(lambda: back(deltas).unsqueeze(-1)),
            "weights": _Deferred(lambda: back(weights).unsqueeze(-1)),
            "stats": {"converged": back(converged), "iterations": back(num_iters), "matched_ratio": back(match_ratio)},
        })

    def check_status(self):
        """Not in the reference: waits for the most recent `icp()` call of this object and returns the number of
        iterations it ran; raises the error a `sync_free` call could not raise at the time (a singular normal-equation
        matrix, for which the reference's torch.linalg.inv raises at ICP.py:201)."""
        if not self._last_status:
            return None
        st = torch.stack(self._last_status).tolist()       # one [iterations run, singular flag] row per kernel-side call
        if any(row[1] != 0.0 for row in st):
            raise RuntimeError("dicp_icp_forward: singular normal-equation matrix (torch.linalg.inv would raise, ICP.py:201)")
        return int(max(row[0] for row in st))

    def _gumbel_loop(self, source, target, T_init, w_init, trim_dist, loss_fn, dim):
        """The Gauss-Newton loop (ICP.py:125-303) for `functionality.gumbel: True` with `differentiable=True`: every
        iteration's neighbours are the Gumbel-softmax soft neighbours of `nn.find_nn` (nn.py:43-70; here the streaming
        online-softmax kernel `dicp_gumbel_nn_*` with counter-based noise, forward and backward), so correspondences are
        soft and gradients also flow through the softmax into the moved points -- which the fused hard-correspondence
        loop does not model.  The per-iteration algebra (residual, trim / robust weights, Jacobian, 6x6 solve, exp-map
        update, bookkeeping) is batched torch on the GPU with autograd, as upstream; this configuration is off in the
        shipped config and outside the benchmarked hot path.  Inputs are the normalised batch tensors of `dICP`."""
        from .loss import loss as _loss
        dev_in = source.device
        dev = self._compute_device(dev_in)
        source, target, T_init, w_init = source.to(dev), target.to(dev), T_init.to(dev), w_init.to(dev)
        N, n = source.shape[0], source.shape[1]
        dt = source.dtype
        pt2pl = self.icp_type == 'pt2pl'
        if dim == 2:                                                           # ICP.py:107-116
            zs = source.new_zeros((N, n, 1))
            source = torch.cat((source[:, :, :2], zs), dim=2)
            zt = target.new_zeros((N, target.shape[1], 1))
            target = torch.cat((target[:, :, :2], zt, target[:, :, 3:5], zt), dim=2) if pt2pl else torch.cat((target[:, :, :2], zt), dim=2)
        steep = self.config['dICP']['parameters']['tanh_steepness']
        trim = _loss("trim", trim_dist, True, steep) if (trim_dist is not None and trim_dist >= 0.0) else None    # ICP.py:152-155
        robust = _loss(loss_fn['name'], loss_fn['metric'], True) if loss_fn is not None else None                 # ICP.py:157-160
        C, r = T_init[:, :3, :3], T_init[:, :3, 3:]
        P = source.transpose(1, 2)
        ncols = 3 if dim == 2 else 6
        reg = 1e-12 * torch.eye(ncols, dtype=dt, device=dev)
        eye3 = torch.eye(3, dtype=dt, device=dev)
        converged = torch.zeros(N, dtype=torch.bool, device=dev)
        num_iters = torch.zeros(N, dtype=dt, device=dev)
        match_ratio = torch.zeros(N, dtype=dt, device=dev)
        deltas, weights, costs = [], [], []
        w = w_init
        k_run = 0
        for k in range(self.max_iterations):
            k_run = k + 1
            A_pts = C @ P
            X = (A_pts + r).transpose(1, 2)                                    # ICP.py:137
            nbr = self.nn.find_nn(X, target)                                   # soft neighbours, (N,n,S)   ICP.py:140
            e3 = X - nbr[:, :, :3]
            a = A_pts.transpose(1, 2)
            if pt2pl:
                nrm = nbr[:, :, 3:]
                err = (e3 * nrm).sum(dim=2, keepdim=True)                      # ICP.py:146
                J = torch.cat((torch.cross(nrm, a, dim=2), -nrm), dim=2)       # ICP.py:172-176
            else:
                err = e3
                z = torch.zeros_like(a[..., 0])
                hat = torch.stack((z, -a[..., 2], a[..., 1], a[..., 2], z, -a[..., 0], -a[..., 1], a[..., 0], z), dim=-1).reshape(N, n, 3, 3)
                J = torch.cat((hat.reshape(N, -1, 3), -eye3.repeat(N, n, 1)), dim=2)      # ICP.py:177-180
            t_w = trim.get_weight(e3) if trim is not None else torch.ones((N, n), dtype=dt, device=dev)
            l_w = robust.get_weight(err) if robust is not None else torch.ones((N, n), dtype=dt, device=dev)
            if not pt2pl:                                                      # ICP.py:162-166
                t_w, l_w = t_w.repeat_interleave(3, dim=1), l_w.repeat_interleave(3, dim=1)
                err = err.reshape(N, -1, 1)
            w = w_init * t_w * l_w                                             # ICP.py:169
            if dim == 2:
                J = J[:, :, 2:5]                                               # ICP.py:186-189
            g = torch.sqrt(w + 1.0e-10) - 1.0e-5                               # ICP.py:194
            err_w, J_w = g.unsqueeze(-1) * err, g.unsqueeze(-1) * J
            step = -torch.linalg.inv(J_w.transpose(1, 2) @ J_w + reg) @ J_w.transpose(1, 2) @ err_w     # ICP.py:200-201
            if dim == 2:                                                       # ICP.py:204-207
                step = torch.cat((step.new_zeros((N, 2, 1)), step, step.new_zeros((N, 1, 1))), dim=1)
            phi = step[:, 0:3, 0]
            zz = torch.zeros_like(phi[:, 0])
            phat = torch.stack((zz, -phi[:, 2], phi[:, 1], phi[:, 2], zz, -phi[:, 0], -phi[:, 1], phi[:, 0], zz), dim=-1).reshape(N, 3, 3)
            C = torch.matrix_exp(phat).transpose(1, 2) @ C                     # ICP.py:210-214
            r = r - step[:, 3:6]                                               # ICP.py:216
            deltas.append(step.detach())
            w_keep = w.detach()
            if weights:                                                        # ICP.py:224-226
                w_keep = torch.where((w_keep.sum(dim=1) == 0.0).unsqueeze(-1), weights[-1].squeeze(-1), w_keep)
            weights.append(w_keep.unsqueeze(-1))
            cost = (err_w.transpose(1, 2) @ err_w).detach()                    # ICP.py:229
            if costs:
                cost = torch.where(cost == 0.0, costs[-1], cost)
            costs.append(cost)
            below = torch.linalg.norm(step, dim=1).detach().squeeze(-1) < self.tolerance    # ICP.py:237
            converged = converged | below
            if bool(below.any()) and not self.const_iter:                      # ICP.py:240-260
                num_iters = torch.where(below & (num_iters == 0), torch.full_like(num_iters, k + 1), num_iters)
                matched = (w > self.match_ratio_thresh).sum(dim=1)
                start = (w_init > self.match_ratio_thresh).sum(dim=1)
                start = torch.where(start == 0, torch.ones_like(start), start)
                match_ratio = torch.where(below & (match_ratio == 0), match_ratio + matched / start, match_ratio)
                w_init = w_init * (~below).to(dt).unsqueeze(-1)
                if bool(below.all()):
                    break
        num_iters = torch.where(num_iters == 0, torch.full_like(num_iters, k_run), num_iters)   # ICP.py:267-271
        matched = (w > self.match_ratio_thresh).sum(dim=1)
        start = (w_init > self.match_ratio_thresh).sum(dim=1)
        start = torch.where(start == 0, torch.ones_like(start), start)
        match_ratio = torch.where(match_ratio == 0, (matched / start).to(dt), match_ratio)
        pc = (C @ P + r).transpose(1, 2)                                       # ICP.py:274
        T = torch.cat((torch.cat((C, r), dim=2), torch.tensor([[[0.0, 0.0, 0.0, 1.0]]], dtype=dt, device=dev).repeat(N, 1, 1)), dim=1)
        back = (lambda t: t) if dev_in == dev else (lambda t: t.to(dev_in))
        return {"pc": back(pc), "T": back(T), "costs": back(torch.stack(costs, dim=1).squeeze(-1)),
                "deltas": back(torch.stack(deltas, dim=1)), "weights": back(torch.stack(weights, dim=1)),
                "stats": {"converged": back(converged), "iterations": back(num_iters), "matched_ratio": back(match_ratio)}}

    def _compute_device(self, input_device):
        _functions.require_cuda()
        if input_device.type == 'cuda':
            return input_device
        return torch.device(self.device) if self.device is not None else torch.device('cuda', torch.cuda.current_device())

    # ---------------------------------------------------------------------------------
    def batch_size_handling(self, source, target, T_init=None, weight=None):
        """Normalise every accepted input form to dense batch tensors (ICP.py:305-511):
        source (n,3|6) | (N,n,3|6) | list of (n_i,3|6)  -> (N,n,3), zero-padded, zero weight on pads
        target (m,3|6) | (N,m,3|6) | list of (m_i,3|6)  -> (N,m,S), padded with max(source)*target_pad_val
        T_init (4,4) | (N,4,4) | list                   -> (N,4,4) or None
        weight None | (n,) | (N,n) | list (entries may be None) -> w (N,n), or (N,3n) for pt2pt.
        Empty / missing clouds become a single phony point with zero weight."""
        rows_per_point = 3 if self.icp_type == 'pt2pt' else 1
        if weight is not None:
            if isinstance(source, list):
                assert len(source) == len(weight), "weight must be list of same length as source"
            else:
                assert source.shape[0] == weight.shape[0], "weight must have same number of rows as source"

        if _is_empty(source) or _is_empty(target):
            # whole source or target missing: phony CPU/fp32 problem that returns T = I (ICP.py:329-346)
            w = torch.zeros((1, rows_per_point), dtype=torch.float32)
            return (torch.zeros((1, 1, 3), dtype=torch.float32), torch.zeros((1, 1, 6), dtype=torch.float32),
                    torch.eye(4, dtype=torch.float32).unsqueeze(0), w)

        # dtype / device / column count come from the first non-empty target (ICP.py:347-358)
        dev, dt, tcols = "cpu", torch.float32, None
        for t_i in target:
            if t_i is not None and len(t_i) > 0:
                dev, dt = t_i.device, t_i.dtype
                tcols = t_i.shape[0] if (not isinstance(target, list) and target.dim() == 2) else t_i.shape[1]
                break

        # ---- source and prior weights ------------------------------------------------
        if isinstance(source, list):
            clouds, priors = [], []
            for i, s_i in enumerate(source):
                empty = len(s_i) == 0
                if empty:
                    s_i = torch.zeros((1, 3), dtype=dt, device=dev)
                if i > 0 and (s_i.dim() != 2 or s_i.shape[1] not in (3, 6)):
                    raise ValueError("source list must contain (n x 3/6) tensors")
                prior = torch.ones(s_i.shape[0], dtype=dt, device=dev)
                if weight is not None and weight[i] is not None and not (empty and i == 0):
                    assert len(weight[i]) == s_i.shape[0], "weight must have same number of rows as source"
                    prior = weight[i] * prior
                if empty:
                    prior = 0.0 * prior
                clouds.append(s_i[:, :3])
                priors.append(prior)
            source_batch = self._stack_ragged(clouds, 0.0)
            w = self._stack_ragged([p_.unsqueeze(-1) for p_ in priors], 0.0).squeeze(-1)
        elif source.dim() == 2 and source.shape[1] in (3, 6):
            source_batch = source[:, :3].unsqueeze(0)
            w = torch.ones((1, source_batch.shape[1]), dtype=dt, device=dev) if weight is None else weight.unsqueeze(0)
        elif source.dim() == 3 and source.shape[2] in (3, 6):
            source_batch = source[:, :, :3]
            w = torch.ones(source_batch.shape[:2], dtype=dt, device=dev) if weight is None else weight
        else:
            raise ValueError("source must be (n x 3/6) or (N x n x 3/6) or list len(N) (n_N x 3/6)")

        if self.source_zeroes_are_pad:
            w = w * (torch.linalg.norm(source_batch, dim=2) != 0.0).to(dt)

        # ---- target ------------------------------------------------------------------
        if isinstance(target, list):
            pad_value = torch.max(source_batch) * self.target_pad_val       # ICP.py:460
            clouds, dead = [], []
            for i, t_i in enumerate(target):
                if len(t_i) == 0:
                    t_i = torch.zeros((1, tcols), dtype=dt, device=dev)
                    dead.append(i)
                if i > 0 and (t_i.dim() != 2 or t_i.shape[1] != tcols):
                    raise ValueError("target list must contain (m x 3/6) tensors. All tensors must have same number of columns")
                clouds.append(t_i)
            target_batch = self._stack_ragged(clouds, pad_value)
            if dead:
                alive = torch.ones((w.shape[0], 1), dtype=w.dtype, device=w.device)
                alive[dead] = 0.0
                w = w * alive                                               # ICP.py:456, 467
        elif target.dim() == 2 and target.shape[1] in (3, 6):
            target_batch = target.unsqueeze(0)
        elif target.dim() == 3 and target.shape[2] in (3, 6):
            target_batch = target
        else:
            raise ValueError("target must be (m x 3/6) or (N x m x 3/6) or list len(N) (m_N x 3/6)")

        # ---- initial transform ---------------------------------------------------------
        if T_init is None:
            T_batch = None
        elif isinstance(T_init, list):
            T_batch = torch.stack(T_init, dim=0)
        elif T_init.shape == (4, 4):
            T_batch = T_init.unsqueeze(0)
        elif T_init.dim() == 3 and T_init.shape[1:] == (4, 4):
            T_batch = T_init
        else:
            raise ValueError("T_init must be (4 x 4) or (N x 4 x 4) or list len(N) (4 x 4)")

        if rows_per_point == 3:
            w = w.repeat_interleave(3, dim=1)                               # ICP.py:508-509
        return source_batch, target_batch, T_batch, w

    @staticmethod
    def _stack_ragged(clouds, value):
        """list of (k_i,c) -> (N, max k_i, c): ONE allocation filled with `value` (a float or a 0-d tensor) and one slice copy
        per cloud, instead of the reference's incremental vstack / hstack of padded copies (ICP.py:361-426, 451-482)."""
        rows = max(c.shape[0] for c in clouds)
        first = clouds[0]
        out = (value * torch.ones((len(clouds), rows, first.shape[1]), dtype=first.dtype, device=first.device))
        for i, c in enumerate(clouds):
            out[i, :c.shape[0]] = c
        return out

    # ---------------------------------------------------------------------------------
    def pt2pt_dICP_SVD(self, source, target, T_init, trim_dist=None, huber_delta=None, dim=3):
        """Kabsch / SVD point-to-point ICP (ICP.py:533-591; "not yet integrated" upstream, so `trim_dist`,
        `huber_delta`, `dim` are accepted and ignored as they are there).  Optional YAML key
        `dICP.functionality.svd: True` makes `icp()` of a 'pt2pt' object take this path.

        source (n,>=3) / target (m,>=3) / T_init (4,4) as in the reference, or batched (N,n,.), (N,m,.),
        (N,4,4).  Returns (moved source, T).  The whole loop runs on the device behind one autograd node
        (`_functions.KabschICPFunction`): per iteration the nearest-neighbour search, the fused centroid +
        cross-covariance reduction, one thread per pair for the 3x3 SVD (one-sided Jacobi in double),
        determinant fix and composition, and the convergence test; the reverse pass is hand-written
        (adjoint of the SVD step included).  A batch stops when every pair is below the tolerance.

        Reference quirks kept: `T_init` is only composed into the returned transform and is not applied
        to the source before the first search (ICP.py:545-547); convergence compares the moved points
        with the neighbours found before the move (ICP.py:585).  One defect is not kept by default:
        ICP.py:566-570 multiplies by V where the Kabsch rotation needs V^T (torch.svd returns V), which is
        only right when V is symmetric, as on the planar test fixture; here R = U diag(1,1,det(U)det(V)) V^T.
        `self.svd_literal = True` switches to the reference's literal product U D V; its value depends on the
        sign convention of the SVD routine (LAPACK's is not specified), so parity with the reference is only
        defined where V = V^T."""
        single = source.dim() == 2
        ps = (source.unsqueeze(0) if single else source)[:, :, 0:3]
        tgt = (target.unsqueeze(0) if target.dim() == 2 else target)[:, :, 0:3]
        T = T_init.unsqueeze(0) if T_init.dim() == 2 else T_init
        out_dev = ps.device
        dev = self._compute_device(out_dev)
        ps, tgt, T = ps.to(dev).contiguous(), tgt.to(dev).contiguous(), T.to(dev).to(ps.dtype)
        pc, T_out = _functions.KabschICPFunction.apply(ps, tgt, T, self.max_iterations, self.tolerance,
                                                       getattr(self, "svd_literal", False), self.check_every)
        if self.verbose:
            print("ICP converged")
        pc, T_out = pc.to(out_dev), T_out.to(out_dev)
        return (pc[0], T_out[0]) if single else (pc, T_out)
