"""torch.autograd.Function wrappers around the C ABI (include/dicp_b200.h).

PyTorch is plumbing here: it owns device memory and streams and carries the autograd
graph edge; every number is produced by libdicp_b200.so."""
import ctypes

import torch

from . import _lib
from ._lib import DicpParams, check, load_library, ptr, require_cuda, stream_ptr, suffix


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def nn_argmin(x, y, return_d2=False, search="brute", m_valid=None):
    """x (N,n,3), y (N,m,3|6) CUDA tensors -> idx (N,n) int32 [, d2 (N,n)] (nn.py:32-35).
    search: 'brute' (tiled scan) or 'grid' (uniform grid, same result).
    m_valid (N,) int32 CUDA tensor or None: real row count of every target cloud of a padded batch."""
    require_cuda()
    lib = load_library()
    assert x.is_cuda and y.is_cuda and x.dtype == y.dtype
    x = x.detach().contiguous()
    y = y.detach().contiguous()
    N, n, _ = x.shape
    m, S = y.shape[1], y.shape[2]
    sfx = suffix(x.dtype)
    idx = torch.empty((N, n), dtype=torch.int32, device=x.device)
    d2 = torch.empty((N, n), dtype=x.dtype, device=x.device) if return_d2 else None
    if search == "grid":
        name = "dicp_nn_argmin_grid_" + sfx
        ws = _workspace(lib.dicp_nn_grid_workspace_bytes(N, n, m, x.element_size()), x.device)
    else:
        name = "dicp_nn_argmin_" + sfx
        ws = _workspace(lib.dicp_nn_workspace_bytes(N, n, m, x.element_size()), x.device)
    with torch.cuda.device(x.device):
        rc = getattr(lib, name)(ptr(x), ptr(y), N, n, m, S, ptr(m_valid), ptr(idx), ptr(d2), ptr(ws), ws.numel(), stream_ptr(x.device))
    check(rc, name)
    return (idx, d2) if return_d2 else idx


class GatherRows(torch.autograd.Function):
    """neighbours = y[b, idx[b,i], :]; backward is the scatter-add into y (nn.py:37-38)."""

    @staticmethod
    def forward(ctx, y, idx):
        lib = load_library()
        y = y.contiguous()
        N, m, S = y.shape
        n = idx.shape[1]
        out = torch.empty((N, n, S), dtype=y.dtype, device=y.device)
        sfx = suffix(y.dtype)
        with torch.cuda.device(y.device):
            rc = getattr(lib, "dicp_gather_rows_" + sfx)(ptr(y), ptr(idx), N, n, m, S, ptr(out), stream_ptr(y.device))
        check(rc, "dicp_gather_rows_" + sfx)
        ctx.save_for_backward(idx)
        ctx.shape = (N, m, S)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = load_library()
        (idx,) = ctx.saved_tensors
        N, m, S = ctx.shape
        n = idx.shape[1]
        g = grad_out.contiguous()
        gy = torch.zeros((N, m, S), dtype=g.dtype, device=g.device)
        sfx = suffix(g.dtype)
        with torch.cuda.device(g.device):
            rc = getattr(lib, "dicp_gather_rows_bwd_" + sfx)(ptr(g), ptr(idx), N, n, m, S, ptr(gy), stream_ptr(g.device))
        check(rc, "dicp_gather_rows_bwd_" + sfx)
        return gy, None


class GumbelNN(torch.autograd.Function):
    """Streaming Gumbel-softmax soft nearest neighbour (nn.py:43-70): x (N,n,3), y (N,m,S) -> (N,n,S).
    The noise is a pure function of (seed, pair, query, target), so the backward regenerates it."""

    @staticmethod
    def forward(ctx, x, y, seed, eps, tau):
        lib = load_library()
        x, y = x.contiguous(), y.contiguous()
        N, n, _ = x.shape
        m, S = y.shape[1], y.shape[2]
        sfx = suffix(x.dtype)
        out = torch.empty((N, n, S), dtype=x.dtype, device=x.device)
        lse = torch.empty((N, n), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            rc = getattr(lib, "dicp_gumbel_nn_" + sfx)(ptr(x), ptr(y), N, n, m, S, int(seed), float(eps), float(tau), ptr(out), ptr(lse),
                                                       stream_ptr(x.device))
        check(rc, "dicp_gumbel_nn_" + sfx)
        ctx.save_for_backward(x, y, out, lse)
        ctx.cfg = (int(seed), float(eps), float(tau))
        return out

    @staticmethod
    def backward(ctx, g):
        lib = load_library()
        x, y, out, lse = ctx.saved_tensors
        seed, eps, tau = ctx.cfg
        N, n, _ = x.shape
        m, S = y.shape[1], y.shape[2]
        sfx = suffix(x.dtype)
        gx, gy = torch.empty_like(x), torch.empty_like(y)
        scratch = torch.empty((N, n), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            rc = getattr(lib, "dicp_gumbel_nn_bwd_" + sfx)(ptr(x), ptr(y), ptr(out), ptr(lse), ptr(g.contiguous()), N, n, m, S, seed, eps, tau,
                                                           ptr(gx), ptr(gy), ptr(scratch), stream_ptr(x.device))
        check(rc, "dicp_gumbel_nn_bwd_" + sfx)
        return gx, gy, None, None, None


def gumbel_noise(N, n, m, seed, eps, dtype, device):
    """The noise tensor G (N,n,m) GumbelNN uses for `seed` (tests)."""
    lib = load_library()
    G = torch.empty((N, n, m), dtype=dtype, device=device)
    with torch.cuda.device(device):
        rc = getattr(lib, "dicp_gumbel_noise_" + suffix(dtype))(N, n, m, int(seed), float(eps), ptr(G), stream_ptr(device))
    check(rc, "dicp_gumbel_noise")
    return G


class KabschSums(torch.autograd.Function):
    """sums[b] = (sum p, sum q, sum q p^T) over the correspondences q_i = y[b, idx[b,i], :3]
    (ICP.py:557-563 without the O(1) 3x3 part).  x (N,n,3), y (N,m,S), idx (N,n) int32 -> (N,16)."""

    @staticmethod
    def forward(ctx, x, y, idx):
        lib = load_library()
        x, y = x.contiguous(), y.contiguous()
        N, n, _ = x.shape
        m, S = y.shape[1], y.shape[2]
        sfx = suffix(x.dtype)
        sums = torch.empty((N, 16), dtype=x.dtype, device=x.device)
        ws = _workspace(lib.dicp_kabsch_workspace_bytes(N, n), x.device)
        with torch.cuda.device(x.device):
            rc = getattr(lib, "dicp_kabsch_sums_" + sfx)(ptr(x), ptr(y), ptr(idx), N, n, m, S, ptr(sums), ptr(ws), ws.numel(),
                                                         stream_ptr(x.device))
        check(rc, "dicp_kabsch_sums_" + sfx)
        ctx.save_for_backward(x, y, idx)
        return sums

    @staticmethod
    def backward(ctx, g):
        lib = load_library()
        x, y, idx = ctx.saved_tensors
        N, n, _ = x.shape
        m, S = y.shape[1], y.shape[2]
        sfx = suffix(x.dtype)
        gx, gy = torch.empty_like(x), torch.empty_like(y)
        with torch.cuda.device(x.device):
            rc = getattr(lib, "dicp_kabsch_sums_bwd_" + sfx)(ptr(x), ptr(y), ptr(idx), N, n, m, S, ptr(g.contiguous()), ptr(gx), ptr(gy),
                                                             stream_ptr(x.device))
        check(rc, "dicp_kabsch_sums_bwd_" + sfx)
        return gx, gy, None


class KabschICPFunction(torch.autograd.Function):
    """The whole Kabsch / SVD point-to-point loop (ICP.py:533-591) as one autograd node: nearest-neighbour search,
    fused centroid / cross-covariance sums, 3x3 SVD + determinant fix + composition and the convergence test run as
    kernels (`dicp_kabsch_icp_forward_*`), and so does the reverse pass (`dicp_kabsch_icp_backward_*`).
    forward(source (N,n,3), target (N,m,S), T_init (N,4,4), max_iterations, tolerance, literal, check_every)
      -> pc (N,n,3), T (N,4,4), iterations run (python int, non-differentiable)."""

    @staticmethod
    def forward(ctx, source, target, T_init, max_iterations, tolerance, literal, check_every):
        require_cuda()
        lib = load_library()
        dev, dt = source.device, source.dtype
        source, target, T_init = source.contiguous(), target.contiguous(), T_init.contiguous()
        N, n, _ = source.shape
        m, S = target.shape[1], target.shape[2]
        K = int(max_iterations)
        sfx = suffix(dt)
        pc = torch.empty((N, n, 3), dtype=dt, device=dev)
        T = torch.empty((N, 4, 4), dtype=dt, device=dev)
        saved_idx = torch.empty((K, N, n), dtype=torch.int32, device=dev)
        saved_state = torch.empty((K + 1, N, _lib.STATE_STRIDE), dtype=torch.float64, device=dev)
        ws = _workspace(lib.dicp_kabsch_icp_workspace_bytes(N, n, m, source.element_size()), dev)
        iters = ctypes.c_int32(0)
        with torch.cuda.device(dev):
            rc = getattr(lib, "dicp_kabsch_icp_forward_" + sfx)(
                ptr(source), ptr(target), ptr(T_init), N, n, m, S, K, float(tolerance), int(bool(literal)), int(check_every),
                ptr(pc), ptr(T), ptr(saved_idx), ptr(saved_state), ctypes.byref(iters), ptr(ws), ws.numel(), stream_ptr(dev))
        check(rc, "dicp_kabsch_icp_forward_" + sfx)
        ctx.k_run = int(iters.value)
        ctx.literal = int(bool(literal))
        ctx.dims = (N, n, m, S)
        ctx.ws = ws
        ctx.save_for_backward(source, target, T_init, saved_idx, saved_state)
        ctx.iterations = ctx.k_run
        return pc, T

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_pc, g_T):
        lib = load_library()
        source, target, T_init, saved_idx, saved_state = ctx.saved_tensors
        N, n, m, S = ctx.dims
        dev, dt = source.device, source.dtype
        sfx = suffix(dt)
        g_pc = g_pc.contiguous() if g_pc is not None else None
        g_T = g_T.contiguous() if g_T is not None else None
        g_src, g_tgt = torch.empty_like(source), torch.empty_like(target)
        g_T0 = torch.empty((N, 4, 4), dtype=dt, device=dev)
        with torch.cuda.device(dev):
            rc = getattr(lib, "dicp_kabsch_icp_backward_" + sfx)(
                ptr(source), ptr(target), ptr(T_init), N, n, m, S, ctx.literal, ctx.k_run, ptr(saved_idx), ptr(saved_state),
                ptr(g_T), ptr(g_pc), ptr(g_src), ptr(g_tgt), ptr(g_T0), ptr(ctx.ws), ctx.ws.numel(), stream_ptr(dev))
        check(rc, "dicp_kabsch_icp_backward_" + sfx)
        return g_src, g_tgt, g_T0, None, None, None, None


def make_params(icp_type, dim, differentiable, loss_name, loss_metric, trim_dist, tanh_steepness, tolerance,
                max_iterations, const_iter, match_ratio_thresh, check_every=1, nn_search="brute", sync_free=False):
    if loss_name not in _lib.LOSS_IDS:
        raise ValueError("Invalid loss name: {}".format(loss_name))          # loss.py:19
    p = DicpParams()
    p.icp_type = _lib.DICP_PT2PL if icp_type == "pt2pl" else _lib.DICP_PT2PT
    p.dim = int(dim)
    p.differentiable = int(bool(differentiable))
    p.loss_id = _lib.LOSS_IDS[loss_name]
    use_trim = trim_dist is not None and trim_dist >= 0.0                    # ICP.py:153
    p.use_trim = int(use_trim)
    p.max_iterations = int(max_iterations)
    p.const_iter = int(bool(const_iter))
    p.check_every = int(check_every)
    if nn_search not in _lib.NN_MODES:
        raise ValueError("nn_search must be 'brute' or 'grid', got {!r}".format(nn_search))
    p.nn_mode = _lib.NN_MODES[nn_search]
    p.flags = _lib.DICP_FLAG_NO_SYNC if (sync_free and const_iter) else 0
    p.loss_metric = float(loss_metric) if loss_name is not None else 1.0
    p.trim_dist = float(trim_dist) if use_trim else 0.0
    p.tanh_steepness = float(tanh_steepness)
    p.tolerance = float(tolerance)
    p.match_ratio_thresh = float(match_ratio_thresh)
    return p


class ICPFunction(torch.autograd.Function):
    """The whole solver loop (ICP.py:125-303) as one autograd node.

    forward(source (N,n,3), target (N,m,S), T_init (N,4,4), w_init (N,n) | None, params)
      -> pc, T (differentiable) and deltas, weights, costs, converged, num_iters, match_ratio
         (detached / non-differentiable, ICP.py:220-229, 288-292).
    m_valid (N,) int32 or None: real row count of every target cloud of a ragged batch.
    n_valid (N,) int32 or None: real row count of every source cloud of a ragged batch (rows beyond it: padding).
    nn_hook: None, or a callable `nn_hook(k, state_k, idx_k)` that replaces the loop's own
      correspondence search (include/dicp_b200.h: dicp_icp_forward_hooked_*): `state_k` is the
      (N, 64) float64 record tensor of iteration k (pass it to `dicp_nn_scan_*`), `idx_k` the
      (N, n) int32 tensor to fill, both views of this call's saved buffers, on the current stream."""

    @staticmethod
    def forward(ctx, source, target, T_init, w_init, params, m_valid=None, nn_hook=None, keep_idx=None, n_valid=None,
                status_out=None):
        require_cuda()
        lib = load_library()
        dev, dt = source.device, source.dtype
        assert source.is_cuda and target.device == dev and T_init.device == dev
        source, target, T_init = source.contiguous(), target.contiguous(), T_init.contiguous()
        if w_init is not None:
            w_init = w_init.contiguous()
        N, n, _ = source.shape
        m, S = target.shape[1], target.shape[2]
        K = params.max_iterations
        R = 1 if params.icp_type == _lib.DICP_PT2PL else 3
        sfx = suffix(dt)
        pc = torch.empty((N, n, 3), dtype=dt, device=dev)
        T = torch.empty((N, 4, 4), dtype=dt, device=dev)
        deltas = torch.empty((N, K, 6), dtype=dt, device=dev)
        weights = torch.empty((N, K, n * R), dtype=dt, device=dev)
        costs = torch.empty((N, K), dtype=dt, device=dev)
        converged = torch.empty((N,), dtype=torch.uint8, device=dev)
        num_iters = torch.empty((N,), dtype=dt, device=dev)
        match_ratio = torch.empty((N,), dtype=dt, device=dev)
        saved_idx = torch.empty((K, N, n), dtype=torch.int32, device=dev)
        saved_state = torch.empty((K + 1, N, _lib.STATE_STRIDE), dtype=torch.float64, device=dev)
        ws = _workspace(lib.dicp_icp_workspace_bytes(N, n, m, source.element_size()), dev)
        iters = ctypes.c_int32(0)
        args = (ptr(source), ptr(target), ptr(T_init), ptr(w_init), ptr(m_valid), ptr(n_valid), N, n, m, S, ctypes.byref(params),
                ptr(pc), ptr(T), ptr(deltas), ptr(weights), ptr(costs), ptr(converged), ptr(num_iters), ptr(match_ratio),
                ptr(saved_idx), ptr(saved_state), ctypes.byref(iters), ptr(ws), ws.numel(), stream_ptr(dev))
        with torch.cuda.device(dev):
            if nn_hook is None:
                rc = getattr(lib, "dicp_icp_forward_" + sfx)(*args)
            else:
                failure = []

                def _c_hook(_user, k, _state_ptr, _idx_ptr):      # the pointers are saved_state[k] / saved_idx[k]
                    try:
                        nn_hook(int(k), saved_state[k], saved_idx[k])
                        return 0
                    except BaseException as exc:                  # noqa: BLE001 - must not unwind through C
                        failure.append(exc)
                        return _lib.DICP_ERR_HOOK
                c_hook = _lib.NN_HOOK(_c_hook)
                rc = getattr(lib, "dicp_icp_forward_hooked_" + sfx)(*args, c_hook, None)
                if failure:
                    raise failure[0]
        check(rc, "dicp_icp_forward_" + sfx)
        k_run = int(iters.value)
        ctx.params = params
        ctx.k_run = k_run
        ctx.dims = (N, n, m, S)
        ctx.has_w = w_init is not None
        ctx.n_valid = n_valid
        del ws                                    # the reverse pass allocates its own, much smaller workspace
        if k_run < K:
            # early stop: the history buffers were sized for max_iterations; keep only the iterations that ran (with the
            # default max_iterations = 100 and a handful of iterations this is the difference between 1.6 GB and 0.1 GB
            # each for `weights` and the saved correspondences of a 256 x 16k batch)
            deltas, weights, costs = deltas[:, :k_run].clone(), weights[:, :k_run].clone(), costs[:, :k_run].clone()
            saved_idx, saved_state = saved_idx[:k_run].clone(), saved_state[:k_run + 1].clone()
        ctx.save_for_backward(source, target, w_init if w_init is not None else source.new_empty(0),
                              saved_idx, saved_state)
        if keep_idx is not None:                 # the caller asked for the correspondences (ICP.keep_correspondences)
            keep_idx.append(saved_idx[:k_run])
        if status_out is not None:               # [iterations run, singular flag] of this call, still on the device
            status_out.append(saved_state[0, 0, _lib.STATE_ITERS_RUN:_lib.STATE_SINGULAR + 1])
        outs = (pc, T, deltas[:, :k_run], weights[:, :k_run], costs[:, :k_run], converged.bool(), num_iters, match_ratio)
        ctx.mark_non_differentiable(*outs[2:])
        return outs

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_pc, g_T, *_unused):
        lib = load_library()
        source, target, w_init, saved_idx, saved_state = ctx.saved_tensors
        N, n, m, S = ctx.dims
        dev, dt = source.device, source.dtype
        sfx = suffix(dt)
        g_pc = g_pc.contiguous() if g_pc is not None else None
        g_T = g_T.contiguous() if g_T is not None else None
        g_src = torch.empty_like(source)
        g_tgt = torch.empty_like(target)
        g_w = torch.empty((N, n), dtype=dt, device=dev) if (ctx.has_w and ctx.needs_input_grad[3]) else None
        g_T0 = torch.empty((N, 4, 4), dtype=dt, device=dev)
        ws = _workspace(lib.dicp_icp_backward_workspace_bytes(N, n, m, source.element_size()), dev)
        with torch.cuda.device(dev):
            rc = getattr(lib, "dicp_icp_backward_" + sfx)(
                ptr(source), ptr(target), ptr(w_init) if ctx.has_w else None, ptr(ctx.n_valid), N, n, m, S, ctypes.byref(ctx.params), ctx.k_run,
                ptr(saved_idx), ptr(saved_state), ptr(g_T), ptr(g_pc), ptr(g_src), ptr(g_tgt), ptr(g_w), ptr(g_T0),
                ptr(ws), ws.numel(), stream_ptr(dev))
        check(rc, "dicp_icp_backward_" + sfx)
        return g_src, g_tgt, g_T0, g_w, None, None, None, None, None, None
