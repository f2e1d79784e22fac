"""Test-only binding of tests/host_emul/libdicp_emul.so: the CUDA kernels' arithmetic
(dicp_b200/csrc/dicp_math.cuh) compiled for the host and driven sequentially.  Lets the
GPU-less suite check the math the kernels run.  Never imported by the product."""
import ctypes
import os
import subprocess

import torch

from dicp_b200._lib import DicpParams, STATE_STRIDE

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "emul.cpp")
LIB = os.path.join(HERE, "host_emul", "libdicp_emul.so")
HDR = os.path.join(HERE, "..", "dicp_b200", "csrc", "dicp_math.cuh")
ABI = os.path.join(HERE, "..", "include", "dicp_b200.h")
_lib = None


def load():
    global _lib
    if _lib is None:
        stale = (not os.path.exists(LIB)) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR), os.path.getmtime(ABI))
        if stale:
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++", SRC, "-o", LIB])
        _lib = ctypes.CDLL(LIB)
    return _lib


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def forward(source, target, T_init, w_init, params):
    """CPU tensors in, dict out (same fields the CUDA Function produces)."""
    lib = load()
    sfx = "f32" if source.dtype == torch.float32 else "f64"
    dt = source.dtype
    source, target, T_init = source.contiguous(), target.contiguous(), T_init.contiguous()
    w_init = w_init.contiguous() if w_init is not None else None
    N, n, _ = source.shape
    m, S = target.shape[1], target.shape[2]
    K = params.max_iterations
    R = 1 if params.icp_type == 0 else 3
    out = dict(pc=torch.empty((N, n, 3), dtype=dt), T=torch.empty((N, 4, 4), dtype=dt), deltas=torch.empty((N, K, 6), dtype=dt),
               weights=torch.empty((N, K, n * R), dtype=dt), costs=torch.empty((N, K), dtype=dt),
               converged=torch.empty((N,), dtype=torch.uint8), iterations=torch.empty((N,), dtype=dt),
               matched_ratio=torch.empty((N,), dtype=dt), saved_idx=torch.empty((K, N, n), dtype=torch.int32),
               saved_state=torch.empty((K + 1, N, STATE_STRIDE), dtype=torch.float64))
    iters = ctypes.c_int32(0)
    fn = getattr(lib, "emul_icp_forward_" + sfx)
    fn.restype = ctypes.c_int
    rc = fn(_p(source), _p(target), _p(T_init), _p(w_init), None, None, N, n, m, S, ctypes.byref(params), _p(out["pc"]), _p(out["T"]),
            _p(out["deltas"]), _p(out["weights"]), _p(out["costs"]), _p(out["converged"]), _p(out["iterations"]),
            _p(out["matched_ratio"]), _p(out["saved_idx"]), _p(out["saved_state"]), ctypes.byref(iters), None,
            ctypes.c_size_t(0), None)
    assert rc == 0, rc
    k = iters.value
    out["k_run"] = k
    for key in ("deltas", "weights", "costs"):
        out[key] = out[key][:, :k]
    out["converged"] = out["converged"].bool()
    return out


def backward(source, target, w_init, params, fwd, grad_T, grad_pc):
    lib = load()
    sfx = "f32" if source.dtype == torch.float32 else "f64"
    dt = source.dtype
    source, target = source.contiguous(), target.contiguous()
    N, n, _ = source.shape
    m, S = target.shape[1], target.shape[2]
    g = dict(source=torch.empty_like(source), target=torch.empty_like(target),
             w=torch.empty((N, n), dtype=dt) if w_init is not None else None, T_init=torch.empty((N, 4, 4), dtype=dt))
    fn = getattr(lib, "emul_icp_backward_" + sfx)
    fn.restype = ctypes.c_int
    rc = fn(_p(source), _p(target), _p(w_init.contiguous()) if w_init is not None else None, None, N, n, m, S, ctypes.byref(params),
            fwd["k_run"], _p(fwd["saved_idx"]), _p(fwd["saved_state"]), _p(grad_T.contiguous()) if grad_T is not None else None,
            _p(grad_pc.contiguous()) if grad_pc is not None else None, _p(g["source"]), _p(g["target"]), _p(g["w"]), _p(g["T_init"]),
            None, ctypes.c_size_t(0), None)
    assert rc == 0, rc
    return g
