"""ctypes binding of libdicp_b200.so (the C ABI in include/dicp_b200.h).

Fails loudly: if the library has not been built, or no CUDA device is present, the product
path raises — there is no CPU / eager fallback anywhere in this package."""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DICP_B200_LIB", os.path.join(_HERE, "lib", "libdicp_b200.so"))

DICP_PT2PL, DICP_PT2PT = 0, 1
LOSS_IDS = {None: 0, "huber": 1, "cauchy": 2}
NN_MODES = {"brute": 0, "grid": 1}
STATE_STRIDE = 64
STATE_ITERS_RUN, STATE_SINGULAR = 62, 63      # DICP_STATE_ITERS_RUN / DICP_STATE_SINGULAR of record [0][0]
DICP_FLAG_NO_SYNC = 1
DICP_ERR_HOOK = -4          # a Python correspondence hook raised (dicp_icp_forward_hooked_*)


class DicpParams(ctypes.Structure):
    """struct dicp_params (include/dicp_b200.h)."""
    _fields_ = [
        ("icp_type", ctypes.c_int32), ("dim", ctypes.c_int32), ("differentiable", ctypes.c_int32),
        ("loss_id", ctypes.c_int32), ("use_trim", ctypes.c_int32), ("max_iterations", ctypes.c_int32),
        ("const_iter", ctypes.c_int32), ("check_every", ctypes.c_int32), ("nn_mode", ctypes.c_int32), ("flags", ctypes.c_int32),
        ("loss_metric", ctypes.c_double), ("trim_dist", ctypes.c_double), ("tanh_steepness", ctypes.c_double),
        ("tolerance", ctypes.c_double), ("match_ratio_thresh", ctypes.c_double),
    ]


_lib = None
_lock = threading.Lock()
_P, _I, _Z = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
_U64, _D = ctypes.c_ulonglong, ctypes.c_double

# int hook(void* user, int iteration, const double* state_k, int32_t* idx_k)   (include/dicp_b200.h: dicp_nn_hook)
NN_HOOK = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p)

_SIGNATURES = {
    "dicp_version": ([], ctypes.c_int),
    "dicp_last_error": ([], ctypes.c_char_p),
    "dicp_nn_workspace_bytes": ([_I, _I, _I, _I], _Z),
    "dicp_icp_workspace_bytes": ([_I, _I, _I, _I], _Z),
    "dicp_icp_backward_workspace_bytes": ([_I, _I, _I, _I], _Z),
    "dicp_nn_packed_bytes": ([_I, _I, _I], _Z),
    "dicp_kabsch_workspace_bytes": ([_I, _I], _Z),
    "dicp_kabsch_icp_workspace_bytes": ([_I, _I, _I, _I], _Z),
    "dicp_nn_grid_workspace_bytes": ([_I, _I, _I, _I], _Z),
}
for _s in ("f32", "f64"):
    _SIGNATURES["dicp_nn_argmin_" + _s] = ([_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _Z, _P], _I)
    _SIGNATURES["dicp_nn_argmin_grid_" + _s] = ([_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _Z, _P], _I)
    _SIGNATURES["dicp_nn_pack_" + _s] = ([_P, _I, _I, _I, _P, _P, _P], _I)
    _SIGNATURES["dicp_nn_scan_" + _s] = ([_P, _P, _P, _I, _I, _I, _P, _P, _P, _Z, _P], _I)
    _SIGNATURES["dicp_gather_rows_" + _s] = ([_P, _P, _I, _I, _I, _I, _P, _P], _I)
    _SIGNATURES["dicp_gather_rows_bwd_" + _s] = ([_P, _P, _I, _I, _I, _I, _P, _P], _I)
    _SIGNATURES["dicp_icp_forward_" + _s] = ([_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, ctypes.POINTER(DicpParams),
                                             _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                             ctypes.POINTER(ctypes.c_int32), _P, _Z, _P], _I)
    _SIGNATURES["dicp_icp_forward_hooked_" + _s] = (_SIGNATURES["dicp_icp_forward_" + _s][0] + [NN_HOOK, _P], _I)
    _SIGNATURES["dicp_icp_backward_" + _s] = ([_P, _P, _P, _P, _I, _I, _I, _I, ctypes.POINTER(DicpParams), _I,
                                              _P, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P], _I)
    _SIGNATURES["dicp_kabsch_sums_" + _s] = ([_P, _P, _P, _I, _I, _I, _I, _P, _P, _Z, _P], _I)
    _SIGNATURES["dicp_kabsch_sums_bwd_" + _s] = ([_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P], _I)
    _SIGNATURES["dicp_kabsch_icp_forward_" + _s] = ([_P, _P, _P, _I, _I, _I, _I, _I, _D, _I, _I, _P, _P, _P, _P,
                                                     ctypes.POINTER(ctypes.c_int32), _P, _Z, _P], _I)
    _SIGNATURES["dicp_kabsch_icp_backward_" + _s] = ([_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P], _I)
    _SIGNATURES["dicp_gumbel_nn_" + _s] = ([_P, _P, _I, _I, _I, _I, _U64, _D, _D, _P, _P, _P], _I)
    _SIGNATURES["dicp_gumbel_nn_bwd_" + _s] = ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _U64, _D, _D, _P, _P, _P, _P], _I)
    _SIGNATURES["dicp_gumbel_noise_" + _s] = ([_I, _I, _I, _U64, _D, _P, _P], _I)
_SIGNATURES["dicp_gn_accumulate_f32"] = ([_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, ctypes.POINTER(DicpParams), _P, _P, _P], _I)
_SIGNATURES["dicp_gn_accumulate_f64"] = ([_P, _P, _P, _P, _P, _I, _I, _I, _I, ctypes.POINTER(DicpParams), _P, _P, _P], _I)
_SIGNATURES["dicp_nn_push_keys_f32"] = ([_P, _P, _I, _Z, ctypes.POINTER(ctypes.c_uint64), _I, _I, _P, _P], _I)
_SIGNATURES["dicp_nn_unpack_keys_f32"] = ([_P, _Z, _P, _P, _P], _I)
for _s in ("f32", "f64"):
    _SIGNATURES["dicp_solve_update_" + _s] = ([_P, _I, _P, _I, ctypes.POINTER(DicpParams), _P, _P, _P, _P], _I)
_SIGNATURES["dicp_gn_blocks_per_pair"] = ([_I, _I], _I)
_SIGNATURES["dicp_probe_fp32_flops"] = ([_I, _P, _P], ctypes.c_double)
_SIGNATURES["dicp_gn_rows_bytes"] = ([_I, _I, _I], _Z)
_SIGNATURES["dicp_gn_pack_rows_f32"] = ([_P, _I, _I, _I, _I, _P, _P], _I)

EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))


def load_library():
    """dlopen the in-tree library and type its entry points (no GPU needed for this)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    "dicp_b200: %s not found. Build it with `python -m dicp_b200.build` (needs nvcc). "
                    "There is no CPU fallback." % LIB_PATH)
            lib = ctypes.CDLL(LIB_PATH)
            for name, (args, res) in _SIGNATURES.items():
                fn = getattr(lib, name)
                fn.argtypes = args
                fn.restype = res
            _lib = lib
    return _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("dicp_b200: a CUDA device is required (sm_100a kernels); there is no CPU fallback.")


def suffix(dtype):
    if dtype == torch.float32:
        return "f32"
    if dtype == torch.float64:
        return "f64"
    raise TypeError("dicp_b200 supports float32 and float64 tensors, got %s" % dtype)


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def check(rc, what):
    if rc != 0:
        msg = load_library().dicp_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (rc=%d): %s" % (what, rc, msg))
