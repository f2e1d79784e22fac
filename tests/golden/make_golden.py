"""Generate golden vectors by running the UNMODIFIED reference (utiasASRL/dICP).

Run in the build container only (the reference is mounted read-only at /root/reference
and does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports `dICP` from /root/reference (with an empty `matplotlib` stub, because
`dICP/ICP.py:8` imports pyplot, which is not installed and is unused on the compute
path), runs the reference's own fixture workloads (`tests/test_ICP.py`,
`tests/test_ICP_inputs.py`, `tests/test_nn.py`) and a few seeded synthetic dense
workloads, and stores inputs and outputs under tests/golden/*.npz.  No reference
source is copied; only numbers are committed.
"""
import os
import sys
import types
import tempfile

import numpy as np
import torch
import yaml

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

# --- import the reference, and only the reference, as `dICP` -------------------------
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
sys.path.insert(0, REF)
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
import dICP  # noqa: E402
assert os.path.abspath(dICP.__file__).startswith(REF), dICP.__file__
from dICP.ICP import ICP  # noqa: E402
from dICP.nn import nn as RefNN  # noqa: E402
from dICP.loss import loss as RefLoss  # noqa: E402

sys.path.insert(0, ROOT)
from oracle.dicp_oracle import make_wavy_pair  # noqa: E402  (input generator only)

torch.manual_seed(0)
np.random.seed(0)
torch.set_num_threads(8)


def ref_config(**overrides):
    """Copy of the reference YAML with some keys overridden, written to a temp file."""
    with open(os.path.join(REF, "config", "dICP_config.yaml")) as f:
        cfg = yaml.safe_load(f)
    for k, v in overrides.items():
        for section in cfg["dICP"].values():
            if k in section:
                section[k] = v
    fd, path = tempfile.mkstemp(suffix=".yaml")
    with os.fdopen(fd, "w") as f:
        yaml.safe_dump(cfg, f)
    return path


class IndexRecorder:
    """Wraps the reference nn object's find_nn to log the hard-NN indices it used."""

    def __init__(self, icp):
        self.log = []
        self._orig = icp.nn.find_nn
        icp.nn.find_nn = self

    def __call__(self, x, y):
        out = self._orig(x, y)
        xq = x.detach().transpose(1, 2) if x.shape[1] == 3 and x.shape[2] != 3 else x.detach()
        if xq.shape[-1] != 3:
            xq = xq.transpose(1, 2)
        idx = torch.argmin(torch.cdist(xq, y.detach()[:, :, :3], p=2), dim=2)
        # the gathered rows must be what the reference returned
        chk = torch.gather(y.detach(), 1, idx.unsqueeze(-1).expand(-1, -1, y.shape[-1]))
        assert torch.equal(chk, out.detach())
        self.log.append(idx.to(torch.int32).numpy())
        return out


def npy(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


def loss_of(res):
    """Backward seed: every entry of T's top 3 rows and every pc coordinate matters."""
    T, pc = res["T"], res["pc"]
    wT = torch.arange(1, 13, dtype=T.dtype).reshape(3, 4)
    wp = torch.linspace(-1.0, 1.0, pc.shape[1] * 3, dtype=pc.dtype).reshape(pc.shape[1], 3)
    return (T[:, :3, :] * wT).sum() + 0.01 * (pc * wp).sum()


def pack_result(res, prefix=""):
    out = {
        prefix + "T": npy(res["T"]), prefix + "pc": npy(res["pc"]),
        prefix + "costs": npy(res["costs"]), prefix + "deltas": npy(res["deltas"]),
        prefix + "weights": npy(res["weights"]),
        prefix + "converged": npy(res["stats"]["converged"]),
        prefix + "iterations": npy(res["stats"]["iterations"]),
        prefix + "matched_ratio": npy(res["stats"]["matched_ratio"]),
    }
    return out


def save(name, **arrays):
    arrays = {k: v for k, v in arrays.items() if v is not None}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


scan = np.load(os.path.join(REF, "tests", "data", "points_scan.npy"))
cmap = np.load(os.path.join(REF, "tests", "data", "points_map.npy"))
save("fixture_clouds", scan=scan, map=cmap)


# --- A. single-pair fixture workloads (tests/test_ICP.py + BASELINE config 1) ---------
def fixture_case(name, icp_type, diff, loss_fn, trim, max_it, tol, ncols_t, rows_s=None, rows_t=None,
                 szap=False, pad_rows=0):
    src = torch.tensor(scan[:rows_s, :3], requires_grad=True)
    tgt = torch.tensor(cmap[:rows_t, :ncols_t], requires_grad=True)
    src_in = src
    if pad_rows:
        src_in = torch.cat((src, torch.zeros((pad_rows, 3))))
    T0 = torch.eye(4, dtype=src.dtype, requires_grad=True)
    icp = ICP(icp_type=icp_type, differentiable=diff, max_iterations=max_it, tolerance=tol)
    icp.source_zeroes_are_pad = szap
    rec = IndexRecorder(icp)
    res = icp.icp(src_in, tgt, T0, trim_dist=trim, loss_fn=loss_fn, dim=2)
    loss_of(res).backward()
    save(name, source=npy(src_in), target=npy(tgt), T_init=npy(T0),
         nn_idx=np.stack(rec.log, 0),
         g_source=npy(src.grad), g_target=npy(tgt.grad), g_T_init=npy(T0.grad),
         **pack_result(res))


fixture_case("fix_pt2pt_diff_huber", "pt2pt", True, {"name": "huber", "metric": 1.0}, 5.0, 100, 1e-10, 3)
fixture_case("fix_pt2pl_diff_huber", "pt2pl", True, {"name": "huber", "metric": 10.0}, 5.0, 100, 1e-10, 6)
fixture_case("fix_pt2pt_nondiff_huber", "pt2pt", False, {"name": "huber", "metric": 10.0}, 5.0, 100, 1e-10, 3)
fixture_case("fix_pt2pl_diff_cauchy", "pt2pl", True, {"name": "cauchy", "metric": 0.5}, 5.0, 100, 1e-10, 6)
fixture_case("fix_pt2pl_nondiff_huber_sub", "pt2pl", False, {"name": "huber", "metric": 1.0}, 5.0, 25, 1e-8, 6, 50, 55)
fixture_case("fix_pt2pl_diff_huber_sub", "pt2pl", True, {"name": "huber", "metric": 1.0}, 5.0, 25, 1e-8, 6, 50, 55)
fixture_case("fix_pt2pl_nondiff_cauchy_sub", "pt2pl", False, {"name": "cauchy", "metric": 0.5}, 5.0, 25, 1e-8, 6, 50, 55)
fixture_case("fix_pt2pt_nondiff_notrim_sub", "pt2pt", False, None, None, 25, 1e-8, 6, 50, 55, szap=True)
fixture_case("fix_pt2pt_nondiff_notrim_pad", "pt2pt", False, None, None, 25, 1e-8, 6, 50, 55, szap=True, pad_rows=20)


# --- B. ragged / empty / weighted list inputs (tests/test_ICP_inputs.py) ---------------
def ragged_case():
    s1 = torch.tensor(scan[:50, :3])
    s1 = torch.cat((s1, torch.tensor([[431.2, 87.5, 903.1]], dtype=s1.dtype)), dim=0)   # fixed outlier
    t1 = torch.tensor(cmap[:55, :], requires_grad=True)
    s2 = torch.tensor(scan[:, :3], requires_grad=True)
    t2 = torch.tensor(cmap[:, :], requires_grad=True)
    s3 = torch.tensor(scan[:55, :3], requires_grad=True)
    t3 = torch.tensor(cmap[:60, :], requires_grad=True)
    T0 = torch.stack([torch.eye(4, dtype=s1.dtype)] * 3)
    icp = ICP(icp_type="pt2pl", differentiable=True, max_iterations=25, tolerance=1e-8)
    res = icp.icp([s1, s2, s3], [t1, t2, t3], T0, trim_dist=5.0, loss_fn={"name": "huber", "metric": 1.0}, dim=2)
    (res["T"][:, :3, :] * torch.arange(1, 13, dtype=s1.dtype).reshape(3, 4)).sum().backward()
    padded = icp.batch_size_handling([s1, s2, s3], [t1, t2, t3], T0, None)
    save("ragged_pt2pl", s1=npy(s1), s2=npy(s2), s3=npy(s3), t1=npy(t1), t2=npy(t2), t3=npy(t3),
         g_s2=npy(s2.grad), g_s3=npy(s3.grad), g_t1=npy(t1.grad), g_t2=npy(t2.grad), g_t3=npy(t3.grad),
         pad_source=npy(padded[0]), pad_target=npy(padded[1]), pad_w=npy(padded[3]),
         **pack_result(res))


def weights_case():
    s1 = torch.tensor(scan[:, :3], requires_grad=True)
    s3 = torch.tensor(np.vstack((scan[:, :3], np.random.rand(10, 3))), requires_grad=True)
    t = torch.tensor(cmap, requires_grad=True)
    w2 = torch.tensor(np.ones(cmap.shape[0]), requires_grad=True)
    w3 = torch.tensor(np.hstack((np.linspace(0.5, 1.0, cmap.shape[0]), np.zeros(10))), requires_grad=True)
    T0 = torch.stack([torch.eye(4, dtype=s1.dtype)] * 3)
    icp = ICP(icp_type="pt2pl", differentiable=True, max_iterations=25, tolerance=1e-8)
    res = icp.icp([s1, s1, s3], [t, t, t], T0, weight=[None, w2, w3], trim_dist=5.0,
                  loss_fn={"name": "huber", "metric": 1.0}, dim=2)
    (res["T"][:, :3, :] * torch.arange(1, 13, dtype=s1.dtype).reshape(3, 4)).sum().backward()
    save("weights_pt2pl", s1=npy(s1), s3=npy(s3), t=npy(t), w2=npy(w2), w3=npy(w3),
         g_s1=npy(s1.grad), g_s3=npy(s3.grad), g_t=npy(t.grad), g_w2=npy(w2.grad), g_w3=npy(w3.grad),
         **pack_result(res))


def zero_case():
    s = torch.tensor(scan)
    t = torch.tensor(cmap)
    T0 = torch.stack([torch.eye(4, dtype=s.dtype)] * 3)
    icp = ICP(icp_type="pt2pl", differentiable=True, max_iterations=25, tolerance=1e-8)
    res = icp.icp([s, [], []], [[], t, []], T0, trim_dist=5.0, loss_fn=None, dim=2)
    save("zero_inputs", **pack_result(res))


ragged_case()
weights_case()
zero_case()


# --- C. dense synthetic dim=3 workloads (SURVEY.md §8(d) generator, small) --------------
def dense_case(name, icp_type, diff, loss_fn, trim, dtype, N, n, m, K, const_iter, tol, seed,
               with_weight=False, T_perturb=False):
    src, tgt = make_wavy_pair(N, n, m, seed, dtype=dtype, with_normals=True)
    if icp_type == "pt2pt":
        tgt = tgt[:, :, :3].contiguous()
    src.requires_grad_(True)
    tgt.requires_grad_(True)
    T0 = torch.eye(4, dtype=dtype).repeat(N, 1, 1)
    if T_perturb:
        c, s = np.cos(0.01), np.sin(0.01)
        T0[:, :3, :3] = torch.tensor([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=dtype)
        T0[:, :3, 3] = torch.tensor([0.05, -0.02, 0.01], dtype=dtype)
    T0.requires_grad_(True)
    w = None
    if with_weight:
        gen = torch.Generator().manual_seed(seed + 7)
        w = (0.25 + 0.75 * torch.rand((N, n), generator=gen, dtype=torch.float64)).to(dtype)
        w.requires_grad_(True)
    path = ref_config(const_iter=const_iter)
    icp = ICP(config_path=path, icp_type=icp_type, differentiable=diff, max_iterations=K, tolerance=tol)
    rec = IndexRecorder(icp)
    res = icp.icp(src, tgt, T0, weight=w, trim_dist=trim, loss_fn=loss_fn, dim=3)
    loss_of(res).backward()
    os.unlink(path)
    save(name, source=npy(src), target=npy(tgt), T_init=npy(T0), weight=npy(w),
         nn_idx=np.stack(rec.log, 0),
         g_source=npy(src.grad), g_target=npy(tgt.grad), g_T_init=npy(T0.grad),
         g_weight=npy(w.grad) if w is not None else None,
         **pack_result(res))


CAUCHY = {"name": "cauchy", "metric": 0.5}
HUBER = {"name": "huber", "metric": 0.3}
dense_case("dense_pt2pl_cauchy_trim_f64", "pt2pl", True, CAUCHY, 2.0, torch.float64, 3, 300, 350, 6, True, 1e-12, 11,
           with_weight=True, T_perturb=True)
dense_case("dense_pt2pl_cauchy_trim_f32", "pt2pl", True, CAUCHY, 2.0, torch.float32, 3, 300, 350, 6, True, 1e-12, 11,
           with_weight=True, T_perturb=True)
dense_case("dense_pt2pt_huber_trim_f64", "pt2pt", True, HUBER, 2.0, torch.float64, 2, 280, 300, 5, True, 1e-12, 12,
           with_weight=True)
dense_case("dense_pt2pt_huber_trim_f32", "pt2pt", True, HUBER, 2.0, torch.float32, 2, 280, 300, 5, True, 1e-12, 12,
           with_weight=True)
dense_case("dense_pt2pl_nondiff_huber_f64", "pt2pl", False, HUBER, 1.0, torch.float64, 2, 256, 256, 30, False, 1e-9, 13)
dense_case("dense_pt2pt_plain_earlystop_f64", "pt2pt", True, None, None, torch.float64, 3, 200, 220, 40, False, 1e-9, 14)
dense_case("dense_pt2pl_plain_f64", "pt2pl", True, None, None, torch.float64, 1, 128, 128, 4, True, 1e-12, 15)


# --- D. nn.find_nn and loss.get_weight known answers ------------------------------------
def nn_cases():
    pts = torch.tensor([(5.0, 4.0, 0.0), (2.0, 6.0, 0.0), (13.0, 3.0, 0.0), (8.0, 7.0, 0.0), (3.0, 1.0, 0.0)])
    q = torch.tensor([[9.0, 4.0, 0.0]])
    hard = RefNN(differentiable=True, use_gumbel=False)
    a1 = hard.find_nn(q, pts)
    pts2 = torch.cat((pts, torch.tensor([[10.0, 2.0, 0.0]])))
    a2 = hard.find_nn(q, pts2)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn((2, 70, 3), generator=gen, dtype=torch.float64) * 3
    y = torch.randn((2, 90, 6), generator=gen, dtype=torch.float64) * 3
    y.requires_grad_(True)
    out = RefNN(differentiable=False, use_gumbel=False).find_nn(x, y)
    gw = torch.randn(out.shape, generator=gen, dtype=torch.float64)
    (out * gw).sum().backward()
    out_cf = hard.find_nn(x.transpose(1, 2), y.detach())          # channel-first query (N,3,n)
    save("nn_known_answers", pts=npy(pts), q=npy(q), a1=npy(a1), pts2=npy(pts2), a2=npy(a2),
         x=npy(x), y=npy(y), out=npy(out), gw=npy(gw), g_y=npy(y.grad), out_cf=npy(out_cf))


def loss_cases():
    gen = torch.Generator().manual_seed(6)
    e3 = torch.randn((2, 40, 3), generator=gen, dtype=torch.float64) * 2
    e1 = torch.randn((2, 40, 1), generator=gen, dtype=torch.float64) * 2
    e2d = torch.randn((40, 3), generator=gen, dtype=torch.float64) * 2
    out = {"e3": npy(e3), "e1": npy(e1), "e2d": npy(e2d)}
    for name, metric in (("huber", 1.3), ("cauchy", 0.7), ("trim", 1.9)):
        for diff in (True, False):
            L = RefLoss(name=name, metric=metric, differentiable=diff, tanh_steepness=5.0)
            out["%s_%d_e3" % (name, diff)] = npy(L.get_weight(e3))
            out["%s_%d_e1" % (name, diff)] = npy(L.get_weight(e1))
            out["%s_%d_e2d" % (name, diff)] = npy(L.get_weight(e2d))
    save("loss_known_answers", **out)


def svd_case():
    """ICP.py:533-591 on the planar fixture (hard NN).  The reference reaches its known answer here."""
    import warnings
    warnings.filterwarnings("ignore")
    icp = ICP(icp_type="pt2pt", max_iterations=50, tolerance=1e-20)
    icp.nn.use_gumbel = False
    src = torch.tensor(scan[:, :3])
    tgt = torch.tensor(cmap[:, :3])
    ps, T = icp.pt2pt_dICP_SVD(src, tgt, torch.eye(4, dtype=src.dtype))
    save("svd_fixture", source=npy(src), target=npy(tgt), pc=npy(ps), T=npy(T))


nn_cases()
loss_cases()
svd_case()
print("done")
