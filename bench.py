#!/usr/bin/env python
"""bench.py — ICP pairs/s (forward + backward, 10 constant iterations, 16k points), learning mode.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2] "C3" per GPU; configs[4] "C5" = 8 GPUs x this): per GPU 256 cloud pairs x 16384
source x 16384 target points, point-to-plane, Cauchy(0.5) + trim(2.0) weights, fp32, dim 3, T_init = I, const_iter,
10 iterations, forward + backward, synthetic wavy-surface clouds.  One "step" = one pass of the hot path over one such
batch in LEARNING MODE: a per-point weight model (WeightModel: small MLP on the source coordinates + a
64^3 voxel table, 0.26 M parameters) feeds `weight=`, the loss weighs the rows of T, the backward reaches the model
(through grad_w, the 124-byte variant of the reverse point pass), the input clouds and T_init.  Pairs are independent,
so N GPUs process N x 256 pairs with no data-path collective (weak scaling); one NCCL all-reduce of
[loss | flattened parameter gradients] (1.05 MB) closes every step.

`value`      : pairs/s with inputs resident in HBM (CUDA events on the launch stream, max over ranks).
`e2e`        : the same step from pinned HOST buffers through the public API: every step's clouds and T_init are copied
               host->device inside the timed region (double-buffered: the copy of step i+1 is enqueued on a side stream
               before the kernels of step i, as an input pipeline would), and [loss, dloss/dtheta] and T are read back.
`roofline`   : the dominant kernel (nearest-neighbour scan, FP32 CUDA-core bound); `frac` is the EXECUTED figure (3 FFMA
               per evaluation, what ncu counts) against the nominal peak, next to the algorithmic 8-FLOP figure and a
               measured FFMA issue peak (dicp_probe_fp32_flops).  `roofline_gn` / `roofline_bwd`: the two HBM-bound
               per-correspondence kernels.
`cpu_baseline` (N = 1 only) and `--impl reference`: the UNMODIFIED reference from baseline/_ref on the host cores, same
               step (same weight model), a bounded sample of the workload.
`theta_only` : round 1's step (weight=None, 13-float all-reduce) for continuity.
`c4_split`   (N > 1): BASELINE configs[3] with the targets of the one pair split over the N GPUs.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PAIRS_PER_GPU = 256
NPTS = 16384
ITERS = 10
LOSS = {"name": "cauchy", "metric": 0.5}
TRIM = 2.0
WORKLOAD = ("C3: %d pairs/GPU x %dx%d pts, pt2pl, cauchy(0.5)+trim(2.0), fp32, dim3, const 10 it, fwd+bwd"
            % (PAIRS_PER_GPU, NPTS, NPTS))
METRIC = "ICP pairs/sec (fwd+bwd, 10 iters, 16k pts)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "sm_max_mhz": float(p.get("sm_max_mhz", 1965.0)), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_batch(n_pairs, first_pair, dtype=torch.float32):
    """seed = 1000*cfg_id + first pair index of the chunk (cfg_id 3)."""
    from dicp_b200.synthetic import surface_pairs
    return surface_pairs(n_pairs, NPTS, NPTS, seed=3000 + first_pair, dtype=dtype, with_normals=True)


def learning_loss(res, theta):
    """Learning-mode objective: theta (12) weighs the top three rows of every T."""
    return (res["T"][:, :3, :].reshape(-1, 12) * theta).sum()


class _PointLinear(torch.autograd.Function):
    """y = x @ w + b over millions of points.  torch's own backward computes grad_w = x^T g as ONE GEMM with K = all points,
    for which cuBLAS picks a SIMT kernel that takes 1 ms at 4 M points; here the reduction over the points is split into
    chunks of 4096 (a batched GEMM) and summed, which takes 0.1 ms.  need_gx: whether a gradient w.r.t. x is wanted."""

    @staticmethod
    def forward(ctx, x, w, b, need_gx):
        ctx.save_for_backward(x, w)
        ctx.need_gx = need_gx
        return x @ w + b

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        x2, g2 = x.reshape(-1, x.shape[-1]), g.reshape(-1, g.shape[-1])
        M = x2.shape[0]
        chunk = 4096
        main = (M // chunk) * chunk
        gw = x2.new_zeros((x2.shape[1], g2.shape[1]))
        if main:
            gw = gw + torch.bmm(x2[:main].reshape(-1, chunk, x2.shape[1]).transpose(1, 2), g2[:main].reshape(-1, chunk, g2.shape[1])).sum(0)
        if main < M:
            gw = gw + x2[main:].t() @ g2[main:]
        gx = (g @ w.t()) if ctx.need_gx else None
        return gx, gw, g2.sum(0), None


class WeightModel(torch.nn.Module):
    """The learnable part of learning mode: a per-point prior weight for `weight=`,
        w_i = sigmoid(MLP(p_i / extent) + table[voxel(p_i)] + 2),
    a 3 -> 16 -> 1 MLP on the source coordinates plus a learnable 64^3 voxel table: 262 144 + 81 parameters, i.e.
    1.05 MB of gradients in the all-reduce, at a per-point cost that stays small next to the ICP itself.  Plain torch
    (it is the model being trained, not the hot path); identical in both arms (seeded init on the CPU)."""

    def __init__(self, grid=64, hidden=16, extent=24.0):
        super().__init__()
        gen = torch.Generator().manual_seed(1234)
        self.grid, self.extent = grid, extent
        self.table = torch.nn.Parameter(0.1 * torch.randn(grid ** 3, generator=gen))
        self.w1 = torch.nn.Parameter(torch.randn((3, hidden), generator=gen) / 3 ** 0.5)
        self.b1 = torch.nn.Parameter(torch.zeros(hidden))
        self.w2 = torch.nn.Parameter(torch.randn((hidden, 1), generator=gen) / hidden ** 0.5)
        self.b2 = torch.nn.Parameter(torch.zeros(1))

    def forward(self, pts):
        x = pts / self.extent
        h = torch.tanh(_PointLinear.apply(x, self.w1, self.b1, False))
        logit = _PointLinear.apply(h, self.w2, self.b2, True).squeeze(-1)
        cell = ((x * 0.5 + 0.5).clamp(0.0, 1.0 - 1e-6) * self.grid).long()
        lin = (cell[..., 0] * self.grid + cell[..., 1]) * self.grid + cell[..., 2]
        # index_select: its backward is an atomic index_add_ (the sort-based backward of advanced indexing costs 1 ms here)
        return torch.sigmoid(logit + self.table.index_select(0, lin.reshape(-1)).reshape(lin.shape) + 2.0)


# ---------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference from baseline/_ref (baseline/install_ref.py),
# driven through its own public API and stock code path (dICP/ICP.py:46-303, dICP/nn.py:11-40).
# Only when baseline/_ref is absent: the oracle port of the same path, labelled kind "port".
# ---------------------------------------------------------------------------------------
def bench_config():
    """The `config` object of the JSON line; identical in both arms."""
    return {"workload": WORKLOAD, "pairs_per_gpu": PAIRS_PER_GPU, "points": NPTS, "iterations": ITERS,
            "learning_mode": "weight= from WeightModel (3-16-1 MLP + 64^3 voxel table, 262225 parameters); loss = sum(theta * T[:, :3, :]); "
                             "backward to the model, the clouds and T_init; one all-reduce of [loss | flat gradients] per step",
            "l2": "inputs larger than L2 (source+target %.0f MB per GPU, plus 168 MB of saved correspondences "
                  "written and re-read per step)" % (PAIRS_PER_GPU * NPTS * 9 * 4 / 1e6)}


_CPU_ARM = {}


def cpu_arm():
    """-> (kind, step(src, tgt)) of the CPU implementation of the path: the reference itself, else the port."""
    if _CPU_ARM:
        return _CPU_ARM["kind"], _CPU_ARM["step"]
    from baseline import ref_arm
    ref = ref_arm.load_reference()
    theta = torch.arange(1, 13, dtype=torch.float32)
    model = WeightModel()
    if ref is not None:
        cfg_path = ref_arm.reference_config(const_iter=True)
        icp = ref.ICP(config_path=cfg_path, icp_type="pt2pl", max_iterations=ITERS, tolerance=1e-12, differentiable=True)

        def step(src, tgt):
            s = src.clone().requires_grad_(True)
            t = tgt.clone().requires_grad_(True)
            model.zero_grad(set_to_none=True)
            res = icp.icp(s, t, torch.eye(4).repeat(src.shape[0], 1, 1), weight=model(src), trim_dist=TRIM, loss_fn=LOSS, dim=3)
            learning_loss(res, theta).backward()
            assert s.grad is not None and t.grad is not None and model.table.grad is not None and res["deltas"].shape[1] == ITERS
            return float(res["T"].detach().sum())
        kind = "reference"
    else:
        from oracle.dicp_oracle import IcpSettings, icp_forward
        cfg = IcpSettings(icp_type="pt2pl", max_iterations=ITERS, tolerance=1e-12, differentiable=True, const_iter=True,
                          trim_dist=TRIM, loss_name=LOSS["name"], loss_metric=LOSS["metric"], dim=3)

        def step(src, tgt):
            s = src.clone().requires_grad_(True)
            t = tgt.clone().requires_grad_(True)
            model.zero_grad(set_to_none=True)
            res = icp_forward(s, t, torch.eye(4).repeat(src.shape[0], 1, 1), model(src), cfg)
            learning_loss(res, theta).backward()
            return float(res["T"].detach().sum())
        kind = "port"
    _CPU_ARM.update(kind=kind, step=step)
    return kind, step


CPU_ARM_DESCRIPTION = {
    "reference": "unmodified utiasASRL/dICP from baseline/_ref: ICP(...).icp through its stock path (cdist + argmin + gather, "
                 "bmm / linalg.inv / matrix_exp, torch autograd), const_iter via a copy of its own YAML",
    "port": "baseline/_ref is absent: oracle port of dICP/ICP.py:132-303 (same torch CPU ops)",
}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    kind, cpu_step = cpu_arm()
    pairs_per_step = 1           # dense (n,m) distance matrix: 1.07 GB fp32 per pair (the reference must be chunked over N)
    src, tgt = make_batch(pairs_per_step, 0)
    for _ in range(args.warmup):
        cpu_step(src, tgt)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(src, tgt)
    dt = time.perf_counter() - t0
    value = pairs_per_step * args.steps / dt
    sample = "%d pair(s) of the C3 workload per step, %d steps, torch %s CPU, %d threads" % (
        pairs_per_step, args.steps, torch.__version__, torch.get_num_threads())
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": bench_config(),
        "reference_arm": CPU_ARM_DESCRIPTION[kind],
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def time_kernel(fn, reps, warm=3):
    """Average duration of one launch: `reps` back-to-back launches between two CUDA events on the launch stream
    (after warm-up, synchronised on both sides), repeated three times, best average."""
    for _ in range(warm):
        fn()
    best = None
    for _ in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) / reps * 1e-3
        best = sec if best is None else min(best, sec)
    return best


def measured_fp32_peak(dev):
    """TFLOP/s a loop of independent FFMAs reaches on this device (dicp_probe_fp32_flops, CUDA events)."""
    from dicp_b200 import _lib
    lib = _lib.load_library()
    sink = torch.zeros(1, device=dev)
    st = _lib.stream_ptr(dev)
    flop = [0.0]

    def launch():
        flop[0] = lib.dicp_probe_fp32_flops(4096, _lib.ptr(sink), st)

    sec = time_kernel(launch, reps=5)
    return flop[0] / sec / 1e12 if flop[0] > 0 else None


def roofline_nn(src, tgt, pk):
    """NN scan kernel alone (dicp_nn_scan_f32), CUDA events on the launch stream."""
    from dicp_b200 import _lib
    lib = _lib.load_library()
    N, n, m = src.shape[0], src.shape[1], tgt.shape[1]
    dev = src.device
    packed = torch.empty(lib.dicp_nn_packed_bytes(N, m, 4), dtype=torch.uint8, device=dev)
    ws = torch.empty(max(lib.dicp_nn_workspace_bytes(N, n, m, 4), 256), dtype=torch.uint8, device=dev)
    idx = torch.empty((N, n), dtype=torch.int32, device=dev)
    st = _lib.stream_ptr(dev)
    _lib.check(lib.dicp_nn_pack_f32(_lib.ptr(tgt), N, m, 6, None, _lib.ptr(packed), st), "pack")

    def launch():
        _lib.check(lib.dicp_nn_scan_f32(_lib.ptr(src), None, _lib.ptr(packed), N, n, m, _lib.ptr(idx), None, _lib.ptr(ws), ws.numel(), st), "scan")

    sec = time_kernel(launch, reps=5)
    evals = float(N) * n * m
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    peak = sms * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
    algorithmic = 8.0 * evals / sec / 1e12   # 3 sub, 3 mul, 2 add per (query, target) evaluation
    executed = 6.0 * evals / sec / 1e12      # what the filter scan issues: 3 FFMA per evaluation (refine + fallback < 2 % more)
    probe = measured_fp32_peak(dev)
    return {"kernel": "nn_scan_f32_kernel<8> + exact fallback (one dicp_nn_scan_f32 call)", "bound": "fp32",
            "achieved": executed, "peak": peak, "unit": "TFLOP/s", "frac": executed / peak,
            "achieved_algorithmic": algorithmic, "frac_algorithmic": algorithmic / peak,
            "peak_measured_ffma": probe, "frac_of_measured_ffma_peak": executed / probe if probe else None,
            "traffic": ncu_traffic("nn_scan", (N, n, m)), "ms_per_launch": sec * 1e3,
            "peak_source": "nominal: %d SMs x 128 lanes x 2 x sm_max_mhz (%s; MEASURED_PEAKS.json has no FP32 CUDA-core figure); "
                           "peak_measured_ffma: independent FFMA chains timed in this run" % (sms, pk["source"]),
            "flop_per_pair_eval": {"algorithmic": 8, "executed": 6}, "units_per_launch": evals}, idx


def ncu_traffic(kernel, shape):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, read from the committed summary of this
    build's `ncu --set full` captures of exactly the launches the probes below time (profiles/r02_ncu_probe_kernels.md, made by
    tools/ncu_summary.py from the .ncu-rep files; C3 shape only) -- null for any other shape or when the file is missing."""
    if tuple(shape) != (256, 16384, 16384):
        return None
    path = os.path.join(ROOT, "profiles", "r02_ncu_probe_kernels.md")
    if not os.path.exists(path):
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, inside = 0.0, False
    with open(path) as f:
        for line in f:
            if line.startswith("## "):
                inside = line[3:].strip() == kernel
            elif inside and line.startswith("| dram__bytes_"):
                cells = [c.strip() for c in line.strip().strip("|").split("|")]
                total += float(cells[1]) * unit.get(cells[2], 1.0)
    return total or None


def roofline_gn(src, tgt, idx, pk):
    """GN accumulate kernel alone (dicp_gn_accumulate_f32), CUDA events on the launch stream."""
    from dicp_b200 import _functions, _lib
    lib = _lib.load_library()
    N, n, m = src.shape[0], src.shape[1], tgt.shape[1]
    dev = src.device
    p = _functions.make_params("pt2pl", 3, True, LOSS["name"], LOSS["metric"], TRIM, 5.0, 1e-12, ITERS, True, 0.0)
    C_r = torch.cat((torch.eye(3, dtype=torch.float64).reshape(1, 9).repeat(N, 1), torch.zeros((N, 3), dtype=torch.float64)), 1).to(dev)
    part = torch.empty((N, lib.dicp_gn_blocks_per_pair(N, n), 32), dtype=torch.float64, device=dev)
    w = torch.ones((N, n), dtype=torch.float32, device=dev)
    wout = torch.empty((N, n), dtype=torch.float32, device=dev)
    st = _lib.stream_ptr(dev)
    # the aligned gather table the loop builds once per ICP call (not part of the per-iteration launch)
    rows = torch.empty(lib.dicp_gn_rows_bytes(N, m, 4), dtype=torch.uint8, device=dev)
    _lib.check(lib.dicp_gn_pack_rows_f32(_lib.ptr(tgt), N, m, 6, p.icp_type, _lib.ptr(rows), st), "gn_pack_rows")

    def launch():
        _lib.check(lib.dicp_gn_accumulate_f32(_lib.ptr(src), _lib.ptr(tgt), _lib.ptr(rows), _lib.ptr(w), _lib.ptr(idx), _lib.ptr(C_r),
                                              N, n, m, 6, ctypes.byref(p), _lib.ptr(wout), _lib.ptr(part), st), "gn")

    sec = time_kernel(launch, reps=20)
    nbytes = 48.0 * N * n                    # p 12 + gathered (q,n) 24 + idx 4 + w_init 4 + weights write 4
    ach = nbytes / sec / 1e9
    traffic = ncu_traffic("gn_accumulate", (N, n, m))
    return {"kernel": "gn_accum_staged_kernel<pt2pl,3> (one dicp_gn_accumulate_f32 call = the loop's single launch)", "bound": "hbm",
            "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
            "traffic": traffic,
            # the same launch by the bytes DRAM actually moved (ncu capture of this build; 32-byte gather records, 128-byte fills)
            "frac_by_dram_traffic": traffic / sec / 1e9 / pk["hbm_gbs"] if traffic else None,
            "ms_per_launch": sec * 1e3, "peak_source": pk["source"], "bytes_per_correspondence": 48,
            "units_per_launch": float(N) * n, "pairs": N}


def roofline_bwd(src, tgt, pk, with_weight):
    """Backward point pass (bwd_points_staged_kernel, HBM bound): the reverse loop of one ICP call is timed through the
    C entry dicp_icp_backward_f32 at K = 10 and K = 5 saved iterations; the difference / 5 is one reverse iteration
    (one bwd_pair_kernel + one bwd_points kernel launch), free of the once-per-call seed / table passes."""
    from dicp_b200 import _functions, _lib
    lib = _lib.load_library()
    N, n, m = src.shape[0], src.shape[1], tgt.shape[1]
    dev = src.device
    p = _functions.make_params("pt2pl", 3, True, LOSS["name"], LOSS["metric"], TRIM, 5.0, 1e-12, ITERS, True, 0.0)
    s = src.detach().clone().requires_grad_(True)
    t = tgt.detach().clone().requires_grad_(True)
    w = torch.ones((N, n), device=dev, requires_grad=True) if with_weight else None
    T0 = torch.eye(4, device=dev).repeat(N, 1, 1)
    outs = _functions.ICPFunction.apply(s, t, T0, w, p)
    ctx = outs[1].grad_fn
    source, target, w_init, saved_idx, saved_state = ctx.saved_tensors
    g_T = torch.zeros((N, 4, 4), device=dev)
    g_T[:, :3, :] = torch.arange(1, 13, dtype=torch.float32, device=dev).reshape(3, 4)
    g_src, g_tgt, g_T0 = torch.empty_like(source), torch.empty_like(target), torch.empty((N, 4, 4), device=dev)
    g_w = torch.empty((N, n), device=dev) if with_weight else None
    st = _lib.stream_ptr(dev)
    ws = torch.empty(lib.dicp_icp_backward_workspace_bytes(N, n, m, 4), dtype=torch.uint8, device=dev)

    def launch(k):
        _lib.check(lib.dicp_icp_backward_f32(_lib.ptr(source), _lib.ptr(target), _lib.ptr(w_init) if with_weight else None, None, N, n, m, 6,
                                             ctypes.byref(p), k, _lib.ptr(saved_idx), _lib.ptr(saved_state), _lib.ptr(g_T), None,
                                             _lib.ptr(g_src), _lib.ptr(g_tgt), _lib.ptr(g_w), _lib.ptr(g_T0), _lib.ptr(ws), ws.numel(), st), "bwd")

    t10 = time_kernel(lambda: launch(ITERS), reps=5)
    t5 = time_kernel(lambda: launch(ITERS // 2), reps=5)
    sec = (t10 - t5) / (ITERS - ITERS // 2)
    per = 124.0 if with_weight else 112.0
    # read p 12 + row 24 + idx 4 (+ w_init 4), RMW grad_source 24 (+ grad_w 8), atomic RMW grad_target 48
    nbytes = per * N * n
    ach = nbytes / sec / 1e9
    traffic = ncu_traffic("bwd_points_124B" if with_weight else "bwd_points_112B", (N, n, m))
    return {"kernel": "bwd_points_staged_kernel<pt2pl> + bwd_pair_kernel, one reverse iteration", "bound": "hbm", "achieved": ach,
            "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
            "traffic": traffic, "frac_by_dram_traffic": traffic / sec / 1e9 / pk["hbm_gbs"] if traffic else None,
            "ms_per_launch": sec * 1e3, "bytes_per_correspondence": per, "with_weight_gradient": with_weight, "units_per_launch": float(N) * n, "pairs": N,
            "whole_backward_ms": t10 * 1e3, "peak_source": pk["source"]}


def run_ours(args):
    import torch.distributed as dist
    from dicp_b200.ICP import ICP          # `dICP.ICP.ICP` of the drop-in alias package is this class
    from dicp_b200 import parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()

    src_h, tgt_h = make_batch(PAIRS_PER_GPU, rank * PAIRS_PER_GPU)
    src_h, tgt_h = src_h.pin_memory(), tgt_h.pin_memory()
    T0_h = torch.eye(4).repeat(PAIRS_PER_GPU, 1, 1).pin_memory()
    icp = ICP(icp_type="pt2pl", differentiable=True, max_iterations=ITERS, tolerance=1e-12)
    icp.const_iter = True
    icp.sync_free = True        # a constant-iteration loop is only enqueued: no host synchronisation inside icp()
    icp.device = dev
    theta = torch.arange(1, 13, dtype=torch.float32, device=dev).requires_grad_(True)
    model = WeightModel().to(dev)
    params = list(model.parameters()) + [theta]
    n_params = sum(p.numel() for p in params)

    src_d = src_h.to(dev).requires_grad_(True)
    tgt_d = tgt_h.to(dev).requires_grad_(True)
    T0_d = T0_h.to(dev)

    def learning_step(s, t, T0):
        """weight model -> ICP forward -> loss -> backward (model, clouds, T_init) -> one all-reduce of [loss | flat grads]"""
        for p_ in params:
            p_.grad = None
        res = icp.icp(s, t, T0, weight=model(s.detach()), trim_dist=TRIM, loss_fn=LOSS, dim=3)
        loss = learning_loss(res, theta)
        loss.backward()
        return parallel.allreduce_loss_and_grads(loss, params), res

    def step_resident():
        src_d.grad = tgt_d.grad = None
        return learning_step(src_d, tgt_d, T0_d)[0]

    def step_theta_only():
        """round 1's step: weight=None, all-reduce of [loss, dloss/dtheta] (13 floats)"""
        src_d.grad = tgt_d.grad = theta.grad = None
        res = icp.icp(src_d, tgt_d, T0_d, trim_dist=TRIM, loss_fn=LOSS, dim=3)
        loss = learning_loss(res, theta)
        loss.backward()
        vec = torch.cat((loss.detach().reshape(1), theta.grad))
        if world > 1:
            dist.all_reduce(vec, op=dist.ReduceOp.SUM)
        return vec

    # e2e: pinned host buffers -> device, double-buffered on a side stream (the copy of step i+1 is enqueued before the
    # kernels of step i), public API, read-back of the reduced loss and of T
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [{"src": torch.empty_like(src_d.detach()), "tgt": torch.empty_like(tgt_d.detach()), "T0": torch.empty_like(T0_d),
              "ready": torch.cuda.Event(), "free": torch.cuda.Event()} for _ in range(2)]

    def enqueue_copy(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(slot["free"])          # the step that last used this slot has finished
            slot["src"].copy_(src_h, non_blocking=True)
            slot["tgt"].copy_(tgt_h, non_blocking=True)
            slot["T0"].copy_(T0_h, non_blocking=True)
            slot["ready"].record(copy_stream)

    # D2H: every step's result ([loss | dloss/dtheta] and T) goes to pinned host memory with an asynchronous copy; the host
    # waits for it one step later (the loss it logs lags by a step, as in any pipelined training loop), so it never drains
    # the launch queue -- with the grid search the step is short enough for that to cost 10 % otherwise
    back = [{"vec": torch.empty(13, dtype=torch.float32).pin_memory(), "T": torch.empty((PAIRS_PER_GPU, 4, 4), dtype=torch.float32).pin_memory(),
             "done": torch.cuda.Event()} for _ in range(2)]

    def e2e_loop(steps):
        cur = torch.cuda.current_stream(dev)
        for sl in slots:
            sl["free"].record(cur)
        enqueue_copy(slots[0])
        for i in range(steps):
            sl, bk = slots[i % 2], back[i % 2]
            if i + 1 < steps:
                enqueue_copy(slots[(i + 1) % 2])
            cur.wait_event(sl["ready"])
            sl["src"].grad = sl["tgt"].grad = None
            s = sl["src"].requires_grad_(True)
            t = sl["tgt"].requires_grad_(True)
            loss, res = learning_step(s, t, sl["T0"])
            assert s.grad is not None and t.grad is not None
            bk["vec"].copy_(torch.cat((loss.reshape(1), theta.grad)), non_blocking=True)            # D2H: the step's result
            bk["T"].copy_(res["T"].detach(), non_blocking=True)
            bk["done"].record(cur)
            sl["src"].requires_grad_(False)
            sl["tgt"].requires_grad_(False)
            sl["free"].record(cur)
            if i > 0:
                back[(i - 1) % 2]["done"].synchronize()                                             # the previous step's result is on the host
        back[(steps - 1) % 2]["done"].synchronize()
        return back[(steps - 1) % 2]["vec"].clone(), back[(steps - 1) % 2]["T"].clone()

    def timed(fn, steps, warmup, sampler=None, loop=False):
        if loop:
            fn(max(1, warmup))
        else:
            for _ in range(warmup):
                fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if loop:
            fn(steps)
        else:
            for _ in range(steps):
                fn()
        e1.record()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) * 1e-3, clocks

    sampler = ClockSampler(local) if rank == 0 else None
    sec, clocks = timed(step_resident, args.steps, args.warmup, sampler)
    steps_e2e = max(3, args.steps)       # as many steps as the resident run: the unoverlapped first copy of the pipeline is amortised alike
    sec_e2e, _ = timed(e2e_loop, steps_e2e, 1, loop=True)
    steps_aux = max(3, args.steps // 2)
    sec_theta, _ = timed(step_theta_only, steps_aux, 1)
    pairs_total = PAIRS_PER_GPU * world
    # the same workload with the exact uniform-grid search (identical correspondences)
    icp.nn_search = "grid"
    sec_grid, _ = timed(step_resident, steps_aux, 1)
    sec_grid_e2e, _ = timed(e2e_loop, steps_aux, 1, loop=True)
    # the theta-only grid step (ICP forward + loss + backward, no weight model, no collective) eager and as ONE CUDA-graph
    # replay per step (dicp_b200.graphs.GraphedStep: a sync-free call contains nothing that cannot be stream-captured)
    from dicp_b200.graphs import GraphedStep

    def grid_theta_step(s, t, T0):
        res = icp.icp(s, t, T0, trim_dist=TRIM, loss_fn=LOSS, dim=3)
        return torch.autograd.grad(learning_loss(res, theta.detach()), [s, t]) + (res["T"],)

    sec_grid_theta, _ = timed(lambda: grid_theta_step(src_d, tgt_d, T0_d), steps_aux, 1)
    sec_grid_graph = None
    if world == 1:      # single process only: a capture must not race with the CUDA calls of a communicator's watchdog thread
        try:
            graphed = GraphedStep(grid_theta_step, (src_d, tgt_d, T0_d))
            sec_grid_graph, _ = timed(graphed.replay, steps_aux, 1)
            del graphed
        except Exception as exc:      # noqa: BLE001 - a secondary figure must never take the bench line down
            sys.stderr.write("bench.py: CUDA-graph replay arm skipped: %r\n" % (exc,))
            torch.cuda.synchronize()
    icp.nn_search = "brute"

    # the collective alone: [loss | flat gradients] SUM all-reduce, as the step issues it
    coll = {"elements": 1 + n_params, "bytes": 4 * (1 + n_params), "ms": 0.0, "backend": "nccl" if world > 1 else "none (single GPU)"}
    if world > 1:
        buf = torch.zeros(1 + n_params, device=dev)
        for _ in range(3):
            dist.all_reduce(buf)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dist.all_reduce(buf)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        coll["ms"] = float(ms.item())

    # BASELINE configs[3] with the targets of the one pair split over the N GPUs
    c4 = c4_split_record(icp, dev, rank, world) if world > 1 else None

    if rank == 0:
        roof, idx = roofline_nn(src_d.detach(), tgt_d.detach(), pk)
        roof_gn = roofline_gn(src_d.detach(), tgt_d.detach(), idx, pk)
        # the same kernels at C5-on-one-GPU size, 2048 pairs = 33.5 M correspondences (measure the
        # HBM fraction where launch ramp and tail do not dominate)
        big = 8
        big_src, big_tgt = src_d.detach().repeat(big, 1, 1), tgt_d.detach().repeat(big, 1, 1)
        keep = ("achieved", "frac", "ms_per_launch", "units_per_launch")      # (traffic fields exist for the C3 shape only)
        roof_gn["at_%d_pairs" % (big * PAIRS_PER_GPU)] = {k: v for k, v in roofline_gn(big_src, big_tgt, idx.repeat(big, 1), pk).items() if k in keep}
        del big_src, big_tgt
        roof_bwd = roofline_bwd(src_d.detach(), tgt_d.detach(), pk, True)
        roof_bwd["without_weight_gradient"] = {k: v for k, v in roofline_bwd(src_d.detach(), tgt_d.detach(), pk, False).items()
                                               if k in keep + ("bytes_per_correspondence", "traffic")}
        mid = 4
        roof_bwd["at_%d_pairs" % (mid * PAIRS_PER_GPU)] = {
            k: v for k, v in roofline_bwd(src_d.detach().repeat(mid, 1, 1), tgt_d.detach().repeat(mid, 1, 1), pk, True).items() if k in keep}
        cpu = None
        if world == 1:    # rank 0 at N=1 only: a bounded sample (8 x 1 pair, 10-30 s of CPU work); at N>1 the ranks waiting
            # in the barrier would starve it, and the driver's own reference arm covers those N
            cores = os.cpu_count()
            torch.set_num_threads(cores)
            kind, cpu_step = cpu_arm()
            s1, t1 = src_h[:1].clone(), tgt_h[:1].clone()
            cpu_step(s1, t1)                        # warm-up (thread pool, allocator)
            t0 = time.perf_counter()
            reps = 8
            for _ in range(reps):
                cpu_step(s1, t1)
            cdt = time.perf_counter() - t0
            cpu = {"value": reps / cdt, "unit": "pairs/s", "cores": cores, "kind": kind,
                   "sample": "%d x 1 pair of the C3 workload (fwd+bwd, 10 it, learning mode), %s, %d threads, %.1f s"
                             % (reps, CPU_ARM_DESCRIPTION[kind], torch.get_num_threads(), cdt)}
        # kernels of libdicp_b200.so per step: forward init, bbox centres, NN pack, gather records + 10 x (filter scan, exact
        # fallback, unpack, GN accumulate, GN update) + finalize; backward gather records, seed + 10 x (pair, points) +
        # final pair + gradient-table fold (memsets and the torch ops of the weight model are not counted)
        launches_per_step = (4 + 5 * ITERS + 1) + (2 + 2 * ITERS + 2)
        h2d = (src_h.numel() + tgt_h.numel() + T0_h.numel()) * 4
        d2h = (PAIRS_PER_GPU * 16 + 13) * 4
        line = {
            "metric": METRIC, "value": pairs_total * args.steps / sec, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(), "pairs_total": pairs_total,
            "learning_mode_allreduce": coll,
            "clocks": clocks,
            "e2e": {"value": pairs_total * steps_e2e / sec_e2e, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * sec_e2e / steps_e2e,
                    "api": "pinned host clouds -> H2D (double-buffered on a side stream) -> WeightModel -> dICP.ICP.ICP.icp -> backward to the "
                           "model, the input points and T_init -> all-reduce -> D2H of [loss, dloss/dtheta] and T into pinned host memory "
                           "(asynchronous, awaited one step later)",
                    "note": "the gradients of the clouds (151 MB) and of the weight model stay on the GPU, where the optimiser of a training "
                            "loop consumes them; a caller who passes HOST tensors to icp() gets every result back on the host (plain dict, as upstream)"},
            "gpu_launches": launches_per_step * args.steps,
            "nn_search": "brute (tiled brute-force scan, the north-star kernel)",
            "sync_free": True,
            "theta_only": {"value": pairs_total * steps_aux / sec_theta, "unit": "pairs/s", "ms_per_step": 1e3 * sec_theta / steps_aux,
                           "note": "round 1's step: weight=None, all-reduce of 13 floats"},
            "alt_nn_grid": {"value": pairs_total * steps_aux / sec_grid, "unit": "pairs/s", "ms_per_step": 1e3 * sec_grid / steps_aux,
                            "e2e": pairs_total * steps_aux / sec_grid_e2e,
                            "theta_only": pairs_total * steps_aux / sec_grid_theta,
                            "theta_only_cuda_graph_replay": pairs_total * steps_aux / sec_grid_graph if sec_grid_graph else None,
                            "note": "same workload with ICP.nn_search='grid': exact uniform-grid NN built once per call, "
                                    "bit-identical correspondences (tests/test_gpu_parity.py::test_icp_grid_mode_equals_brute_mode)"},
            "roofline": roof, "roofline_gn": roof_gn, "roofline_bwd": roof_bwd, "cpu_baseline": cpu, "c4_split": c4,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def c4_split_record(icp_unused, dev, rank, world):
    """BASELINE configs[3] (one 130 000 x 130 000 pair, pt2pl, fwd+bwd, 10 it) with the TARGETS split over the ranks
    (parallel.SplitTargetSearch: slice scans + one MIN merge per iteration over NVLink peer memory or NCCL); rank 0 also runs
    the plain single-GPU call and checks the split run reproduces it bit for bit.  Timing: CUDA events, max over ranks."""
    import torch.distributed as dist
    from dicp_b200.ICP import ICP
    from dicp_b200.parallel import SplitTargetSearch
    from dicp_b200.synthetic import surface_pairs
    src, tgt = surface_pairs(1, 130000, 130000, seed=4000, dtype=torch.float32, extent=60.0)
    src, tgt = src.to(dev), tgt.to(dev)
    T0 = torch.eye(4, device=dev)[None]
    theta = torch.arange(1, 13, dtype=torch.float32, device=dev)

    def step(search):
        icp = ICP(icp_type="pt2pl", differentiable=True, max_iterations=ITERS, tolerance=1e-12)
        icp.const_iter = True
        icp.nn_search = search
        s, t = src.clone().requires_grad_(True), tgt.clone().requires_grad_(True)
        res = icp.icp(s, t, T0, trim_dist=TRIM, loss_fn=LOSS, dim=3)
        (res["T"][:, :3, :].reshape(-1, 12) * theta).sum().backward()
        return res["T"].detach(), s.grad

    def timed(search, reps, warm):
        for _ in range(warm):
            out = step(search)
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            out = step(search)
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    search = SplitTargetSearch(transport="auto")
    ms_split, (T_split, g_split) = timed(search, 4, 2)
    rec = {"workload": "C4: 1 pair x 130000 x 130000, pt2pl, cauchy(0.5)+trim(2.0), fp32, 10 it, fwd+bwd, targets split over %d GPUs" % world,
           "ms_per_step": ms_split, "pairs_per_s": 1e3 / ms_split, "transport": search.used_transport}
    if rank == 0:
        step("brute")
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            T_one, g_one = step("brute")
        b.record()
        torch.cuda.synchronize()
        rec["single_gpu_ms_per_step"] = a.elapsed_time(b) / 3
        rec["speedup"] = rec["single_gpu_ms_per_step"] / ms_split
        rec["bit_identical_to_single_gpu"] = bool(torch.equal(T_one, T_split) and torch.equal(g_one, g_split))
    dist.barrier()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: a CUDA device is required for --impl ours (no CPU fallback)")
        run_ours(args)


if __name__ == "__main__":
    main()
