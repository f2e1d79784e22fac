#!/usr/bin/env python
"""bench.py — ICP pairs/s (forward + backward, 10 constant iterations, 16k points).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], SURVEY.md §8(d) "C3"): per GPU 256 cloud pairs x 16384
source x 16384 target points, point-to-plane, Cauchy(0.5) + trim(2.0) weights, fp32, dim 3,
T_init = I, const_iter, 10 iterations, forward + backward, synthetic wavy-surface clouds.
One "step" = one pass of the hot path over one such batch.  Pairs are independent, so N GPUs
process N x 256 pairs with no data-path collective (weak scaling; at N=8 this is configs[4]);
in learning mode one NCCL all-reduce of [loss, d loss / d theta] closes every step.

`value`  : pairs/s with inputs resident in HBM (timed with CUDA events on the launch stream).
`e2e`    : the same step (forward + backward) starting from pinned HOST buffers: the host->device
           copies of the clouds and T_init, the call of the public API `dICP.ICP.ICP.icp`, the
           backward to the input points and the device->host read of [loss, dloss/dtheta] and T
           are all inside the timed region.
`roofline`: the dominant kernel (nearest-neighbour scan, FP32 CUDA-core bound) timed alone
           with CUDA events; `roofline_gn`: the Gauss-Newton accumulate (HBM bound).
`cpu_baseline`: the oracle port of the reference's torch-CPU path on a bounded sample.
`--impl reference`: the same oracle port timed alone on all host threads (the reference is
           pure Python/torch and cannot travel to the GPU box).
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


# ---------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference from baseline/_ref (baseline/install_ref.py),
# driven through its own public API and stock code path (dICP/ICP.py:46-303, dICP/nn.py:11-40).
# Only when baseline/_ref is absent: the oracle port of the same path, labelled kind "port".
# ---------------------------------------------------------------------------------------
def bench_config():
    """The `config` object of the JSON line; identical in both arms."""
    return {"workload": WORKLOAD, "pairs_per_gpu": PAIRS_PER_GPU, "points": NPTS, "iterations": ITERS,
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
    if ref is not None:
        cfg_path = ref_arm.reference_config(const_iter=True)
        icp = ref.ICP(config_path=cfg_path, icp_type="pt2pl", max_iterations=ITERS, tolerance=1e-12, differentiable=True)

        def step(src, tgt):
            s = src.clone().requires_grad_(True)
            t = tgt.clone().requires_grad_(True)
            res = icp.icp(s, t, torch.eye(4).repeat(src.shape[0], 1, 1), trim_dist=TRIM, loss_fn=LOSS, dim=3)
            learning_loss(res, theta).backward()
            assert s.grad is not None and t.grad is not None and res["deltas"].shape[1] == ITERS
            return float(res["T"].sum())
        kind = "reference"
    else:
        from oracle.dicp_oracle import IcpSettings, icp_forward
        cfg = IcpSettings(icp_type="pt2pl", max_iterations=ITERS, tolerance=1e-12, differentiable=True, const_iter=True,
                          trim_dist=TRIM, loss_name=LOSS["name"], loss_metric=LOSS["metric"], dim=3)

        def step(src, tgt):
            s = src.clone().requires_grad_(True)
            t = tgt.clone().requires_grad_(True)
            res = icp_forward(s, t, torch.eye(4).repeat(src.shape[0], 1, 1), torch.ones(src.shape[:2]), cfg)
            learning_loss(res, theta).backward()
            return float(res["T"].sum())
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
    flop = 8.0 * evals                       # algorithmic: 3 sub, 3 mul, 2 add per (query, target) evaluation
    peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
    ach = flop / sec / 1e12
    executed = 6.0 * evals / sec / 1e12      # what the filter scan issues: 3 FFMA per evaluation (refine + fallback < 2 % more)
    return {"kernel": "nn_scan_f32_kernel<8> + exact fallback (one dicp_nn_scan_f32 call)", "bound": "fp32",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of nn_scan_f32_kernel<8> at these shapes (profiles/r01_ncu_full_top_kernels.md)
            "traffic": 201.7e6 if (N, n, m) == (256, 16384, 16384) else None, "ms_per_launch": sec * 1e3,
            "executed_tflops": executed, "frac_executed": executed / peak,
            "peak_source": "148 SMs x 128 lanes x 2 x sm_max_mhz (%s MEASURED_PEAKS.json has no FP32 CUDA-core figure)" % pk["source"],
            # context: a loop of independent FFMAs issues at 1.13 cycles per warp instruction and scheduler on this part
            # (tools/ubench_fp32.cu, profiles/r01_nn_tuning.md), i.e. 88 % of the nominal peak is the most FFMA code reaches
            "ffma_issue_peak_measured": peak / 1.13, "frac_executed_of_ffma_issue_peak": executed / (peak / 1.13),
            "flop_per_pair_eval": {"algorithmic": 8, "executed": 6}, "units_per_launch": evals}, idx


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
    return {"kernel": "gn_accum_kernel<float,pt2pl,3> (one dicp_gn_accumulate_f32 call = the loop's single launch)", "bound": "hbm",
            "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
            # dram bytes of gn_accum_kernel at 256 pairs from the ncu capture in profiles/: 201.5 MB read + 15.8 MB written
            "traffic": 217.3e6 if (N, n, m) == (256, 16384, 16384) else None,
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

    def launch(k):
        _lib.check(lib.dicp_icp_backward_f32(_lib.ptr(source), _lib.ptr(target), _lib.ptr(w_init) if with_weight else None, None, N, n, m, 6,
                                             ctypes.byref(p), k, _lib.ptr(saved_idx), _lib.ptr(saved_state), _lib.ptr(g_T), None,
                                             _lib.ptr(g_src), _lib.ptr(g_tgt), _lib.ptr(g_w), _lib.ptr(g_T0), _lib.ptr(ctx.ws), ctx.ws.numel(), st), "bwd")

    t10 = time_kernel(lambda: launch(ITERS), reps=5)
    t5 = time_kernel(lambda: launch(ITERS // 2), reps=5)
    sec = (t10 - t5) / (ITERS - ITERS // 2)
    per = 124.0 if with_weight else 112.0
    # read p 12 + row 24 + idx 4 (+ w_init 4), RMW grad_source 24 (+ grad_w 8), atomic RMW grad_target 48
    nbytes = per * N * n
    ach = nbytes / sec / 1e9
    return {"kernel": "bwd_points_staged_kernel<pt2pl> + bwd_pair_kernel, one reverse iteration", "bound": "hbm", "achieved": ach,
            "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": None, "ms_per_launch": sec * 1e3,
            "bytes_per_correspondence": per, "with_weight_gradient": with_weight, "units_per_launch": float(N) * n, "pairs": N,
            "whole_backward_ms": t10 * 1e3, "peak_source": pk["source"]}


def run_ours(args):
    import torch.distributed as dist
    from dicp_b200.ICP import ICP          # `dICP.ICP.ICP` of the drop-in alias package is this class

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
    icp.device = dev
    theta = torch.arange(1, 13, dtype=torch.float32, device=dev).requires_grad_(True)

    src_d = src_h.to(dev).requires_grad_(True)
    tgt_d = tgt_h.to(dev).requires_grad_(True)
    T0_d = T0_h.to(dev)

    def allreduce(vec):
        if world > 1:
            dist.all_reduce(vec, op=dist.ReduceOp.SUM)
        return vec

    def step_resident():
        """inputs already in HBM; forward + backward; learning-mode all-reduce of [loss, dloss/dtheta]."""
        src_d.grad = tgt_d.grad = theta.grad = None
        res = icp.icp(src_d, tgt_d, T0_d, trim_dist=TRIM, loss_fn=LOSS, dim=3)
        loss = learning_loss(res, theta)
        loss.backward()
        return allreduce(torch.cat((loss.detach().reshape(1), theta.grad)))

    def step_e2e():
        """the same step from pinned HOST buffers: the clouds and T_init are copied to the GPU inside the step (H2D),
        `icp.icp` runs forward, the backward reaches the input points as in step_resident, and the reduced
        [loss, dloss/dtheta] and the transforms are read back to the host (D2H)."""
        theta.grad = None
        s = src_h.to(dev, non_blocking=True).requires_grad_(True)
        t = tgt_h.to(dev, non_blocking=True).requires_grad_(True)
        res = icp.icp(s, t, T0_h.to(dev, non_blocking=True), trim_dist=TRIM, loss_fn=LOSS, dim=3)
        loss = learning_loss(res, theta)
        loss.backward()
        assert s.grad is not None and t.grad is not None
        out = allreduce(torch.cat((loss.detach().reshape(1), theta.grad)))
        return out.cpu(), res["T"].detach().cpu()

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
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
    sec_e2e, _ = timed(step_e2e, max(2, args.steps // 2), 1)
    steps_e2e = max(2, args.steps // 2)
    pairs_total = PAIRS_PER_GPU * world
    # the same workload with the exact uniform-grid search (identical correspondences)
    icp.nn_search = "grid"
    sec_grid, _ = timed(step_resident, args.steps, args.warmup)
    sec_grid_e2e, _ = timed(step_e2e, steps_e2e, 1)
    icp.nn_search = "brute"

    line = None
    if rank == 0:
        roof, idx = roofline_nn(src_d.detach(), tgt_d.detach(), pk)
        roof_gn = roofline_gn(src_d.detach(), tgt_d.detach(), idx, pk)
        # the same kernel at C5-on-one-GPU size, 2048 pairs = 33.5 M correspondences (measure the
        # HBM fraction where launch latency does not dominate)
        big = 8
        roof_gn["at_%d_pairs" % (big * PAIRS_PER_GPU)] = {
            k: v for k, v in roofline_gn(src_d.detach().repeat(big, 1, 1), tgt_d.detach().repeat(big, 1, 1), idx.repeat(big, 1), pk).items()
            if k in ("achieved", "frac", "ms_per_launch", "units_per_launch")}
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
                   "sample": "%d x 1 pair of the C3 workload (fwd+bwd, 10 it), %s, %d threads, %.1f s"
                             % (reps, CPU_ARM_DESCRIPTION[kind], torch.get_num_threads(), cdt)}
        # kernels of libdicp_b200.so per step (profiles/r01_launches.md): forward init, bbox centres, NN pack, gather
        # records + 10 x (filter scan, exact fallback, unpack, GN accumulate, GN update) + finalize;
        # backward seed + 10 x (pair, points) + final pair
        launches_per_step = (4 + 5 * ITERS + 1) + (1 + 2 * ITERS + 1)
        # per step: clouds + T_init up; T and the reduced [loss, dloss/dtheta] down (pc / weights / deltas / costs and
        # the gradients of the clouds stay on the GPU, where a training loop would consume them)
        h2d = (src_h.numel() + tgt_h.numel() + T0_h.numel()) * 4
        d2h = (PAIRS_PER_GPU * 16 + 13) * 4
        line = {
            "metric": METRIC, "value": pairs_total * args.steps / sec, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(), "pairs_total": pairs_total,
            "learning_mode_allreduce": "13 floats per step over NCCL" if world > 1 else "single GPU: no collective",
            "clocks": clocks,
            "e2e": {"value": pairs_total * steps_e2e / sec_e2e, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * sec_e2e / steps_e2e, "api": "pinned host clouds -> H2D -> dICP.ICP.ICP.icp (forward) -> backward to the input points -> D2H of [loss, dloss/dtheta] and T"},
            "gpu_launches": launches_per_step * args.steps,
            "nn_search": "brute (tiled brute-force scan, the north-star kernel)",
            "alt_nn_grid": {"value": pairs_total * args.steps / sec_grid, "unit": "pairs/s", "ms_per_step": 1e3 * sec_grid / args.steps,
                            "e2e": pairs_total * steps_e2e / sec_grid_e2e,
                            "note": "same workload with ICP.nn_search='grid': exact uniform-grid NN built once per call, "
                                    "bit-identical correspondences (tests/test_gpu_parity.py::test_icp_grid_mode_equals_brute_mode)"},
            "roofline": roof, "roofline_gn": roof_gn, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
