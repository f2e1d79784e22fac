"""Experiment: cudaLimitMaxL2FetchGranularity (32 / 64 / 128 B) against the two HBM-bound kernels, whose random
32-byte row gathers pay for the other half of every 64-byte granule.  One subprocess per setting (the limit belongs
to the context).   python tools/exp_l2gran.py"""
import ctypes
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIMIT_L2_FETCH = 0x05       # cudaLimitMaxL2FetchGranularity


def child(gran):
    import torch
    sys.path.insert(0, ROOT)
    import bench
    from dicp_b200 import _functions
    torch.zeros(1, device="cuda")
    import nvidia
    rt = ctypes.CDLL(glob.glob(os.path.join(nvidia.__path__[0], "cuda_runtime", "lib", "libcudart.so*"))[0])
    val = ctypes.c_size_t(0)
    rt.cudaDeviceGetLimit(ctypes.byref(val), LIMIT_L2_FETCH)
    before = val.value
    rc = 0
    if gran > 0:
        rc = rt.cudaDeviceSetLimit(LIMIT_L2_FETCH, ctypes.c_size_t(gran))
    rt.cudaDeviceGetLimit(ctypes.byref(val), LIMIT_L2_FETCH)
    pk = bench.peaks()
    src, tgt = bench.make_batch(256, 0)
    src, tgt = src.cuda(), tgt.cuda()
    idx = _functions.nn_argmin(src, tgt, search="grid")
    out = {"limit_before": before, "limit_now": val.value, "rc": rc}
    out["gn256_us"] = bench.roofline_gn(src, tgt, idx, pk)["ms_per_launch"] * 1e3
    out["gn2048_us"] = bench.roofline_gn(src.repeat(8, 1, 1), tgt.repeat(8, 1, 1), idx.repeat(8, 1), pk)["ms_per_launch"] * 1e3
    out["bwd256_us"] = bench.roofline_bwd(src, tgt, pk, True)["ms_per_launch"] * 1e3
    a = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
    b = torch.empty_like(a)
    out["copy_2GiB_GBs"] = 2 * a.numel() * 4 / bench.time_kernel(lambda: b.copy_(a), reps=5) / 1e9
    print("RESULT " + json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(int(sys.argv[2]))
    else:
        for g in (0, 32, 64, 128):
            pr = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(g)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            res = [l for l in pr.stdout.splitlines() if l.startswith("RESULT ")]
            print("granularity %3d: %s" % (g, res[0][7:] if res else "FAILED\n" + pr.stdout[-1500:]), flush=True)
