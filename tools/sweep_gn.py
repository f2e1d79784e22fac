"""Time the GN accumulate stage for every library variant under variants/ (tools/build_variants.py) and check that
all of them produce the same per-pair sums.  One subprocess per variant (DICP_B200_LIB selects the library).

    python tools/sweep_gn.py [pairs ...]          # default 256 1024
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(pairs):
    import ctypes
    import torch
    sys.path.insert(0, ROOT)
    import bench
    from dicp_b200 import _functions, _lib
    lib = _lib.load_library()
    pk = bench.peaks()
    out = {}
    for N in pairs:
        src, tgt = bench.make_batch(min(N, 256), 0)
        rep = (N + 255) // 256
        src, tgt = src.cuda().repeat(rep, 1, 1)[:N].contiguous(), tgt.cuda().repeat(rep, 1, 1)[:N].contiguous()
        idx = _functions.nn_argmin(src, tgt, search="grid")
        r = bench.roofline_gn(src, tgt, idx, pk)
        # per-pair sums of one launch
        n, m = src.shape[1], tgt.shape[1]
        p = _functions.make_params("pt2pl", 3, True, "cauchy", 0.5, 2.0, 5.0, 1e-12, 10, True, 0.0)
        C_r = torch.cat((torch.eye(3, dtype=torch.float64).reshape(1, 9).repeat(N, 1), torch.zeros((N, 3), dtype=torch.float64)), 1).cuda()
        part = torch.full((N, lib.dicp_gn_blocks_per_pair(N, n), 32), float("nan"), dtype=torch.float64, device="cuda")
        wout = torch.empty((N, n), dtype=torch.float32, device="cuda")
        w = torch.ones((N, n), dtype=torch.float32, device="cuda")
        rows = torch.empty(lib.dicp_gn_rows_bytes(N, m, 4), dtype=torch.uint8, device="cuda")
        st = _lib.stream_ptr(src.device)
        _lib.check(lib.dicp_gn_pack_rows_f32(_lib.ptr(tgt), N, m, 6, p.icp_type, _lib.ptr(rows), st), "pack")
        _lib.check(lib.dicp_gn_accumulate_f32(_lib.ptr(src), _lib.ptr(tgt), _lib.ptr(rows), _lib.ptr(w), _lib.ptr(idx), _lib.ptr(C_r),
                                              N, n, m, 6, ctypes.byref(p), _lib.ptr(wout), _lib.ptr(part), st), "gn")
        torch.cuda.synchronize()
        sums = part.sum(dim=1)
        out[str(N)] = {"us": r["ms_per_launch"] * 1e3, "frac": r["frac"], "rows": part.shape[1],
                       "check": [float(sums[:, :31].abs().sum()), float(sums[N // 2, 0]), float(wout.double().sum())]}
        if os.environ.get("SWEEP_BWD") and N <= 1024:      # one reverse iteration (pair kernel + point pass), with the weight gradient
            rb = bench.roofline_bwd(src, tgt, pk, True)
            out[str(N)]["bwd_us"] = rb["ms_per_launch"] * 1e3
            out[str(N)]["bwd_frac"] = rb["frac"]
    print("RESULT " + json.dumps(out))


def main():
    pairs = [int(a) for a in sys.argv[1:]] or [256, 1024]
    vdir = os.path.join(ROOT, "variants")
    libs = [("default", os.path.join(ROOT, "dicp_b200", "lib", "libdicp_b200.so"))]
    if os.path.isdir(vdir):
        libs += [(v, os.path.join(vdir, v, "libdicp_b200.so")) for v in sorted(os.listdir(vdir)) if os.path.exists(os.path.join(vdir, v, "libdicp_b200.so"))]
    for name, path in libs:
        env = dict(os.environ, DICP_B200_LIB=path)
        pr = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"] + [str(p) for p in pairs], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        res = [l for l in pr.stdout.splitlines() if l.startswith("RESULT ")]
        if not res:
            print("%-12s FAILED\n%s" % (name, pr.stdout[-1500:]))
            continue
        d = json.loads(res[0][7:])
        print("%-12s " % name + "  ".join("N=%s: %6.1f us %.3f rows=%d chk=%s%s" % (
            k, v["us"], v["frac"], v["rows"], ["%.9g" % c for c in v["check"]],
            " bwd %.1f us %.3f" % (v["bwd_us"], v["bwd_frac"]) if "bwd_us" in v else "") for k, v in d.items()), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child([int(a) for a in sys.argv[2:]])
    else:
        main()
