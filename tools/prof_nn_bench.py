"""One `dicp_nn_scan_f32` call (filter scan + exact fallback) on the benchmark's C3 surface clouds, as bench.py times it."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
src, tgt = bench.make_batch(N, 0)
r, _ = bench.roofline_nn(src.cuda(), tgt.cuda(), bench.peaks())
print("nn scan call: %.3f ms, %.1f %% executed, %.1f %% algorithmic" % (r["ms_per_launch"], 100 * r["frac"], 100 * r["frac_algorithmic"]))
