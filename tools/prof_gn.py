"""Time the GN accumulate stage (dicp_gn_accumulate_f32) and the full backward on C3-shaped data."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dicp_b200 import _functions  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pk = bench.peaks()
src, tgt = bench.make_batch(min(N, 256), 0)
rep = (N + 255) // 256
src, tgt = src.cuda().repeat(rep, 1, 1)[:N].contiguous(), tgt.cuda().repeat(rep, 1, 1)[:N].contiguous()
idx = _functions.nn_argmin(src, tgt, search="grid")
r = bench.roofline_gn(src, tgt, idx, pk)
print("gn_accumulate: %.1f us per call, %.0f GB/s, %.1f %% of measured HBM peak" % (r["ms_per_launch"] * 1e3, r["achieved"], 100 * r["frac"]))
ar = torch.arange(src.shape[1], dtype=torch.int32, device="cuda").repeat(N, 1).contiguous()
r = bench.roofline_gn(src, tgt, ar, pk)
print("  with idx = identity (sequential gather): %.1f us, %.0f GB/s" % (r["ms_per_launch"] * 1e3, r["achieved"]))
