"""Time the uniform-grid NN against the brute-force scan on the benchmark clouds (C3 shapes)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dicp_b200 import _functions  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
src, tgt = bench.make_batch(N, 0)
src, tgt = src.cuda(), tgt.cuda()
for mode in ("brute", "grid"):
    for _ in range(2):
        idx = _functions.nn_argmin(src, tgt, search=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        idx = _functions.nn_argmin(src, tgt, search=mode)
    e1.record()
    torch.cuda.synchronize()
    print(mode, "ms per search (incl. build/pack):", e0.elapsed_time(e1) / 3)
