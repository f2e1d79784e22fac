"""NN scan time for the benchmark clouds at the origin and shifted 1 km / 100 km away (fallback share)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dicp_b200 import _functions  # noqa: E402

N = 64
src, tgt = bench.make_batch(N, 0)
for off in (0.0, 1000.0, 100000.0):
    s = (src.double() + off).float().cuda()
    t = tgt.clone()
    t[:, :, :3] = (tgt[:, :, :3].double() + off).float()
    t = t.cuda()
    for _ in range(2):
        _functions.nn_argmin(s, t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        _functions.nn_argmin(s, t)
    e1.record()
    torch.cuda.synchronize()
    print("offset %8.0f m: %.3f ms per search (64 pairs x 16k x 16k)" % (off, e0.elapsed_time(e1) / 3))
