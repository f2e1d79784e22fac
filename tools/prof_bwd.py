"""Time one reverse iteration of the backward (bench.roofline_bwd) with and without the weight gradient."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pk = bench.peaks()
src, tgt = bench.make_batch(min(N, 256), 0)
rep = (N + 255) // 256
src, tgt = src.cuda().repeat(rep, 1, 1)[:N].contiguous(), tgt.cuda().repeat(rep, 1, 1)[:N].contiguous()
for ww in (False, True):
    r = bench.roofline_bwd(src, tgt, pk, ww)
    print("N=%d weight_grad=%s: %.1f us per reverse iteration, %.0f GB/s, %.3f of HBM peak; whole backward %.3f ms" % (
        N, ww, r["ms_per_launch"] * 1e3, r["achieved"], r["frac"], r["whole_backward_ms"]))
