"""Run only the NN scan (and optionally GN / backward stage kernels) on the C3 shapes — a short
command for ncu captures.   python tools/prof_nn.py [N] [n] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dicp_b200 import _lib  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lib = _lib.load_library()
dev = torch.device("cuda")
g = torch.Generator(device="cpu").manual_seed(0)
tgt = torch.empty((N, n, 6), device=dev).uniform_(-20, 20)
src = tgt[:, torch.randperm(n, generator=g).to(dev), :3].contiguous() + 0.02 * torch.randn((N, n, 3), device=dev)
packed = torch.empty(lib.dicp_nn_packed_bytes(N, n, 4), dtype=torch.uint8, device=dev)
ws = torch.empty(max(lib.dicp_nn_workspace_bytes(N, n, n, 4), 256), dtype=torch.uint8, device=dev)
idx = torch.empty((N, n), dtype=torch.int32, device=dev)
st = _lib.stream_ptr(dev)
_lib.check(lib.dicp_nn_pack_f32(_lib.ptr(tgt), N, n, 6, None, _lib.ptr(packed), st), "pack")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for i in range(reps):
    _lib.check(lib.dicp_nn_scan_f32(_lib.ptr(src), None, _lib.ptr(packed), N, n, n, _lib.ptr(idx), None, _lib.ptr(ws), ws.numel(), st), "scan")
    ev[i + 1].record()
torch.cuda.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
print("nn scan ms per launch:", ts, " TFLOP/s(8/pair):", [8.0 * N * n * n / (t * 1e-3) / 1e12 for t in ts])
