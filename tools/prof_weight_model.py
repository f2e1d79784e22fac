import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda')
m = bench.WeightModel().to(dev)
x = (torch.rand(256, 16384, 3, device=dev) * 40 - 20)
g = torch.rand(256, 16384, device=dev)
def step():
    for p in m.parameters(): p.grad = None
    w = m(x); w.backward(g)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
print("model fwd+bwd: %.3f ms" % (e0.elapsed_time(e1) / 10))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
