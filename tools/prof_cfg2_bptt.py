"""Per-kernel GPU time of one cfg2-bptt window (torch.profiler / CUPTI, no replay).  Usage: prof_cfg2_bptt.py [B] [T]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "graph-conv-memory_b200"), ROOT]
import torch
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda:0")
N, F, H = 128, 32, 32
mod = bench.build_dense(dev, N, F, H, [("temporal", (1, 2, 4), "forward")], pre=bool(os.environ.get("GCM_PROF_PRE")))   # GCM_PROF_PRE=1: with RayDenseGCM's Linear preprocessor (its parameters train too)
mod.bptt_capacity = T
opt = torch.optim.SGD(mod.parameters(), lr=1e-3)
gen = torch.Generator().manual_seed(1003)
obs_tb = torch.randn(T, B, F, generator=gen).to(dev)
with torch.no_grad():
    _, carry = mod.forward_sequence(torch.randn(N + 8, B, F, device=dev), None, time_major=True)
carry = [carry]


def window():
    opt.zero_grad(set_to_none=True)
    beliefs, hidden = mod.forward_sequence(obs_tb, carry[0].detach(), time_major=True)
    beliefs.mean().backward()
    opt.step()
    carry[0] = hidden


for _ in range(2):
    window()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); window(); e1.record(); torch.cuda.synchronize()
print(f"window: {e0.elapsed_time(e1):.2f} ms")
import time
torch.cuda.synchronize()
t0 = time.perf_counter()
e0.record()
for _ in range(5):
    window()
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"5 windows back to back: {e0.elapsed_time(e1) / 5:.2f} ms per window on the device, host issue time {(t1 - t0) * 200:.2f} ms per window")
if os.environ.get("GCM_TRACE_JSON"):
    from torch.profiler import profile as _p, ProfilerActivity as _A
    with _p(activities=[_A.CUDA, _A.CPU]) as pr:
        window(); window()
        torch.cuda.synchronize()
    pr.export_chrome_trace(os.environ["GCM_TRACE_JSON"])
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    window()
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
tot = sum(r[2] for r in rows)
print(f"total kernel time {tot/1e3:.2f} ms")
print("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
for k, n, t in sorted(rows, key=lambda r: -r[2])[:25]:
    print(f"| `{k[:90]}` | {n} | {t/1e3:.2f} | {t/n:.1f} | {100*t/tot:.1f} % |")
