"""Per-kernel GPU time of one cfg3 BPTT window (torch.profiler / CUPTI, no replay).  Usage: prof_cfg3.py [B] [cache]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "graph-conv-memory_b200"), ROOT]
import torch
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cache = sys.argv[2] if len(sys.argv) > 2 else "bf16"
seq = len(sys.argv) > 3 and sys.argv[3] == "seq"
dev = torch.device("cuda:0")
N, F, H, T = 256, 128, 128, 64
mod = bench.build_dense(dev, N, F, H, [("dense",)])
mod.bptt_capacity = T
mod.compute_dtype = torch.bfloat16 if cache == "bf16" else None
opt = torch.optim.SGD(mod.parameters(), lr=1e-3)
gen = torch.Generator().manual_seed(1003)
obs = (0.5 * torch.randn(T, B, F, generator=gen)).to(dev)
obs_bt = obs.transpose(0, 1).contiguous()
nn0 = torch.full((B,), N - T, dtype=torch.long, device=dev)
nodes0 = 0.5 * torch.randn(B, N, F, device=dev)
nodes0[:, N - T:] = 0
adj0 = torch.zeros(B, N, N, device=dev)
adj0[:, : N - T, : N - T] = 1


def window():
    hidden = (nodes0, adj0, torch.zeros(0, device=dev), nn0)
    opt.zero_grad(set_to_none=True)
    if seq:
        beliefs, hidden = mod.forward_sequence(obs_bt, hidden)
        tot = beliefs.mean() * T
    else:
        tot = 0
        for t in range(T):
            belief, hidden = mod(obs[t], hidden)
            tot = tot + belief.mean()
    (tot / T).backward()
    opt.step()


for _ in range(2):
    window()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); window(); e1.record(); torch.cuda.synchronize()
print(f"window: {e0.elapsed_time(e1):.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    window()
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
tot = sum(r[2] for r in rows)
print(f"total kernel time {tot/1e3:.2f} ms")
print("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
for k, n, t in sorted(rows, key=lambda r: -r[2])[:25]:
    print(f"| `{k[:80]}` | {n} | {t/1e3:.2f} | {t/n:.1f} | {100*t/tot:.1f} % |")
