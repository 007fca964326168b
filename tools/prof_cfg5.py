"""Per-kernel GPU time of one cfg5 SparseGCM.forward (torch.profiler / CUPTI, no replay)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "graph-conv-memory_b200"), ROOT]
import torch
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
grad = len(sys.argv) > 2 and sys.argv[2] == "grad"
dev = torch.device("cuda:0")
N, F, H = int(os.environ.get("GCM_PROF_N", "4096")), 64, 64      # GCM_PROF_N=256 with B=16384: small graphs, same node count
mod = bench.build_sparse(dev, N, F, H)
gen = torch.Generator().manual_seed(1005)
x = torch.randn(B, N, F, generator=gen)
x[..., 0:2] = torch.cumsum(0.1 * torch.randn(B, N, 2, generator=gen), dim=1)
x = x.to(dev)
taus = torch.full((B,), N, dtype=torch.long, device=dev)


def call():
    if grad:
        out, hid = mod(x, taus, None)
        out.mean().backward()
    else:
        with torch.no_grad():
            out, hid = mod(x, taus, None)


for _ in range(2):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); call(); e1.record(); torch.cuda.synchronize()
print(f"call: {e0.elapsed_time(e1):.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    call()
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
tot = sum(r[2] for r in rows)
print(f"total kernel time {tot/1e3:.2f} ms")
print("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
for k, n, t in sorted(rows, key=lambda r: -r[2])[:22]:
    print(f"| `{k[:80]}` | {n} | {t/1e3:.2f} | {t/n:.1f} | {100*t/tot:.1f} % |")
