"""Per-kernel GPU time of the cfg2-pre rollout (Linear preprocessor + temporal chain) through forward_sequence."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "graph-conv-memory_b200"), ROOT]
import torch
import bench
B, K, N, F, H = 65536, 20, 128, 32, 32
dev = torch.device("cuda:0")
mod = bench.build_dense(dev, N, F, H, [("temporal", (1, 2, 4), "forward")], pre=True)
gen = torch.Generator().manual_seed(7)
x = torch.randn(K, B, F, generator=gen).to(dev)
with torch.no_grad():
    _, hidden = mod.forward_sequence(torch.randn(N + 8, B, F, device=dev), None, time_major=True)
    for _ in range(3):
        out, hidden = mod.forward_sequence(x, hidden, time_major=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out, hidden = mod.forward_sequence(x, hidden, time_major=True); e1.record(); torch.cuda.synchronize()
    print(f"call: {e0.elapsed_time(e1) * 1e3 / K:.2f} us per step")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        out, hidden = mod.forward_sequence(x, hidden, time_major=True)
        torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
tot = sum(r[2] for r in rows)
print(f"total kernel time {tot/1e3:.3f} ms for {K} steps")
for k, n, t in sorted(rows, key=lambda r: -r[2])[:12]:
    print(f"| `{k[:90]}` | {n} | {t/1e3:.3f} ms | {100*t/tot:.1f} % |")
