"""ONE cfg3 BPTT window (ingest, 64 recorded steps, backward, optimizer step) at full size, nothing else: the
target of the ncu captures under profiles/ (launch list and --set full of the ones-path kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "graph-conv-memory_b200"), ROOT]
import torch
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
windows = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
N, F, H, T = 256, 128, 128, 64
mod = bench.build_dense(dev, N, F, H, [("dense",)])
mod.bptt_capacity = T
mod.compute_dtype = torch.bfloat16
opt = torch.optim.SGD(mod.parameters(), lr=1e-3)
gen = torch.Generator().manual_seed(1003)
obs = (0.5 * torch.randn(T, B, F, generator=gen)).to(dev)
nn0 = torch.full((B,), N - T, dtype=torch.long, device=dev)
nodes0 = 0.5 * torch.randn(B, N, F, device=dev)
nodes0[:, N - T:] = 0
adj0 = torch.zeros(B, N, N, device=dev)
adj0[:, : N - T, : N - T] = 1
for _ in range(windows):
    hidden = (nodes0, adj0, torch.zeros(0, device=dev), nn0)
    opt.zero_grad(set_to_none=True)
    tot = 0
    for t in range(T):
        belief, hidden = mod(obs[t], hidden)
        tot = tot + belief.mean()
    (tot / T).backward()
    opt.step()
torch.cuda.synchronize()
print("done", float(tot))
