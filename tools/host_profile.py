"""cProfile of the per-call rollout path (DenseGCM.forward on a live handle, cfg2 shapes at a small batch)."""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "graph-conv-memory_b200"), ROOT]
import torch
import bench

dev = torch.device("cuda:0")
B, N, F, H = 4096, 128, 32, 32
mod = bench.build_dense(dev, N, F, H, [("temporal", (1, 2, 4), "forward")])
x = torch.randn(B, F, device=dev)
hidden = None
with torch.no_grad():
    for _ in range(200):
        _, hidden = mod(x, hidden)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5000):
        _, hidden = mod(x, hidden)
    pr.disable()
    torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(18)
