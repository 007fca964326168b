"""Host cost of one DenseGCM.forward call on the fused rollout path (cProfile, tiny batch so the GPU never blocks)."""
import cProfile, pstats, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "graph-conv-memory_b200"), ROOT]
import torch
import bench

dev = torch.device("cuda:0")
mod = bench.build_dense(dev, 128, 32, 32, [("temporal", (1, 2, 4), "forward")])
obs = torch.randn(64, 32, device=dev)
hidden = None
with torch.no_grad():
    for _ in range(300):
        b, hidden = mod(obs, hidden)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5000):
        b, hidden = mod(obs, hidden)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"host per call: {(t1 - t0) / 5000 * 1e6:.2f} us")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5000):
        b, hidden = mod(obs, hidden)
    pr.disable()
    torch.cuda.synchronize()
ps = pstats.Stats(pr).sort_stats("tottime")
ps.print_stats(18)
