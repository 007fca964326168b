#!/usr/bin/env python
"""Host-side cost of one DenseGCM.forward call on the fused path (cProfile over many small steps)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "graph-conv-memory_b200"))
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda:0")
mod = bench.build_dense(dev, 128, 32, 32, [("temporal", (1, 2, 4), "forward")])
obs = torch.randn(256, 32, device=dev)
hidden = None
with torch.no_grad():
    for _ in range(200):
        b, hidden = mod(obs, hidden)
    torch.cuda.synchronize()
    n = 5000
    t0 = time.perf_counter()
    for _ in range(n):
        b, hidden = mod(obs, hidden)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"host time per call: {(t1 - t0) / n * 1e6:.2f} us (B=256, queue may throttle)")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        b, hidden = mod(obs, hidden)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
