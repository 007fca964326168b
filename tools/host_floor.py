"""How much of the per-step host time is the C call itself (ctypes marshalling + cudaLaunchKernelEx)?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "graph-conv-memory_b200"), ROOT]
import torch
import bench
from gcm import _cabi

dev = torch.device("cuda:0")
mod = bench.build_dense(dev, 128, 32, 32, [("temporal", (1, 2, 4), "forward")])
obs = torch.randn(64, 32, device=dev)
hidden = None
lib = _cabi.lib()
real = lib.gcm_dense_step_fwd_cached
last = {}

def rec(*a):
    last["a"] = a
    return real(*a)

with torch.no_grad():
    for _ in range(300):
        b, hidden = mod(obs, hidden)
    lib.gcm_dense_step_fwd_cached = rec
    b, hidden = mod(obs, hidden)
    lib.gcm_dense_step_fwd_cached = real
    torch.cuda.synchronize()
    n = 5000
    t0 = time.perf_counter()
    for _ in range(n):
        b, hidden = mod(obs, hidden)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    a = last["a"]
    t2 = time.perf_counter()
    for _ in range(n):
        real(*a)
    t3 = time.perf_counter()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    for _ in range(n):
        torch.empty(64, 32, device=dev)
    t5 = time.perf_counter()
    k = mod.fused_plan().gnn
    t6 = time.perf_counter()
    for _ in range(n):
        k.current_key(dev)
    t7 = time.perf_counter()
print(f"full API call {1e6*(t1-t0)/n:.2f} us | bare C call {1e6*(t3-t2)/n:.2f} us | torch.empty {1e6*(t5-t4)/n:.2f} us | weights key {1e6*(t7-t6)/n:.2f} us")
