"""Per-warp phase timestamps of CTA 0 of one multi-step k_step_temporal_hc launch (debug build: GCM_NVCC_EXTRA="-DHC_TRACE
-DHC_TRACE_FROM=16").  consumer phases: 0 before full wait, 1 data there, 2 operands built, 3 state written, 4 D1 ready,
5 layer-1 epilogue done (MMA2 issued), 6 D2 ready, 7 belief written; producer: 0 before empty wait, 1 stage free, 2 issued"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "graph-conv-memory_b200"))
import numpy as np
import torch
import bench
from gcm import _cabi
dev = torch.device("cuda:0")
B = 65536
mod = bench.build_dense(dev, 128, 32, 32, [("temporal", (1, 2, 4), "forward")])
x = torch.randn(B, 140, 32, device=dev)
with torch.no_grad():
    _, hidden = mod.forward_sequence(x, None)
    torch.cuda.synchronize()
    _, hidden = mod.forward_sequence(x[:, :32].contiguous(), hidden)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (12 * 8 * 8))()
fn = ctypes.CDLL(_cabi.lib_path()).gcm_debug_hc_trace
fn.argtypes = [ctypes.c_void_p]
print("rc", fn(buf))
a = np.array(buf, dtype=np.int64).reshape(12, 8, 8)
t0 = a[a > 0].min()
for w in range(12):
    for it in range(8):
        row = a[w, it]
        if row.max() > 0:
            print("warp", w, "it", it, " ".join(f"{(v - t0) / 1000:7.2f}" if v > 0 else "    -  " for v in row))
