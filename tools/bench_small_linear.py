"""Latency of the tensor-core Linear kernels on few rows (the per-step products of the DenseEdge path): us per launch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "graph-conv-memory_b200"), ROOT]
import torch
from gcm import ones, _cabi

dev = torch.device("cuda:0")
lib = _cabi.lib()
for K, Ho in ((128, 128), (64, 64), (32, 32)):
    w = torch.randn(Ho, K, device=dev) / K ** 0.5
    b = torch.randn(Ho, device=dev)
    for rows in (128, 2048, 16384, 131072, 1048576):
        x = torch.randn(rows, K, device=dev)
        out = torch.empty(rows, Ho, device=dev)
        res = {}
        def bf16():
            _cabi.check(lib.gcm_linear_tc(x.data_ptr(), K, K, w.data_ptr(), b.data_ptr(), 0, rows, Ho, out.data_ptr(), Ho, 0,
                                          _cabi.stream_ptr(dev)), "gcm_linear_tc")
        for name, fn in (("tc32", lambda: ones._lin_tc32(x, w, bias=b, act=1, out=out)),
                         ("tc_bf16", bf16),
                         ("lin2", lambda: ones._lin2(x, w, bias=b, act=1, out=out))):
            for _ in range(5):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 100 if rows <= 131072 else 20
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / n * 1e3
        print(f"K={K} Ho={Ho} rows={rows}: tc32 {res['tc32']:.1f} us, tc bf16 {res['tc_bf16']:.1f} us, "
              f"lin2 (CUDA cores) {res['lin2']:.1f} us")
