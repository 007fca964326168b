"""Per-phase timeline of k_temporal_window_bwd tiles (CTA 0, one warp), from the kernel's clock64 trace hook."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "graph-conv-memory_b200"), ROOT]
import torch
from gcm import _cabi
dev = torch.device("cuda:0")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 68 * 65536
lib = _cabi.lib()
X = torch.randn(rows, 64, device=dev) * 0.3
U = torch.randn(rows, 64, device=dev) * 0.3
w1 = torch.randn(32, 64, device=dev) * 0.1
w2 = torch.randn(32, 64, device=dev) * 0.1
b1 = torch.zeros(32, device=dev)
ws = torch.empty(int(lib.gcm_temporal_window_bwd_workspace()), device=dev)
acc = torch.zeros(4160, device=dev)
names = ["st TMEM+arrive", "issue rows", "(gap)", "wait G(prev)", "A' stores", "prefetch", "wait rows MMA", "epilogue+B' stores", "issue G"]
for warp in (0, 5):
    tr = torch.zeros(64, 10, dtype=torch.int64, device=dev)
    lib.gcm_temporal_window_bwd_set_trace(tr.data_ptr(), warp)
    for _ in range(2):
        _cabi.check(lib.gcm_temporal_window_bwd(X.data_ptr(), U.data_ptr(), rows, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), 1,
                                                ws.data_ptr(), acc.data_ptr(), acc[2048:].data_ptr(), acc[4096:].data_ptr(),
                                                acc[4128:].data_ptr(), None, _cabi.stream_ptr(dev)), "wb")
    torch.cuda.synchronize()
    t = tr.cpu().double()
    d = t[8:56, 1:9] - t[8:56, 0:8]
    per = (t[9:57, 0] - t[8:56, 0]).mean()
    print(f"warp {warp}: period {per:.0f} clk per tile")
    m = d.mean(0)
    # phases: 0->1 st, 1->2 issue rows (issuer only), 2->3 wait G, 3->4 A' stores, 4->5 prefetch, 5->6 wait ab, 6->7 epilogue, 7->8 issue G
    for k, nm in enumerate(["st TMEM + arrive", "issue rows (issuer)", "wait G(prev)", "A' stores", "prefetch issue", "wait rows MMA", "epilogue + B' stores + arrive", "issue G (issuer)"]):
        print(f"   {nm:32s} {m[k]:8.0f} clk")
lib.gcm_temporal_window_bwd_set_trace(None, 0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    lib.gcm_temporal_window_bwd(X.data_ptr(), U.data_ptr(), rows, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), 1, ws.data_ptr(),
                                acc.data_ptr(), acc[2048:].data_ptr(), acc[4096:].data_ptr(), acc[4128:].data_ptr(), None,
                                _cabi.stream_ptr(dev))
e1.record(); torch.cuda.synchronize()
print(f"kernel: {e0.elapsed_time(e1) / 5:.3f} ms for {rows} rows")
