import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "graph-conv-memory_b200"))
import torch
import bench
from gcm import _cabi
dev = torch.device("cuda:0")
lib = _cabi.lib()
B = 65536
mod = bench.build_dense(dev, 128, 32, 32, [("temporal", (1, 2, 4), "forward")])
obs = torch.randn(B, 32, device=dev)
hidden = None
with torch.no_grad():
    for t in range(140):
        b, hidden = mod(obs, hidden)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (12 * 8 * 8))()
fn = ctypes.CDLL(_cabi.lib_path()).gcm_debug_hc_trace
fn.argtypes = [ctypes.c_void_p]
print("rc", fn(buf))
import numpy as np
a = np.array(buf, dtype=np.int64).reshape(12, 8, 8)
t0 = a[a > 0].min()
for w in range(12):
    for it in range(8):
        row = a[w, it]
        if row.max() > 0:
            print("warp", w, "it", it, " ".join(f"{(x - t0) / 1000:7.2f}" if x > 0 else "    -  " for x in row))

buf2 = (ctypes.c_ulonglong * (160 * 2))()
fn2 = ctypes.CDLL(_cabi.lib_path()).gcm_debug_hc_cta
fn2.argtypes = [ctypes.c_void_p]
print("rc", fn2(buf2))
c = np.array(buf2, dtype=np.int64).reshape(160, 2)[:148]
t0 = c[:, 0].min()
print("cta start (us): min %.2f max %.2f" % ((c[:, 0].min() - t0) / 1e3, (c[:, 0].max() - t0) / 1e3))
print("cta end   (us): min %.2f max %.2f" % ((c[:, 1].min() - t0) / 1e3, (c[:, 1].max() - t0) / 1e3))
dur = (c[:, 1] - c[:, 0]) / 1e3
print("cta duration: min %.2f median %.2f max %.2f" % (dur.min(), np.median(dur), dur.max()))
print("slowest CTAs:", np.argsort(-dur)[:10], np.sort(-dur)[:10] * -1)
