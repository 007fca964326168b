import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "graph-conv-memory_b200"))
import torch
import bench
from gcm import _cabi
dev = torch.device("cuda:0")
lib = _cabi.lib()
for B in [int(x) for x in sys.argv[1:]]:
    mod = bench.build_dense(dev, 128, 32, 32, [("temporal", (1, 2, 4), "forward")])
    obs = torch.randn(B, 32, device=dev)
    hidden = None
    with torch.no_grad():
        for t in range(6):
            b, hidden = mod(obs, hidden)
            torch.cuda.synchronize()
            print("B", B, "step", t, lib.gcm_last_kernel().decode(), float(b.abs().mean()), flush=True)
