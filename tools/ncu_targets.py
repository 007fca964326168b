"""Small driver for ncu captures of the kernels bench.py's default command does not reach:
  tc       k_step_temporal_tc      (the recomputing tensor-core step kernel: cache fill / after weight updates), cfg2 size
  general  k_step_general          (chained selectors: TemporalBackedge + DenseEdge), B=8192 N=128 F=H=32
  bwd      k_step_bwd_general      (BPTT of a CosineEdge chain), B=2048 N=64 F=H=32, 4 steps
Usage: ncu ... python tools/ncu_targets.py tc|general|bwd"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "graph-conv-memory_b200"), ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import torch
import bench
from gcm import _cabi
from gcm.gcm import DenseGCM
from helpers import make_dense_gnn, make_selector
import gcm_oracle as oracle

what = sys.argv[1]
dev = torch.device("cuda:0")
if what == "tc":
    mod = bench.build_dense(dev, 128, 32, 32, [("temporal", (1, 2, 4), "forward")])
    _cabi.lib().gcm_set_temporal_kernel(_cabi.TK_TC)
    x = torch.randn(140, 65536, 32, device=dev)
    with torch.no_grad():
        h = None
        for t in range(136):
            _, h = mod(x[t], h)
    torch.cuda.synchronize()
elif what == "general":
    p = oracle.make_params(32, 32)
    gnn, _ = make_dense_gnn(32, 32, p, ("tanh", "tanh"))
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector([("temporal", (1, 2, 4), "forward"), ("dense",)]), graph_size=128)
    x = torch.randn(40, 8192, 32, device=dev)
    with torch.no_grad():
        h = None
        for t in range(40):
            _, h = mod(x[t], h)
    torch.cuda.synchronize()
    print(_cabi.lib().gcm_last_kernel().decode())
else:
    p = oracle.make_params(32, 32)
    gnn, _ = make_dense_gnn(32, 32, p, ("tanh", "tanh"))
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector([("cosine", 0.5)]), graph_size=64)
    gen = torch.Generator().manual_seed(1)
    x = bench.synth_obs(gen, 44, 2048, 32, [("cosine", 0.5)]).to(dev)
    with torch.no_grad():
        h = None
        for t in range(40):
            _, h = mod(x[t], h)
    h = h.detach()
    outs = []
    for t in range(40, 44):
        o, h = mod(x[t], h)
        outs.append(o)
    torch.stack(outs).mean().backward()
    torch.cuda.synchronize()
    print(_cabi.lib().gcm_last_kernel().decode())
