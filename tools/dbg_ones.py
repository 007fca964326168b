import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "graph-conv-memory_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import gcm_oracle as oracle
from helpers import make_dense_gnn, make_selector, named_grads, rel_err
from gcm import _cabi
from gcm.gcm import DenseGCM

dev = torch.device("cuda:0")
B, N, F, H, T = 6, 40, 32, 64, 12
spec = [("dense",)]
for acts in (("tanh", "tanh"), ("relu", "none")):
  for bf16 in (False, True):
    gen = torch.Generator().manual_seed(5)
    p = oracle.make_params(F, H)
    nn0 = torch.tensor([33, 0, 40, 17, 39, 5])
    nodes0 = 0.5 * torch.randn(B, N, F, generator=gen)
    adj0 = torch.zeros(B, N, N)
    for b in range(B):
        nodes0[b, int(nn0[b]):] = 0
        adj0[b, : int(nn0[b]), : int(nn0[b])] = 1
    obs = 0.5 * torch.randn(T, B, F, generator=gen)
    w = torch.randn(T, B, H, generator=gen)
    o = obs.double().clone().requires_grad_(True)
    pp = {k: v.double().clone().requires_grad_(True) for k, v in p.items()}
    outs, _ = oracle.dense_gcm_rollout(o, (nodes0.double(), adj0.double(), torch.zeros(0, dtype=torch.float64), nn0.clone()),
                                       spec, pp, acts, graph_size=N)
    (outs * w.double()).sum().backward()
    for x_grad in (True, False):
        gnn, convs = make_dense_gnn(F, H, p, acts)
        mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
        mod.bptt_capacity = T
        if bf16:
            mod.compute_dtype = torch.bfloat16
        x = obs.to(dev).requires_grad_(x_grad)
        hidden = (nodes0.to(dev), adj0.to(dev), torch.zeros(0, device=dev), nn0.to(dev))
        got = []
        for t in range(T):
            belief, hidden = mod(x[t], hidden)
            got.append(belief)
        got = torch.stack(got)
        print(acts, bf16, x_grad, "fwd last kernel", _cabi.lib().gcm_last_kernel().decode(), "plan", mod._plan is not None,
              "rc_bf16", hidden.claim().rc_bf16)
        (got * w.to(dev)).sum().backward()
        print("  bwd last kernel", _cabi.lib().gcm_last_kernel().decode())
        print("  belief err", rel_err(got, outs.detach()))
        if x_grad:
            print("  dx err", rel_err(x.grad, o.grad))
        g = named_grads(convs)
        for k in g:
            print("  ", k, rel_err(g[k], pp[k].grad))
