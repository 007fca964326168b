#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It puts `tests/golden/standin` (a restatement of the few torch_geometric /
torch_scatter / torchtyping symbols the reference imports; none of them is
installed here) and `/root/reference/src` on sys.path, imports the reference's
own `gcm` package, runs it on small seeded inputs and stores inputs + outputs as
`tests/golden/*.pt`.  While doing so it also checks `oracle/gcm_oracle.py`
against the reference on every case (bit-exact for node slots / adjacency /
num_nodes / edge lists, allclose for beliefs and gradients).

The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "standin"))
sys.path.insert(1, "/root/reference/src")
sys.path.insert(2, os.path.join(ROOT, "oracle"))

import torch  # noqa: E402
import torch_geometric  # noqa: E402  (the stand-in)
from gcm.gcm import DenseGCM  # noqa: E402  (the REFERENCE package)
from gcm.sparse_gcm import SparseGCM  # noqa: E402
from gcm.edge_selectors.temporal import TemporalBackedge  # noqa: E402
from gcm.edge_selectors.dense import DenseEdge  # noqa: E402
from gcm.edge_selectors.distance import EuclideanEdge, CosineEdge, SpatialEdge  # noqa: E402
from gcm.sparse_edge_selectors.temporal import TemporalEdge  # noqa: E402
from gcm.sparse_edge_selectors.spatial import SpatialRadiusEdge  # noqa: E402

import gcm_oracle as oracle  # noqa: E402

assert "/root/reference/src" in sys.modules["gcm.gcm"].__file__

torch.set_num_threads(4)
ACT = {"tanh": torch.nn.Tanh, "relu": torch.nn.ReLU, "none": torch.nn.Identity}


class RefDenseGNN(torch.nn.Module):
    """The README's user GNN (README.md:52-62), with configurable activations."""

    def __init__(self, F, H, p, acts):
        super().__init__()
        self.gc0 = torch_geometric.nn.DenseGraphConv(F, H)
        self.gc1 = torch_geometric.nn.DenseGraphConv(H, H)
        self.a0, self.a1 = ACT[acts[0]](), ACT[acts[1]]()
        load(self, p)

    def forward(self, x, adj, weights, B, N):
        x = self.a0(self.gc0(x, adj))
        return self.a1(self.gc1(x, adj))


class RefSparseGNN(torch.nn.Module):
    def __init__(self, F, H, p, acts):
        super().__init__()
        self.gc0 = torch_geometric.nn.GraphConv(F, H)
        self.gc1 = torch_geometric.nn.GraphConv(H, H)
        self.a0, self.a1 = ACT[acts[0]](), ACT[acts[1]]()
        load(self, p)

    def forward(self, x, edges, weights):
        x = self.a0(self.gc0(x, edges, weights))
        return self.a1(self.gc1(x, edges, weights))


def load(m, p):
    with torch.no_grad():
        m.gc0.lin_rel.weight.copy_(p["w_rel1"]); m.gc0.lin_rel.bias.copy_(p["b1"])
        m.gc0.lin_root.weight.copy_(p["w_root1"])
        m.gc1.lin_rel.weight.copy_(p["w_rel2"]); m.gc1.lin_rel.bias.copy_(p["b2"])
        m.gc1.lin_root.weight.copy_(p["w_root2"])


def ref_selector(spec):
    mods = []
    for s in spec:
        if s[0] == "temporal":
            mods.append(TemporalBackedge(list(s[1]), direction=s[2]))
        elif s[0] == "dense":
            mods.append(DenseEdge())
        elif s[0] == "euclidean":
            mods.append(EuclideanEdge(s[1]))
        elif s[0] == "cosine":
            mods.append(CosineEdge(s[1]))
        elif s[0] == "spatial":
            mods.append(SpatialEdge(s[1], s[2], s[3]))
    if not mods:
        return None
    if len(mods) == 1:
        return mods[0]
    return torch_geometric.nn.Sequential(
        "x, adj, weights, num_nodes, B",
        [(m, "x, adj, weights, num_nodes, B -> adj, weights") for m in mods],
    )


def clustered_obs(g, T, B, F, K=4, noise=0.02, shared=True):
    """SURVEY.md §8(d) c4 recipe: K well separated centres, one shared schedule."""
    centres = torch.randn(K, F, generator=g) * 2.0
    sched = torch.randint(0, K, (T,), generator=g)
    obs = centres[sched].unsqueeze(1).expand(T, B, F) + noise * torch.randn(T, B, F, generator=g)
    return obs.contiguous()


def dense_case(name, B, N, F, H, T, spec, acts=("tanh", "tanh"), seed=0, obs=None,
               init=None, grads=False, snap=()):
    g = torch.Generator().manual_seed(1000 + seed)
    if obs is None:
        obs = torch.randn(T, B, F, generator=g)
    p = oracle.make_params(F, H, seed=7 + seed)
    gnn = RefDenseGNN(F, H, p, acts)
    mod = DenseGCM(gnn, edge_selectors=ref_selector(spec), graph_size=N)
    hidden = None if init is None else tuple(t.clone() for t in init)
    o_hidden = None if init is None else tuple(t.clone() for t in init)
    obs_ref = obs.clone().requires_grad_(grads)
    obs_or = obs.clone().requires_grad_(grads)
    p_or = {k: v.clone().requires_grad_(grads) for k, v in p.items()}
    beliefs, snaps = [], {}
    o_beliefs = []
    for t in range(T):
        mx, hidden = mod(obs_ref[t], hidden)
        omx, o_hidden = oracle.dense_gcm_step(obs_or[t], o_hidden, spec, p_or, acts, graph_size=N)
        beliefs.append(mx)
        o_beliefs.append(omx)
        # bit-exact GCM-owned state
        assert torch.equal(hidden[0].detach(), o_hidden[0].detach()), (name, t, "nodes")
        assert torch.equal(hidden[1].float(), o_hidden[1].float()), (name, t, "adj")
        assert torch.equal(hidden[3], o_hidden[3]), (name, t, "num_nodes")
        assert torch.allclose(mx, omx, rtol=1e-5, atol=1e-6), (name, t, (mx - omx).abs().max())
        if t in snap:
            snaps[t] = tuple(h.detach().clone() for h in hidden)
    beliefs = torch.stack(beliefs)
    out = {
        "name": name, "B": B, "N": N, "F": F, "H": H, "T": T, "spec": spec, "acts": acts,
        "obs": obs, "params": p, "init": init,
        "beliefs": beliefs.detach().clone(),
        "final": tuple(h.detach().clone() for h in hidden), "snaps": snaps,
    }
    if grads:
        w = torch.randn(beliefs.shape, generator=g)
        out["loss_w"] = w
        (beliefs * w).sum().backward()
        (torch.stack(o_beliefs) * w).sum().backward()
        out["d_obs"] = obs_ref.grad.clone()
        names = {"w_rel1": gnn.gc0.lin_rel.weight, "b1": gnn.gc0.lin_rel.bias,
                 "w_root1": gnn.gc0.lin_root.weight, "w_rel2": gnn.gc1.lin_rel.weight,
                 "b2": gnn.gc1.lin_rel.bias, "w_root2": gnn.gc1.lin_root.weight}
        out["d_params"] = {k: v.grad.clone() for k, v in names.items()}
        assert torch.allclose(out["d_obs"], obs_or.grad, rtol=1e-4, atol=1e-6), name
        for k in names:
            assert torch.allclose(out["d_params"][k], p_or[k].grad, rtol=1e-4, atol=1e-5), (name, k)
    torch.save(out, os.path.join(HERE, name + ".pt"))
    print("wrote", name, "adj nnz", int(hidden[1].sum()))


def preproc_case(name, B, N, F_raw, F, H, T, spec, seed=0, pre_act=None):
    """The reference DenseGCM WITH a preprocessor (what RayDenseGCM builds, ray_gcm.py:118,133-136): a Linear (+ optional
    activation) applied by the reference to all N rows every step (gcm.py:290-291).  Also pins the equivalence the fused
    path relies on: for a per-row preprocessor the beliefs equal those of the plain step fed the PREPROCESSED
    observations, while the hidden state keeps the RAW ones."""
    g = torch.Generator().manual_seed(2000 + seed)
    obs = torch.randn(T, B, F_raw, generator=g)
    p = oracle.make_params(F, H, seed=7 + seed)
    lin = torch.nn.Linear(F_raw, F)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(F, F_raw, generator=g) / F_raw ** 0.5)
        lin.bias.copy_(0.1 * torch.randn(F, generator=g))
    pre = lin if pre_act is None else torch.nn.Sequential(lin, ACT[pre_act]())
    mod = DenseGCM(RefDenseGNN(F, H, p, ("tanh", "tanh")), preprocessor=pre, edge_selectors=ref_selector(spec), graph_size=N)
    hidden, o_hidden, beliefs = None, None, []
    with torch.no_grad():
        for t in range(T):
            mx, hidden = mod(obs[t], hidden)
            omx, o_hidden = oracle.dense_gcm_step(pre(obs[t]), o_hidden, spec, p, ("tanh", "tanh"), graph_size=N)
            assert torch.allclose(mx, omx, rtol=1e-5, atol=1e-6), (name, t, (mx - omx).abs().max())
            assert torch.equal(hidden[1].float(), o_hidden[1].float()) and torch.equal(hidden[3], o_hidden[3]), (name, t)
            beliefs.append(mx)
    out = {"name": name, "B": B, "N": N, "F_raw": F_raw, "F": F, "H": H, "T": T, "spec": spec, "pre_act": pre_act,
           "obs": obs, "params": p, "pre_weight": lin.weight.detach().clone(), "pre_bias": lin.bias.detach().clone(),
           "beliefs": torch.stack(beliefs).clone(), "final": tuple(h.detach().clone() for h in hidden)}
    torch.save(out, os.path.join(HERE, name + ".pt"))
    print("wrote", name, "adj nnz", int(hidden[1].sum()))


def preproc_grad_case(name, B, N, F_raw, F, H, T0, T1, spec, seed=0):
    """Training through the reference DenseGCM WITH a preprocessor (RayDenseGCM's configuration): T0 steps without
    autograd (the window wraps), the hidden state handed on as plain detached tensors, then a BPTT window of T1 steps.
    The reference maps ALL stored rows through the preprocessor at every step (gcm.py:290-291), so the preprocessor's
    gradients also collect contributions through the rows written BEFORE the window.  Stored: beliefs, dL/dobs of the
    window, gradients of the six GNN tensors and of the preprocessor."""
    g = torch.Generator().manual_seed(2000 + seed)
    obs = torch.randn(T0 + T1, B, F_raw, generator=g)
    w = torch.randn(T1, B, H, generator=g)
    p = oracle.make_params(F, H, seed=7 + seed)
    lin = torch.nn.Linear(F_raw, F)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(F, F_raw, generator=g) / F_raw ** 0.5)
        lin.bias.copy_(0.1 * torch.randn(F, generator=g))
    gnn = RefDenseGNN(F, H, p, ("tanh", "tanh"))
    mod = DenseGCM(gnn, preprocessor=lin, edge_selectors=ref_selector(spec), graph_size=N)
    hidden = None
    with torch.no_grad():
        for t in range(T0):
            _, hidden = mod(obs[t], hidden)
    hidden = tuple(h.detach().clone() for h in hidden)
    start = tuple(h.clone() for h in hidden)
    x = obs[T0:].clone().requires_grad_(True)
    outs = []
    for t in range(T1):
        mx, hidden = mod(x[t], hidden)
        outs.append(mx)
    outs = torch.stack(outs)
    (outs * w).sum().backward()
    d_params = {"w_rel1": gnn.gc0.lin_rel.weight.grad, "b1": gnn.gc0.lin_rel.bias.grad, "w_root1": gnn.gc0.lin_root.weight.grad,
                "w_rel2": gnn.gc1.lin_rel.weight.grad, "b2": gnn.gc1.lin_rel.bias.grad, "w_root2": gnn.gc1.lin_root.weight.grad}
    out = {"name": name, "B": B, "N": N, "F_raw": F_raw, "F": F, "H": H, "T0": T0, "T1": T1, "spec": spec, "obs": obs,
           "loss_w": w, "params": p, "pre_weight": lin.weight.detach().clone(), "pre_bias": lin.bias.detach().clone(),
           "start": start, "beliefs": outs.detach().clone(), "d_obs": x.grad.clone(),
           "d_params": {k: v.clone() for k, v in d_params.items()},
           "d_pre_weight": lin.weight.grad.clone(), "d_pre_bias": lin.bias.grad.clone(),
           "final": tuple(h.detach().clone() for h in hidden)}
    torch.save(out, os.path.join(HERE, name + ".pt"))
    print("wrote", name)


def coo_to_idx(adj):
    return adj.coalesce().indices()


def sparse_case(name, B, N, F, H, calls, spec, aux=None, acts=("tanh", "tanh"), seed=0,
                max_hops=None, grads=False, walk=False):
    """calls: list of taus lists (one SparseGCM.forward per entry)."""
    g = torch.Generator().manual_seed(2000 + seed)
    p = oracle.make_params(F, H, seed=7 + seed)
    gnn = RefSparseGNN(F, H, p, acts)

    def mk(spec):
        if not spec:
            return None
        s = spec[0]
        assert len(spec) == 1
        if s[0] == "temporal":
            return TemporalEdge(list(s[1]))
        return SpatialRadiusEdge(s[1], s[2])

    mod = SparseGCM(gnn, edge_selectors=mk(spec), aux_edge_selectors=mk(aux), graph_size=N,
                    max_hops=max_hops)
    p_or = {k: v.clone().requires_grad_(grads) for k, v in p.items()}
    hidden, o_hidden = None, None
    xs, outs, o_outs, x_or = [], [], [], []
    for taus in calls:
        taus = torch.tensor(taus, dtype=torch.long)
        tmax = int(taus.max())
        x = torch.randn(B, tmax, F, generator=g)
        if walk:
            x[..., 0:2] = torch.cumsum(0.1 * torch.randn(B, tmax, 2, generator=g), dim=1)
        for b in range(B):
            x[b, int(taus[b]):] = 0
        xr = x.clone().requires_grad_(grads)
        xo = x.clone().requires_grad_(grads)
        mx, hidden = mod(xr, taus, hidden)
        omx, o_hidden = oracle.sparse_gcm_forward(xo, taus, o_hidden, spec, p_or, acts,
                                                  graph_size=N, max_hops=max_hops,
                                                  aux_selectors=aux)
        assert torch.equal(hidden[0].detach(), o_hidden[0].detach()), (name, "nodes")
        assert torch.equal(coo_to_idx(hidden[1]), o_hidden[1]), (name, "edges")
        assert torch.equal(hidden[2], o_hidden[2]), (name, "T")
        assert torch.allclose(mx, omx, rtol=1e-5, atol=1e-6), (name, (mx - omx).abs().max())
        xs.append((x, taus, xr, xo))
        outs.append(mx)
        o_outs.append(omx)
    out = {"name": name, "B": B, "N": N, "F": F, "H": H, "spec": spec, "aux": aux,
           "acts": acts, "max_hops": max_hops, "params": p,
           "calls": [(x, taus) for x, taus, _, _ in xs],
           "outs": [o.detach().clone() for o in outs],
           "final_nodes": hidden[0].detach().clone(), "final_edges": coo_to_idx(hidden[1]).clone(),
           "final_T": hidden[2].clone()}
    if grads:
        ws = [torch.randn(o.shape, generator=g) for o in outs]
        out["loss_w"] = ws
        sum((o * w).sum() for o, w in zip(outs, ws)).backward()
        sum((o * w).sum() for o, w in zip(o_outs, ws)).backward()
        out["d_x"] = [xr.grad.clone() for _, _, xr, _ in xs]
        names = {"w_rel1": gnn.gc0.lin_rel.weight, "b1": gnn.gc0.lin_rel.bias,
                 "w_root1": gnn.gc0.lin_root.weight, "w_rel2": gnn.gc1.lin_rel.weight,
                 "b2": gnn.gc1.lin_rel.bias, "w_root2": gnn.gc1.lin_root.weight}
        out["d_params"] = {k: v.grad.clone() for k, v in names.items()}
        for (_, _, xr, xo) in xs:
            assert torch.allclose(xr.grad, xo.grad, rtol=1e-4, atol=1e-6), name
        for k in names:
            assert torch.allclose(out["d_params"][k], p_or[k].grad, rtol=1e-4, atol=1e-5), (name, k)
    torch.save(out, os.path.join(HERE, name + ".pt"))
    print("wrote", name, "E", int(out["final_edges"].shape[1]))


def main():
    # ---- dense -------------------------------------------------------------
    dense_case("dense_temporal1_wrap", B=4, N=8, F=5, H=6, T=20, spec=[("temporal", (1,), "forward")],
               snap=(3, 7, 8, 12))
    dense_case("dense_temporal124_fwd", B=3, N=16, F=4, H=8, T=40,
               spec=[("temporal", (1, 2, 4), "forward")], seed=1, snap=(15, 16, 20))
    dense_case("dense_temporal12_bwd", B=3, N=8, F=4, H=8, T=20,
               spec=[("temporal", (1, 2), "backward")], seed=2, snap=(9,))
    dense_case("dense_temporal13_both", B=3, N=8, F=4, H=8, T=20,
               spec=[("temporal", (1, 3), "both")], seed=3, snap=(9,))
    dense_case("dense_denseedge_wrap", B=3, N=6, F=5, H=7, T=14, spec=[("dense",)], seed=4, snap=(5, 6, 9))
    dense_case("dense_noedge_relu", B=2, N=5, F=3, H=4, T=8, spec=[], acts=("relu", "none"), seed=5)
    dense_case("dense_chain_t1_t2", B=5, N=10, F=11, H=11, T=14,
               spec=[("temporal", (1,), "forward"), ("temporal", (2,), "forward")], seed=6)
    dense_case("dense_chain_t1_dense", B=2, N=6, F=4, H=4, T=9,
               spec=[("temporal", (1,), "backward"), ("dense",)], acts=("tanh", "relu"), seed=7)
    g = torch.Generator().manual_seed(77)
    dense_case("dense_euclid", B=6, N=10, F=7, H=5, T=25, spec=[("euclidean", 1.0)], seed=8,
               obs=clustered_obs(g, 25, 6, 7), snap=(9, 10))
    dense_case("dense_cosine", B=4, N=12, F=6, H=5, T=20, spec=[("cosine", 0.5)], seed=9,
               obs=clustered_obs(g, 20, 4, 6))
    obs = clustered_obs(g, 20, 4, 8)
    dense_case("dense_spatial", B=4, N=9, F=8, H=5, T=20, spec=[("spatial", 1.0, slice(0, 2), None)],
               seed=10, obs=obs)
    dense_case("dense_spatial_ab", B=4, N=9, F=8, H=5, T=20,
               spec=[("spatial", 1.0, slice(0, 3), slice(2, 5))], seed=11, obs=obs * 0.3)
    # user-supplied (clean) hidden with ragged num_nodes, incl. a full graph
    # (tests/test_gcm.py:89-184 shape: N=7, num_nodes [1, 7])
    N, F, B = 7, 5, 3
    nn_ = torch.tensor([1, 7, 4])
    nodes = torch.arange(B * N * F, dtype=torch.float32).reshape(B, N, F) * 0.01
    adj = torch.zeros(B, N, N)
    for b in range(B):
        n = int(nn_[b])
        nodes[b, n:] = 0
        for i in range(1, n):
            adj[b, i, i - 1] = 1
        if n > 2:
            adj[b, 0, n - 1] = 1          # a "future" edge into row 0
            adj[b, n - 1, 0] = 1
    dense_case("dense_userhidden_ragged", B=B, N=N, F=F, H=4, T=10,
               spec=[("temporal", (1, 2), "forward")], seed=12,
               init=(nodes, adj, torch.zeros(0), nn_), snap=(0, 1))
    # gradients (BPTT)
    dense_case("dense_grad_temporal", B=3, N=6, F=4, H=5, T=10,
               spec=[("temporal", (1, 2), "forward")], seed=13, grads=True)
    dense_case("dense_grad_both", B=2, N=8, F=3, H=4, T=7,
               spec=[("temporal", (1, 3), "both")], seed=14, grads=True)
    dense_case("dense_grad_denseedge", B=2, N=8, F=3, H=4, T=7, spec=[("dense",)], seed=15, grads=True)
    dense_case("dense_grad_relu", B=2, N=5, F=3, H=4, T=8, spec=[("temporal", (1,), "forward")],
               acts=("relu", "none"), seed=16, grads=True)
    # preprocessor (RayDenseGCM's configuration), window wraps
    preproc_case("dense_preproc_temporal", B=4, N=8, F_raw=6, F=8, H=8, T=20, spec=[("temporal", (1, 2), "forward")], seed=17)
    preproc_case("dense_preproc_denseedge", B=3, N=6, F_raw=5, F=4, H=8, T=14, spec=[("dense",)], seed=18, pre_act="tanh")
    # ---- sparse ------------------------------------------------------------
    sparse_case("sparse_temporal12_once", B=3, N=8, F=3, H=4, calls=[[8, 5, 7]],
                spec=[("temporal", (1, 2))], grads=True)
    sparse_case("sparse_temporal12_steps", B=3, N=8, F=3, H=4, calls=[[1, 1, 1]] * 6,
                spec=[("temporal", (1, 2))], seed=1)
    sparse_case("sparse_temporal_ragged_calls", B=4, N=16, F=5, H=6,
                calls=[[3, 1, 4, 2], [2, 5, 1, 3], [4, 4, 4, 1]], spec=[("temporal", (1, 3))],
                seed=2, grads=True)
    sparse_case("sparse_temporal_radius", B=3, N=24, F=6, H=5, calls=[[20, 24, 11]],
                spec=[("temporal", (1,))], aux=[("spatial_radius", slice(0, 2), 0.25)],
                seed=3, walk=True, grads=True)
    sparse_case("sparse_temporal_radius_steps", B=2, N=16, F=6, H=5, calls=[[4, 2], [3, 6], [5, 5]],
                spec=[("temporal", (1,))], aux=[("spatial_radius", slice(0, 2), 0.3)],
                seed=4, walk=True)
    sparse_case("sparse_temporal12_maxhops2", B=3, N=8, F=3, H=4, calls=[[2, 1, 2], [3, 3, 1], [1, 2, 2]],
                spec=[("temporal", (1, 2))], seed=5, max_hops=2)
    # subgraphs smaller than the two layers' receptive field change the numbers (k_hop_subgraph, sparse_gcm.py:182-199)
    sparse_case("sparse_temporal12_maxhops1", B=3, N=8, F=3, H=4, calls=[[2, 1, 2], [3, 3, 1], [1, 2, 2]],
                spec=[("temporal", (1, 2))], seed=6, max_hops=1)
    sparse_case("sparse_temporal12_maxhops0", B=3, N=8, F=3, H=4, calls=[[2, 1, 2], [3, 3, 1], [1, 2, 2]],
                spec=[("temporal", (1, 2))], seed=7, max_hops=0)
    sparse_case("sparse_radius_steps_maxhops1", B=2, N=16, F=6, H=5, calls=[[4, 2], [1, 1], [1, 1], [3, 6], [1, 1]],
                spec=[("temporal", (1,))], aux=[("spatial_radius", slice(0, 2), 0.3)], seed=8, walk=True, max_hops=1)
    sparse_case("sparse_radius_rollout_maxhops3", B=2, N=16, F=6, H=5, calls=[[1, 1]] * 12,
                spec=[("temporal", (1,))], aux=[("spatial_radius", slice(0, 2), 0.3)], seed=9, walk=True, max_hops=3)


def pack_case(name, B, N, max_edges, seed):
    """util.pack_hidden / util.unpack_hidden of the UNMODIFIED reference (util.py:323-382) on a random ragged COO
    adjacency with non-unit weights, one empty graph included; the oracle's restatement is checked on the way."""
    from gcm import util as ref_util

    g = torch.Generator().manual_seed(seed)
    rows = []
    for b in range(B):
        n_e = 0 if b == 1 else int(torch.randint(1, max_edges - 1, (1,), generator=g))
        pairs = torch.randperm(N * N, generator=g)[:n_e]
        rows.append(torch.stack([torch.full((n_e,), b), pairs // N, pairs % N]))
    idx = torch.cat(rows, dim=1).long()
    vals = torch.rand(idx.shape[1], generator=g) + 0.5
    adj = torch.sparse_coo_tensor(idx, vals, size=(B, N, N)).coalesce()
    nodes = torch.randn(B, N, 3, generator=g)
    T = torch.randint(0, N, (B,), generator=g)
    _, edges, weights, _ = ref_util.pack_hidden((nodes, adj, T), B, max_edges)
    _, adj2, _ = ref_util.unpack_hidden((nodes, edges, weights, T), B)
    o = oracle.pack_hidden((nodes, adj, T), B, max_edges)
    assert torch.equal(o[1], edges) and torch.equal(o[2], weights)
    o2 = oracle.unpack_hidden((nodes, edges, weights, T), B)[1]
    assert torch.equal(o2._indices(), adj2._indices()) and torch.equal(o2._values(), adj2._values())
    torch.save({"name": name, "B": B, "N": N, "max_edges": max_edges, "indices": adj.indices(), "values": adj.values(),
                "nodes": nodes, "T": T, "edges": edges, "weights": weights,
                "unpacked_indices": adj2._indices(), "unpacked_values": adj2._values()},
               os.path.join(HERE, name + ".pt"))
    print("wrote", name)


if __name__ == "__main__":
    def extras():
        pack_case("pack_ragged", B=5, N=12, max_edges=20, seed=21)
        pack_case("pack_wide", B=9, N=40, max_edges=70, seed=22)
        preproc_grad_case("train_preproc_temporal", B=4, N=12, F_raw=6, F=32, H=32, T0=15, T1=9,
                          spec=[("temporal", (1, 2, 4), "forward")], seed=31)
        preproc_grad_case("train_preproc_temporal_young", B=3, N=16, F_raw=5, F=8, H=32, T0=2, T1=7,
                          spec=[("temporal", (1, 3), "forward")], seed=32)

    if len(sys.argv) > 1 and sys.argv[1] == "extras":
        extras()
    else:
        main()
        extras()
