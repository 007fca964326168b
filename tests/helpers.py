"""Shared test helpers: build the product modules / oracle arguments from a golden fixture."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ACT = {"tanh": torch.nn.Tanh, "relu": torch.nn.ReLU, "none": torch.nn.Identity}


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def dense_cases():
    return sorted(f[:-3] for f in os.listdir(GOLDEN) if f.startswith("dense_") and f.endswith(".pt")
                  and not f.startswith("dense_preproc"))


def preproc_cases():
    """fixtures of the reference DenseGCM WITH a preprocessor (tests/golden/make_golden.py: preproc_case)"""
    return sorted(f[:-3] for f in os.listdir(GOLDEN) if f.startswith("dense_preproc") and f.endswith(".pt"))


def make_preprocessor(g):
    lin = torch.nn.Linear(g["F_raw"], g["F"])
    with torch.no_grad():
        lin.weight.copy_(g["pre_weight"])
        lin.bias.copy_(g["pre_bias"])
    return lin if g["pre_act"] is None else torch.nn.Sequential(lin, ACT[g["pre_act"]]())


def sparse_cases():
    return sorted(f[:-3] for f in os.listdir(GOLDEN) if f.startswith("sparse_") and f.endswith(".pt"))


def make_dense_gnn(F, H, params, acts, style="readme"):
    """The README's user GNN built from gcm.nn layers (README.md:52-62 of the reference)."""
    from gcm.nn import DenseGraphConv, Sequential

    if style == "sequential":
        g = Sequential("x, adj, weights, B, N", [
            (DenseGraphConv(F, H), "x, adj -> x"), ACT[acts[0]](),
            (DenseGraphConv(H, H), "x, adj -> x"), ACT[acts[1]](),
        ])
        convs = [m for m in g.modules() if type(m).__name__ == "DenseGraphConv"]
    else:
        class GNN(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.gc0 = DenseGraphConv(F, H)
                self.a0 = ACT[acts[0]]()
                self.gc1 = DenseGraphConv(H, H)
                self.a1 = ACT[acts[1]]()

            def forward(self, x, adj, weights, B, N):
                x = self.a0(self.gc0(x, adj))
                return self.a1(self.gc1(x, adj))

        g = GNN()
        convs = [g.gc0, g.gc1]
    with torch.no_grad():
        convs[0].lin_rel.weight.copy_(params["w_rel1"]); convs[0].lin_rel.bias.copy_(params["b1"])
        convs[0].lin_root.weight.copy_(params["w_root1"])
        convs[1].lin_rel.weight.copy_(params["w_rel2"]); convs[1].lin_rel.bias.copy_(params["b2"])
        convs[1].lin_root.weight.copy_(params["w_root2"])
    return g, convs


def make_selector(spec):
    from gcm.edge_selectors.dense import DenseEdge
    from gcm.edge_selectors.distance import CosineEdge, EuclideanEdge, SpatialEdge
    from gcm.edge_selectors.temporal import TemporalBackedge
    from gcm.nn import Sequential

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
    return Sequential("x, adj, weights, num_nodes, B",
                      [(m, "x, adj, weights, num_nodes, B -> adj, weights") for m in mods])


def named_grads(convs):
    return {"w_rel1": convs[0].lin_rel.weight.grad, "b1": convs[0].lin_rel.bias.grad,
            "w_root1": convs[0].lin_root.weight.grad, "w_rel2": convs[1].lin_rel.weight.grad,
            "b2": convs[1].lin_rel.bias.grad, "w_root2": convs[1].lin_root.weight.grad}


def rel_err(a, b):
    """max |a-b| / max(1, max|b|): the parity tolerances of BASELINE.json are relative."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(1.0, float(b.abs().max())))


def make_sparse_gnn(F, H, params, acts, style="readme"):
    from gcm.nn import GraphConv, Sequential

    if style == "sequential":
        g = Sequential("x, edges, weights", [
            (GraphConv(F, H), "x, edges, weights -> x"), ACT[acts[0]](),
            (GraphConv(H, H), "x, edges, weights -> x"), ACT[acts[1]](),
        ])
        convs = [m for m in g.modules() if type(m).__name__ == "GraphConv"]
    else:
        class GNN(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.gc0 = GraphConv(F, H)
                self.a0 = ACT[acts[0]]()
                self.gc1 = GraphConv(H, H)
                self.a1 = ACT[acts[1]]()

            def forward(self, x, edges, weights):
                x = self.a0(self.gc0(x, edges, weights))
                return self.a1(self.gc1(x, edges, weights))

        g = GNN()
        convs = [g.gc0, g.gc1]
    with torch.no_grad():
        convs[0].lin_rel.weight.copy_(params["w_rel1"]); convs[0].lin_rel.bias.copy_(params["b1"])
        convs[0].lin_root.weight.copy_(params["w_root1"])
        convs[1].lin_rel.weight.copy_(params["w_rel2"]); convs[1].lin_rel.bias.copy_(params["b2"])
        convs[1].lin_root.weight.copy_(params["w_root2"])
    return g, convs


def make_sparse_selector(spec):
    from gcm.sparse_edge_selectors.spatial import SpatialRadiusEdge
    from gcm.sparse_edge_selectors.temporal import TemporalEdge

    if not spec:
        return None
    assert len(spec) == 1
    s = spec[0]
    if s[0] == "temporal":
        return TemporalEdge(list(s[1]))
    return SpatialRadiusEdge(s[1], s[2])
