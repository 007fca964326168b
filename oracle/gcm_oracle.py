"""CPU oracle for the GCM memory-update-and-aggregate hot path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import this module.
The product package (`graph-conv-memory_b200/gcm`) never imports it and has no CPU
fallback.

What it is: a functional restatement, in plain CPU torch ops, of the algorithm the
reference (proroklab/graph-conv-memory v0.0.7, pure Python) runs for one
`DenseGCM.forward` step, its dense edge selectors, the 2-layer DenseGraphConv
stack, and the SparseGCM + GraphConv path.  Each function cites the reference
file:line it follows.  torch (CPU, fp32 or fp64) is used rather than numpy because
the path is floating point and the reference itself executes on ATen; autograd on
these functions is the gradient oracle for the backward kernels.

Third-party arithmetic: `torch_geometric` (pinned by the reference only as
`>= 1.7.0`, setup.cfg:22-26) is NOT vendored in /root/reference and not installed
here.  DenseGraphConv / GraphConv / coalesce / k_hop_subgraph are restated from
their published definitions (see the functions below).

Parity pinning: `tests/golden/make_golden.py` imports the UNMODIFIED reference
sources from /root/reference/src (against the stand-in under
tests/golden/standin) and stores its outputs in tests/golden/*.pt; the CPU test
suite checks this oracle against those fixtures and against the hand-written
known-answer targets of the reference's own tests (tests/test_gcm.py,
tests/test_sparse_gcm.py).  Status: PINNED for everything GCM owns (node slots,
wrap, every selector's edge set, sparse edge lists, flat indexing).  The
DenseGraphConv/GraphConv numerics for random weights are pinned only against the
stand-in's restatement of PyG (no PyG numeric golden exists in the reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------- #
# hidden state                                                                 #
# --------------------------------------------------------------------------- #
def initial_hidden(B: int, N: int, F: int, edge_weights: bool = False,
                   dtype=torch.float32):
    """gcm.py:194-211 — zeros nodes[B,N,F], adj[B,N,N], weights [0]|[B,N,N], num_nodes i64[B]."""
    nodes = torch.zeros(B, N, F, dtype=dtype)
    adj = torch.zeros(B, N, N, dtype=dtype)
    weights = torch.zeros(B, N, N, dtype=dtype) if edge_weights else torch.zeros(0, dtype=dtype)
    num_nodes = torch.zeros(B, dtype=torch.long)
    return nodes, adj, weights, num_nodes


def wrap_overflow(nodes: Tensor, adj: Tensor, weights: Tensor, num_nodes: Tensor):
    """gcm.py:323-355 — graphs with num_nodes+1 > N drop node 0 and shift down by one.

    Row/col 0 of adj (and weights) and node row 0 are zeroed, everything is rotated
    by -1 so the freed slot is the last one, and num_nodes is decremented.
    Operates on copies (the reference clones before calling, gcm.py:262-271).
    """
    N = nodes.shape[1]
    full = num_nodes + 1 > N
    if not bool(full.any()):
        return nodes, adj, weights, num_nodes
    idx = full.nonzero().flatten()
    nodes = nodes.clone()
    adj = adj.clone()
    sub_n = nodes[idx]
    sub_n[:, 0] = 0
    nodes[idx] = torch.roll(sub_n, shifts=-1, dims=1)
    sub_a = adj[idx]
    sub_a[:, 0, :] = 0
    sub_a[:, :, 0] = 0
    adj[idx] = torch.roll(sub_a, shifts=(-1, -1), dims=(1, 2))
    if weights.numel() != 0:
        weights = weights.clone()
        sub_w = weights[idx]
        sub_w[:, 0, :] = 0
        sub_w[:, :, 0] = 0
        weights[idx] = torch.roll(sub_w, shifts=(-1, -1), dims=(1, 2))
    num_nodes = torch.where(full, num_nodes - 1, num_nodes)
    return nodes, adj, weights, num_nodes


# --------------------------------------------------------------------------- #
# dense edge selectors (all OR 1s into `adj`, row = sink, column = source)     #
# --------------------------------------------------------------------------- #
def temporal_backedge(adj: Tensor, num_nodes: Tensor, hops: Sequence[int],
                      direction: str = "forward") -> Tensor:
    """edge_selectors/temporal.py:72-88 — for each hop with t >= hop:
    forward: adj[b,t,t-hop]=1; backward: adj[b,t-hop,t]=1; both: both."""
    assert direction in ("forward", "backward", "both")
    adj = adj.clone()
    for hop in hops:
        b = (num_nodes >= hop).nonzero().flatten()
        t = num_nodes[b]
        if direction in ("forward", "both"):
            adj[b, t, t - hop] = 1
        if direction in ("backward", "both"):
            adj[b, t - hop, t] = 1
    return adj


def dense_edge(adj: Tensor, num_nodes: Tensor) -> Tensor:
    """edge_selectors/dense.py:16-21 — adj[b,t,:t]=1; adj[b,:t,t]=1; adj[b,t,t]=1."""
    adj = adj.clone()
    B, N, _ = adj.shape
    j = torch.arange(N).view(1, N)
    le = j <= num_nodes.view(B, 1)                     # [B,N]: j <= t
    bidx = torch.arange(B)
    row = adj[bidx, num_nodes]                         # [B,N]
    adj[bidx, num_nodes] = torch.where(le, torch.ones_like(row), row)
    col = adj[bidx, :, num_nodes]
    adj[bidx, :, num_nodes] = torch.where(le, torch.ones_like(col), col)
    return adj


def euclidean_dist(cur: Tensor, nodes: Tensor) -> Tensor:
    """edge_selectors/distance.py:48-49 — torch.cdist(a[B,F], b[B,N,F]).mean(dim=1).

    The 2-D `a` broadcasts against the batched `b`, so the result is
    d[b,j] = mean over ALL batch elements p of ||cur_p - nodes[b,j]||_2
    (probed).  Restated with the same ATen call."""
    return torch.cdist(cur, nodes).mean(dim=1)


def cosine_dist(cur: Tensor, nodes: Tensor) -> Tensor:
    """edge_selectors/distance.py:59-61 — CosineSimilarity(dim=2, eps=1e-8) of cur_b vs nodes[b,j]."""
    a = cur.unsqueeze(1).expand_as(nodes)
    return torch.nn.functional.cosine_similarity(a, nodes, dim=2, eps=1e-8)


def spatial_dist(cur: Tensor, nodes: Tensor, a_slice: slice, b_slice: slice) -> Tensor:
    """edge_selectors/distance.py:77-81 — ||cur_b[a_slice] - nodes[b,j][b_slice]||_2.

    The reference computes a [B,N,N] cdist of N identical rows and averages them;
    the mean of N identical values is the value itself up to rounding."""
    ra = cur[:, a_slice].unsqueeze(1)                  # [B,1,k]
    rb = nodes[:, :, b_slice]                          # [B,N,k]
    return torch.cdist(ra.expand(-1, nodes.shape[1], -1), rb).mean(dim=1)


def distance_edges(nodes: Tensor, adj: Tensor, num_nodes: Tensor, kind: str,
                   max_distance: float, a_slice: Optional[slice] = None,
                   b_slice: Optional[slice] = None, dist_param: Optional[Tensor] = None,
                   return_dists: bool = False):
    """edge_selectors/distance.py:18-39 — connect t to every j < t with d[b,j] < max_distance.

    learned=True (dist_param given): nodes are divided by dist_param and the
    threshold becomes 1.0 (distance.py:13-16,21-22)."""
    B, N, _ = nodes.shape
    if dist_param is not None:
        nodes = nodes / dist_param
        max_distance = 1.0
    bidx = torch.arange(B)
    cur = nodes[bidx, num_nodes]
    if kind == "euclidean":
        d = euclidean_dist(cur, nodes)
    elif kind == "cosine":
        d = cosine_dist(cur, nodes)
    elif kind == "spatial":
        d = spatial_dist(cur, nodes, a_slice, b_slice if b_slice is not None else a_slice)
    else:
        raise ValueError(kind)
    j = torch.arange(N).view(1, N)
    sel = (d < max_distance) & (j < num_nodes.view(B, 1))
    adj = adj.clone()
    row = adj[bidx, num_nodes]
    adj[bidx, num_nodes] = torch.where(sel, torch.ones_like(row), row)
    if return_dists:
        return adj, d
    return adj


def apply_selectors(nodes: Tensor, adj: Tensor, num_nodes: Tensor, selectors) -> Tensor:
    """Chained selectors compose by OR (tests/test_gcm.py:646-658).

    `selectors`: list of tuples
      ("temporal", hops, direction) | ("dense",) | ("euclidean", max_d) |
      ("cosine", max_d) | ("spatial", max_d, a_slice, b_slice)"""
    for s in selectors or []:
        k = s[0]
        if k == "temporal":
            adj = temporal_backedge(adj, num_nodes, s[1], s[2] if len(s) > 2 else "forward")
        elif k == "dense":
            adj = dense_edge(adj, num_nodes)
        elif k in ("euclidean", "cosine"):
            adj = distance_edges(nodes, adj, num_nodes, k, s[1],
                                 dist_param=s[2] if len(s) > 2 else None)
        elif k == "spatial":
            adj = distance_edges(nodes, adj, num_nodes, k, s[1], s[2],
                                 s[3] if len(s) > 3 else None,
                                 dist_param=s[4] if len(s) > 4 else None)
        else:
            raise ValueError(k)
    return adj


# --------------------------------------------------------------------------- #
# GNN layers (published torch_geometric definitions; not in /root/reference)   #
# --------------------------------------------------------------------------- #
def _act(x: Tensor, kind: str) -> Tensor:
    if kind in (None, "none", "identity"):
        return x
    if kind == "tanh":
        return torch.tanh(x)
    if kind == "relu":
        return torch.relu(x)
    raise ValueError(kind)


def dense_graph_conv(x: Tensor, adj: Tensor, w_rel: Tensor, bias: Optional[Tensor],
                     w_root: Tensor) -> Tensor:
    """torch_geometric.nn.DenseGraphConv(aggr='add'): lin_rel(adj @ x) + lin_root(x),
    one bias vector (call sites README.md:56-62, invoked at gcm.py:308)."""
    out = torch.matmul(adj.to(x.dtype), x) @ w_rel.t() + x @ w_root.t()
    if bias is not None:
        out = out + bias
    return out


def dense_gnn2(x: Tensor, adj: Tensor, p: Dict[str, Tensor], acts=("tanh", "tanh")) -> Tensor:
    """README.md:52-62 — act(gc1(act(gc0(x, adj)), adj)) over all N rows."""
    h = _act(dense_graph_conv(x, adj, p["w_rel1"], p.get("b1"), p["w_root1"]), acts[0])
    return _act(dense_graph_conv(h, adj, p["w_rel2"], p.get("b2"), p["w_root2"]), acts[1])


def graph_conv(x: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor],
               w_rel: Tensor, bias: Optional[Tensor], w_root: Tensor) -> Tensor:
    """torch_geometric.nn.GraphConv(aggr='add'):
    out_i = lin_rel(sum_{(j->i)} w_ji x_j) + lin_root(x_i); edge_index[0]=source, [1]=sink
    (call sites ray_sparse_gcm.py:37-40; invoked at sparse_gcm.py:178,199)."""
    src, dst = edge_index[0], edge_index[1]
    msg = x[src]
    if edge_weight is not None:
        msg = msg * edge_weight.view(-1, 1)
    agg = torch.zeros_like(x[:, : x.shape[1]]).index_add(0, dst, msg)
    out = agg @ w_rel.t() + x @ w_root.t()
    if bias is not None:
        out = out + bias
    return out


def sparse_gnn2(x, edge_index, edge_weight, p, acts=("tanh", "tanh")):
    h = _act(graph_conv(x, edge_index, edge_weight, p["w_rel1"], p.get("b1"), p["w_root1"]), acts[0])
    return _act(graph_conv(h, edge_index, edge_weight, p["w_rel2"], p.get("b2"), p["w_root2"]), acts[1])


def make_params(F: int, H: int, seed: int = 7, dtype=torch.float32, H2: Optional[int] = None):
    """Default nn.Linear init for the 2-layer stack (weights seed 7)."""
    H2 = H if H2 is None else H2
    g = torch.Generator().manual_seed(seed)

    def lin(o, i):
        bound = 1.0 / math.sqrt(i)
        return (torch.rand(o, i, generator=g, dtype=torch.float64) * 2 - 1) * bound

    def bias(o, i):
        bound = 1.0 / math.sqrt(i)
        return (torch.rand(o, generator=g, dtype=torch.float64) * 2 - 1) * bound

    p = {
        "w_rel1": lin(H, F), "b1": bias(H, F), "w_root1": lin(H, F),
        "w_rel2": lin(H2, H), "b2": bias(H2, H), "w_root2": lin(H2, H),
    }
    return {k: v.to(dtype) for k, v in p.items()}


# --------------------------------------------------------------------------- #
# DenseGCM.forward                                                             #
# --------------------------------------------------------------------------- #
def dense_gcm_step(x: Tensor, hidden, selectors, params, acts=("tanh", "tanh"),
                   graph_size: int = 128, check_finite: bool = True):
    """gcm.py:213-321 — one environment step.

    hidden None -> zeros (gcm.py:241-242).  Returns (mx[B,H], (nodes, adj, weights, num_nodes+1)).
    The caller's tensors are never modified (the reference's in-place decrement of the
    caller's num_nodes on overflow, gcm.py:354, is a side effect this oracle does not copy)."""
    if hidden is None:
        hidden = initial_hidden(x.shape[0], graph_size, x.shape[1], dtype=x.dtype)
    nodes, adj, weights, num_nodes = hidden
    assert num_nodes.dtype == torch.long and num_nodes.dim() == 1
    B = x.shape[0]
    N = nodes.shape[1]
    assert N == adj.shape[1] == adj.shape[2], "N must be equal for adj mat and node mat"
    nodes, adj, weights, num_nodes = wrap_overflow(nodes, adj, weights, num_nodes)
    bidx = torch.arange(B)
    nodes = nodes.clone()
    nodes[bidx, num_nodes] = x                                   # gcm.py:274
    adj = apply_selectors(nodes, adj, num_nodes, selectors)      # gcm.py:284-287
    feats = dense_gnn2(nodes, adj, params, acts)                 # gcm.py:308
    mx = feats[bidx, num_nodes]                                  # gcm.py:314
    if check_finite:
        assert torch.all(torch.isfinite(mx)), "Got NaN in returned memory, try using tanh activation"
    return mx, (nodes, adj, weights, num_nodes + 1)


def dense_gcm_rollout(obs: Tensor, hidden, selectors, params, acts=("tanh", "tanh"),
                      graph_size: int = 128):
    """User loop of README.md:79-82 over obs[T,B,F]."""
    outs = []
    for t in range(obs.shape[0]):
        mx, hidden = dense_gcm_step(obs[t], hidden, selectors, params, acts, graph_size)
        outs.append(mx)
    return torch.stack(outs), hidden


# --------------------------------------------------------------------------- #
# SparseGCM                                                                    #
# --------------------------------------------------------------------------- #
def batch_offsets(counts: Tensor):
    """util.py:234-240 — exclusive / inclusive cumsum of per-graph node counts."""
    ends = counts.cumsum(0)
    return ends - counts, ends


def temporal_edges_sparse(T: Tensor, taus: Tensor, hops: Sequence[int]) -> Tensor:
    """sparse_edge_selectors/temporal.py:19-63 — edges (b, sink=s, source=s-h) for
    s in [T_b, T_b+tau_b), h in hops, kept when source >= 0 and sink > 0.
    Returned as int64 [3,E] ordered by (b, s, hop order)."""
    out = []
    for b in range(T.numel()):
        s = torch.arange(int(T[b]), int(T[b] + taus[b]))
        for_h = torch.tensor(list(hops), dtype=torch.long)
        sink = s.view(-1, 1).expand(-1, for_h.numel()).reshape(-1)
        src = (s.view(-1, 1) - for_h.view(1, -1)).reshape(-1)
        keep = (src >= 0) & (sink > 0)
        sink, src = sink[keep], src[keep]
        out.append(torch.stack([torch.full_like(sink, b), sink, src]))
    return torch.cat(out, dim=1) if out else torch.zeros(3, 0, dtype=torch.long)


def spatial_radius_edges_sparse(nodes: Tensor, T: Tensor, taus: Tensor, pos_slice: slice,
                                radius: float) -> Tensor:
    """sparse_edge_selectors/spatial.py:74-115 + util.py:242-263 (causal branch) —
    all pairs (sink in [T,T+tau), source < sink) with ||pos_sink - pos_source||_2 < radius;
    empty when (T+taus).max() <= 1."""
    if int((T + taus).max()) <= 1:
        return torch.zeros(3, 0, dtype=torch.long)
    out = []
    for b in range(T.numel()):
        n = int(T[b] + taus[b])
        pos = nodes[b, :n, pos_slice]
        sink, src = torch.tril_indices(n, n, offset=-1)
        keep = sink >= int(T[b])
        sink, src = sink[keep], src[keep]
        d = ((pos[sink] - pos[src]) ** 2).sum(-1).sqrt()
        hit = d < radius
        sink, src = sink[hit], src[hit]
        out.append(torch.stack([torch.full_like(sink, b), sink, src]))
    return torch.cat(out, dim=1)


def coalesce_edges(idx: Tensor) -> Tensor:
    """COO coalesce (sparse_gcm.py:107,132,139): sort by (b, sink, source), drop duplicates."""
    if idx.numel() == 0:
        return idx
    M = int(idx.max()) + 1
    key = (idx[0] * M + idx[1]) * M + idx[2]
    key = torch.unique(key, sorted=True)
    return torch.stack([key // (M * M), (key // M) % M, key % M])


def sparse_gcm_forward(x: Tensor, taus: Tensor, hidden, selectors, params,
                       acts=("tanh", "tanh"), graph_size: int = 128,
                       max_hops: Optional[int] = None, aux_selectors=None):
    """sparse_gcm.py:72-212.

    x[B,t,F] zero-padded, taus[B]; hidden None or (nodes[B,N,F], edges int64 [3,E] rows
    (b, sink, source), T[B]).  Returns (mx[B,t,H], (nodes, edges, T+taus)).
    `selectors` / `aux_selectors`: list of ("temporal", hops) | ("spatial_radius", slice, radius).
    The reference's COO adjacency is represented by its coalesced index list (values are
    forced to 1.0 at sparse_gcm.py:160-164)."""
    B, tmax, F = x.shape
    if hidden is None:
        nodes = torch.zeros(B, graph_size, F, dtype=x.dtype)
        edges = torch.zeros(3, 0, dtype=torch.long)
        T = torch.zeros(B, dtype=torch.long)
    else:
        nodes, edges, T = hidden
    N = nodes.shape[1]
    if int((T + taus).max()) - 1 >= N:                      # sparse_gcm.py:120-121
        raise Exception("Overflow")
    nodes = nodes.clone()
    for b in range(B):                                      # sparse_gcm.py:111-123
        nodes[b, int(T[b]): int(T[b] + taus[b])] = x[b, : int(taus[b])]

    def run(sel_list, edges):
        for s in sel_list or []:
            if s[0] == "temporal":
                new = temporal_edges_sparse(T, taus, s[1])
            elif s[0] == "spatial_radius":
                new = spatial_radius_edges_sparse(nodes, T, taus, s[1], s[2])
            else:
                raise ValueError(s[0])
            edges = coalesce_edges(torch.cat([edges, new], dim=1))
        return edges

    edges = run(selectors, edges)
    edges = run(aux_selectors, edges)

    counts = T + taus
    starts, _ = batch_offsets(counts)                       # util.py:234-240
    flat_nodes = torch.cat([nodes[b, : int(counts[b])] for b in range(B)])  # util.py:426-434
    out_idx = torch.cat([torch.arange(int(starts[b] + T[b]), int(starts[b] + counts[b]))
                         for b in range(B)])                # util.py:439-451
    sink = edges[1] + starts[edges[0]]                      # util.py:287-304
    src = edges[2] + starts[edges[0]]
    ei = torch.stack([src, sink])                           # flip, sparse_gcm.py:170
    assert torch.all(ei[0] < ei[1]), "Causality violated"
    w = torch.ones(ei.shape[1], dtype=x.dtype)
    if max_hops is None:
        feats = sparse_gnn2(flat_nodes, ei, w, params, acts)
        mx = feats[out_idx]
    else:                                                   # sparse_gcm.py:182-199
        keep = torch.zeros(flat_nodes.shape[0], dtype=torch.bool)
        keep[out_idx] = True
        frontier = keep.clone()
        for _ in range(max_hops):
            hit = frontier[ei[1]]
            nxt = torch.zeros_like(keep)
            nxt[ei[0][hit]] = True
            frontier = nxt
            keep |= nxt
        sub = keep.nonzero().flatten()
        remap = torch.full((flat_nodes.shape[0],), -1, dtype=torch.long)
        remap[sub] = torch.arange(sub.numel())
        em = keep[ei[0]] & keep[ei[1]]
        feats = sparse_gnn2(flat_nodes[sub], remap[ei[:, em]], w[em], params, acts)
        mx = feats[remap[out_idx]]
    assert torch.all(torch.isfinite(mx)), "Got NaN in returned memory, try using tanh activation"
    H = mx.shape[-1]
    dense = torch.zeros(B, tmax, H, dtype=x.dtype)
    pos = 0
    for b in range(B):                                      # sparse_gcm.py:205-208
        k = int(taus[b])
        dense[b, :k] = mx[pos: pos + k]
        pos += k
    return dense, (nodes, edges, T + taus)


# --------------------------------------------------------------------------- #
# RLlib state wire format (util.py:323-382)                                    #
# --------------------------------------------------------------------------- #
def pack_hidden(hidden, B: int, max_edges: int, edge_fill: int = -1, weight_fill: float = 1.0):
    """util.py:323-353 -- COO adjacency -> [B, 2, max_edges] edge list + [B, 1, max_edges] weights, one graph at a time:
    the graph's edges (rows 1 and 2 of the coalesced indices, in coalesced order) first, fill values after."""
    nodes, adj, T = hidden
    adj = adj.coalesce()
    idx, val = adj.indices(), adj.values()
    dense_edges = torch.full((B, 2, max_edges), edge_fill, dtype=torch.long)
    dense_weights = torch.full((B, 1, max_edges), weight_fill, dtype=torch.float)
    for b in range(B):
        sel = torch.nonzero(idx[0] == b).reshape(-1)
        assert sel.shape[-1] < max_edges, f"Cannot pack {sel.shape[-1]} edges into {max_edges}, increase max edges"
        for s, e in enumerate(sel.tolist()):
            dense_edges[b, 0, s] = idx[1, e]
            dense_edges[b, 1, s] = idx[2, e]
            dense_weights[b, 0, s] = val[e]
    return nodes, dense_edges, dense_weights, T


def unpack_hidden(hidden, B: int):
    """util.py:355-382 -- the slots whose first row is >= 0, in (graph, slot) order, back to a COO adjacency."""
    nodes, edges, weights, T = hidden
    bs, i1, i2, vs = [], [], [], []
    for b in range(edges.shape[0]):
        for s in range(edges.shape[2]):
            if int(edges[b, 0, s]) >= 0:
                bs.append(b); i1.append(int(edges[b, 0, s])); i2.append(int(edges[b, 1, s])); vs.append(float(weights[b, 0, s]))
    idx = torch.tensor([bs, i1, i2], dtype=torch.long).reshape(3, -1)
    adj = torch.sparse_coo_tensor(indices=idx, values=torch.tensor(vs, dtype=torch.float),
                                  size=(B, nodes.shape[1], nodes.shape[1]))
    return nodes, adj, T
