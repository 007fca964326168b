"""SparseGCM — drop-in for the reference's `gcm.sparse_gcm.SparseGCM`
(/root/reference/src/gcm/sparse_gcm.py:12-212).

forward(x[B,t,F], taus[B], hidden) -> (mx[B,t,H], (nodes[B,N,F], adj sparse_coo[B,N,N], T[B])), with the
adjacency indexed (batch, sink, source) like the reference.  For the hot-path configuration (two
GraphConv layers + activation, TemporalEdge / SpatialRadiusEdge selectors) one call is:

  gcm_sparse_write_flatten  node write + flat gather              (sparse_gcm.py:111-123, util.py:426-452)
  gcm_sparse_build_edges    all selectors fused, emitted coalesced (sparse_gcm.py:130-152)
  gcm_sparse_graphconv_fwd  x2: CSR gather-reduce + Linear + act   (sparse_gcm.py:178)

with no COO coalesce, no sort and no Python loop over the batch.  The second layer is evaluated only on
the rows that are returned (the new nodes).  Other configurations (preprocessor, positional encoder,
learned edges, arbitrary GNNs, max_hops < 2) take `_forward_generic` (torch ops)."""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch

from gcm import _cabi, sparse_ops, util
from gcm.fused import _ACT_OF


def _is_graph_conv(m) -> bool:
    return (type(m).__name__ == "GraphConv" and isinstance(getattr(m, "lin_rel", None), torch.nn.Linear)
            and isinstance(getattr(m, "lin_root", None), torch.nn.Linear))


def _match_sparse_gnn(gnn):
    """-> (conv1, conv2, act1, act2) for [GraphConv, act?, GraphConv, act?] stacks, else None."""
    atoms = []

    def walk(m) -> bool:
        if _is_graph_conv(m):
            if getattr(m, "aggr", "add") != "add":
                return False
            atoms.append(m)
            return True
        kids = list(m.children())
        if not kids:
            if type(m).__name__ in _ACT_OF:
                atoms.append(_ACT_OF[type(m).__name__])
                return True
            return False
        if any(f is not None for f in getattr(m, "fns", [])) or any(True for _ in m.parameters(recurse=False)):
            return False
        return all(walk(k) for k in kids)

    if any(f is not None for f in getattr(gnn, "fns", [])):
        return None
    kids = list(gnn.children())
    if not kids or not all(walk(k) for k in kids) or any(True for _ in gnn.parameters(recurse=False)):
        return None
    shape = "".join("a" if isinstance(a, str) else "c" for a in atoms)
    convs = [a for a in atoms if not isinstance(a, str)]
    acts = [a for a in atoms if isinstance(a, str)]
    if shape == "caca":
        a1, a2 = acts
    elif shape == "cca":
        a1 = a2 = acts[0]
    elif shape == "cc":
        a1 = a2 = "none"
    elif shape == "cac":
        a1, a2 = acts[0], "none"
    else:
        return None
    c1, c2 = convs
    dims = (c1.lin_rel.in_features, c1.lin_rel.out_features, c2.lin_rel.out_features)
    if c1.lin_rel.out_features != c2.lin_rel.in_features or max(dims) > 128:
        return None
    return c1, c2, a1, a2


def _one_bias(conv):
    b_rel, b_root = conv.lin_rel.bias, conv.lin_root.bias
    if b_rel is not None and b_root is not None:
        return b_rel + b_root
    return b_rel if b_rel is not None else b_root


def _all_finite(t: torch.Tensor) -> bool:
    """torch.all(torch.isfinite(t)) as ONE pass (gcm_any_nonfinite) -- the check of sparse_gcm.py:203."""
    t = t.detach()
    if not (t.is_cuda and t.dtype is torch.float32 and t.is_contiguous() and t.data_ptr() % 16 == 0):
        return bool(torch.all(torch.isfinite(t)))
    flag = torch.zeros(1, dtype=torch.int32, device=t.device)
    _cabi.check(_cabi.lib().gcm_any_nonfinite(t.data_ptr(), t.numel(), flag.data_ptr(), _cabi.stream_ptr(t.device)),
                "gcm_any_nonfinite")
    return int(flag.item()) == 0


class SparseGCM(torch.nn.Module):
    """Graph Associative Memory using sparse-graph representations"""

    did_warn = False

    def __init__(
        self,
        gnn: torch.nn.Module,
        preprocessor: torch.nn.Module = None,
        edge_selectors: torch.nn.Module = None,
        aux_edge_selectors: torch.nn.Module = None,
        graph_size: int = 128,
        max_hops: Union[int, None] = None,
        positional_encoder: torch.nn.Module = None,
    ):
        super().__init__()
        self.preprocessor = preprocessor
        self.gnn = gnn
        self.graph_size = graph_size
        self.edge_selectors = edge_selectors
        self.aux_edge_selectors = aux_edge_selectors
        self.positional_encoder = positional_encoder
        self.max_hops = max_hops
        self.ste = util.StraightThroughEstimator()
        self._plan = None
        self._plan_built = False

    def get_initial_hidden_state(self, x):
        """Zeros nodes, empty COO adjacency, T = 0 (reference sparse_gcm.py:55-70)."""
        assert x.dim() == 3
        B, _, feats = x.shape
        nodes = torch.zeros(B, self.graph_size, feats, device=x.device)
        adj = torch.zeros((B, self.graph_size, self.graph_size), device=x.device, layout=torch.sparse_coo)
        T = torch.zeros(B, dtype=torch.long, device=x.device)
        return nodes, adj, T

    # ------------------------------------------------------------------ fused plan
    def fused_plan(self):
        if not self._plan_built:
            self._plan_built = True
            self._plan = None
            if self.preprocessor is None and self.positional_encoder is None and (
                    self.max_hops is None or self.max_hops >= 0):
                g = _match_sparse_gnn(self.gnn)
                hops: List[int] = []
                radius = None
                ok = g is not None
                for sel in (self.edge_selectors, self.aux_edge_selectors):
                    if sel is None:
                        continue
                    spec = sel.fused_spec() if hasattr(sel, "fused_spec") else None
                    if spec is None:
                        ok = False
                    elif spec[0] == "temporal":
                        hops += list(spec[1])
                        ok = ok and all(h >= 1 for h in spec[1])
                    elif spec[0] == "spatial_radius" and radius is None:
                        radius = (spec[1], spec[2])
                    else:
                        ok = False
                if ok:
                    self._plan = (g, tuple(hops), radius)
        return self._plan

    def forward(self, x, taus, hidden):
        """Add tau_b observations to every graph b and query the memory for each of them."""
        plan = self.fused_plan()
        if plan is None:
            return self._forward_generic(x, taus, hidden)
        small_k = self.max_hops is not None and self.max_hops < 2
        if small_k and torch.is_grad_enabled() and (
                x.requires_grad or any(p.requires_grad for p in self.parameters())
                or (hidden is not None and hidden[0].requires_grad)):
            # a subgraph SMALLER than the two layers' receptive field changes the result (sparse_gcm.py:182-199); its
            # masked aggregation has no backward kernel
            return self._forward_generic(x, taus, hidden)
        _cabi.require_cuda(x, "SparseGCM.forward(x)")
        (c1, c2, a1, a2), hops, radius = plan
        assert x.dim() == 3 and x.dtype == torch.float32
        dev = x.device
        if hidden is None:
            hidden = self.get_initial_hidden_state(x)
        nodes, adj, T = hidden
        B, tmax, F = x.shape
        N = nodes.shape[1]
        T = T.to(dev).long().contiguous()
        taus = taus.to(dev).long().contiguous()
        counts = T + taus
        n_flat, max_count, n_new = (int(v) for v in torch.stack([counts.sum(), counts.max(), taus.sum()]).tolist())
        if max_count - 1 >= N:
            raise Exception("Overflow")
        offsets = sparse_ops._excl_cumsum(counts)
        new_off = sparse_ops._excl_cumsum(taus)

        # node write + flat gather
        if torch.is_grad_enabled() and (x.requires_grad or nodes.requires_grad):
            nodes, flat = sparse_ops._WriteFlattenFn.apply(nodes, x, T, taus, offsets, n_flat)
        else:
            # nothing to record: when every graph ends up full, the flat rows are the node tensor itself (no second copy)
            nodes, flat = sparse_ops.write_flatten_oop(nodes.detach(), x.detach().contiguous(), T, taus, offsets, n_flat,
                                                       alias_full=True)

        # edges: previous (sinks < T) + the new nodes' (sinks >= T), both sorted by (b, sink, source)
        old = adj.coalesce().indices() if adj._nnz() else torch.zeros(3, 0, dtype=torch.long, device=dev)
        csr = None
        if old.shape[1] == 0 and n_new == n_flat:
            # every row is new (hidden=None, all-at-once): the builder's per-sink offsets ARE the CSR row pointer
            # and it writes the flat column ids in the same pass
            new, edge_off, flat_col = sparse_ops.build_edges(nodes, T, taus, new_off, n_new, tmax, hops, radius, offsets)
            csr = sparse_ops.Csr(edge_off, flat_col, n_flat, node_off=offsets, max_nodes=max_count,
                                 sink_local=new[1])        # n_new == n_flat: T == 0, so sink index == row in graph
        else:
            new = sparse_ops.build_edges(nodes, T, taus, new_off, n_new, tmax, hops, radius)
        if old.shape[1] == 0:
            edges = new
        elif new.shape[1] == 0:
            edges = old
        else:
            old_cnt = torch.bincount(old[0], minlength=B)
            new_cnt = torch.bincount(new[0], minlength=B)
            old_end = torch.cumsum(old_cnt, 0)
            new_start = torch.cumsum(new_cnt, 0) - new_cnt
            edges = torch.empty(3, old.shape[1] + new.shape[1], dtype=torch.long, device=dev)
            edges[:, torch.arange(old.shape[1], device=dev) + new_start[old[0]]] = old
            edges[:, torch.arange(new.shape[1], device=dev) + old_end[new[0]]] = new
        if old.shape[1]:
            assert bool((edges[2] < edges[1]).all()), "Causality violated"
        if csr is None:
            base = offsets[edges[0]]
            csr = sparse_ops.Csr.from_sorted_edges(edges[1] + base, edges[2] + base, n_flat)

        # two GraphConv layers; the second only on the rows that are returned, the first only on the rows the second
        # reads: the returned rows and their in-neighbours (the k-hop subgraph of sparse_gcm.py:182-199 for the two
        # layers' receptive field -- max_hops >= 2 and max_hops = None give the same numbers).  In a
        # step-wise rollout (tau = 1) that is 1 + in-degree rows per graph instead of all T_b + 1.
        if n_new == n_flat:
            rows = None
        else:
            ob, ok_ = sparse_ops.ragged_arange(taus, n_new)
            rows = (offsets[ob] + T[ob] + ok_).contiguous()
        mask = None
        if small_k:
            # max_hops in {0, 1}: only edges whose source is inside the subgraph count, in BOTH layers
            keep = torch.zeros(n_flat, dtype=torch.bool, device=dev)
            if rows is None:
                keep.fill_(True)
            else:
                keep[rows] = True
                if self.max_hops == 1 and new.shape[1]:
                    keep[offsets[new[0]] + new[2]] = True
            mask = keep[csr.col].to(torch.float32).contiguous()
        if rows is None:
            h = sparse_ops.graph_conv_csr(flat, csr, None, c1.lin_rel.weight, _one_bias(c1), c1.lin_root.weight, a1,
                                          edge_mask=mask)
        else:
            rows1 = rows
            if new.shape[1] and not (small_k and self.max_hops == 0):
                rows1 = torch.cat([rows, offsets[new[0]] + new[2]])     # the new nodes' in-neighbours (duplicates possible)
            recording = torch.is_grad_enabled() and (flat.requires_grad or any(p.requires_grad for p in self.parameters()))
            if recording:
                rows1 = torch.unique(rows1)                                # autograd must see every row once
            h_sub = sparse_ops.graph_conv_csr(flat, csr, rows1.contiguous(), c1.lin_rel.weight, _one_bias(c1),
                                              c1.lin_root.weight, a1, edge_mask=mask)
            if recording:
                h = torch.zeros(n_flat, h_sub.shape[1], device=dev).index_copy(0, rows1, h_sub)
            else:
                h = (torch.zeros if small_k else torch.empty)(n_flat, h_sub.shape[1], device=dev)
                h[rows1] = h_sub                                            # rows nobody reads stay unwritten
        mx = sparse_ops.graph_conv_csr(h, csr, rows, c2.lin_rel.weight, _one_bias(c2), c2.lin_root.weight, a2,
                                       edge_mask=mask)
        assert _all_finite(mx), "Got NaN in returned memory, try using tanh activation"

        if n_new == B * tmax:
            mx_dense = mx.view(B, tmax, mx.shape[-1])       # no padding anywhere: the flat rows ARE the padded layout
        else:
            db, dk = sparse_ops.ragged_arange(taus, n_new)
            mx_dense = torch.zeros((B, tmax, mx.shape[-1]), device=dev)
            mx_dense[db, dk] = mx
        adj_out = torch.sparse_coo_tensor(indices=edges, values=torch.ones(edges.shape[1], device=dev),
                                          size=adj.shape, is_coalesced=True)
        return mx_dense, (nodes, adj_out, counts)

    # ------------------------------------------------------------------ generic (unfused) path
    def _forward_generic(self, x, taus, hidden):
        """Reference semantics (sparse_gcm.py:72-212) with torch ops, for configurations outside the
        fused hot path."""
        if hidden is None:
            hidden = self.get_initial_hidden_state(x)
        nodes, adj, T = hidden
        adj = adj.coalesce()
        N = nodes.shape[1]
        B = x.shape[0]
        b_new, k_new = util.get_new_node_idxs(T, taus, B)
        b_pad, k_pad = util.get_nonpadded_idxs(T, taus, B)
        nodes = nodes.clone()
        if int(k_new.max()) >= N:
            raise Exception("Overflow")
        nodes[b_new, k_new] = x[b_pad, k_pad]
        dirty = nodes.clone()

        def merge(adj, sel, inp):
            new = sel(inp, T, taus, B).coalesce()
            return torch.sparse_coo_tensor(indices=torch.cat([adj.indices(), new.indices()], dim=-1),
                                           values=torch.cat([adj.values(), new.values()], dim=-1),
                                           size=adj.shape).coalesce()

        if self.edge_selectors:
            adj = merge(adj, self.edge_selectors, dirty)
        if self.preprocessor:
            dirty = self.preprocessor(dirty)
        if self.positional_encoder:
            dirty = self.positional_encoder(dirty, T + taus)
        if self.aux_edge_selectors:
            adj = merge(adj, self.aux_edge_selectors, dirty)
        adj = torch.sparse_coo_tensor(indices=adj.indices(), values=adj.values() / adj.values().detach(),
                                      size=adj.shape)
        flat_nodes, out_idx = util.flatten_nodes(dirty, T, taus, B)
        edges, weights, _ = util.flatten_adj(adj, T, taus, B)
        edges = torch.flip(edges, (0,))
        assert torch.all(edges[0] < edges[1]), "Causality violated"
        if self.max_hops is None:
            mx = self.gnn(flat_nodes, edges, weights)[out_idx]
        else:
            n = flat_nodes.shape[0]
            keep = torch.zeros(n, dtype=torch.bool, device=x.device)
            keep[out_idx] = True
            frontier = keep.clone()
            for _ in range(self.max_hops):
                nxt = torch.zeros_like(keep)
                nxt[edges[0][frontier[edges[1]]]] = True
                keep |= nxt
                frontier = nxt
            sub = keep.nonzero().flatten()
            remap = torch.full((n,), -1, dtype=torch.long, device=x.device)
            remap[sub] = torch.arange(sub.numel(), device=x.device)
            em = keep[edges[0]] & keep[edges[1]]
            mx = self.gnn(flat_nodes[sub], remap[edges[:, em]], weights[em])[remap[out_idx]]
        assert torch.all(torch.isfinite(mx)), "Got NaN in returned memory, try using tanh activation"
        mx_dense = torch.zeros((*x.shape[:-1], mx.shape[-1]), device=x.device)
        mx_dense[b_pad, k_pad] = mx
        return mx_dense, (nodes, adj, T + taus)
