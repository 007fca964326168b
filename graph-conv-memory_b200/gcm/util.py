"""Index helpers of the GCM package, same names and results as the reference's gcm/util.py.

The reference builds these index tensors with Python loops over the batch and `torch.cat`
(util.py:176-231, 426-452); here they are closed-form vectorised tensor expressions (cumsum /
repeat_interleave / arange arithmetic) that run on the tensors' device without host loops.
The fused kernels do this index math on the device themselves; these helpers remain for
callers of the public API (RaySparseGCM uses pack_hidden / unpack_hidden, ray_sparse_gcm.py:195-213).
The learned-edge utilities of the reference (Spardmax, sparse gumbel softmax, util.py:29-172)
are outside the hot path and are not provided.
"""
from __future__ import annotations

from typing import List, Tuple

import torch


class STEFunction(torch.autograd.Function):
    """Straight-through estimator: forward (x > 0), backward identity (reference util.py:9-17)."""

    @staticmethod
    def forward(ctx, input):
        return (input > 0).float()

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output


class StraightThroughEstimator(torch.nn.Module):
    def forward(self, x):
        return STEFunction.apply(x)


class Hardmax(torch.nn.Module):
    """Hard softmax with straight-through gradient (reference util.py:45-56)."""

    def __init__(self, dim=-1, cutoff=0.2):
        super().__init__()
        self.dim, self.cutoff = dim, cutoff

    def forward(self, x):
        y_soft = torch.softmax(x, self.dim)
        y_hard = (y_soft > self.cutoff).float()
        return y_hard - y_soft.detach() + y_soft


def _ragged_arange(lengths: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(owner index, position within owner) for every element of a ragged batch."""
    lengths = lengths.long()
    owners = torch.repeat_interleave(torch.arange(lengths.numel(), device=lengths.device), lengths)
    starts = torch.cumsum(lengths, 0) - lengths
    pos = torch.arange(int(lengths.sum()), device=lengths.device) - starts[owners]
    return owners, pos


def get_nonpadded_idxs(T: torch.Tensor, taus: torch.Tensor, B: int):
    """(batch, time) indices of the valid entries of a zero-padded [B, t, F] observation batch
    (reference util.py:176-188)."""
    return _ragged_arange(taus)


def get_new_node_idxs(T: torch.Tensor, taus: torch.Tensor, B: int):
    """(batch, node) indices of the nodes added this call: nodes T[b] .. T[b]+tau[b]-1
    (reference util.py:191-208)."""
    b, k = _ragged_arange(taus)
    return b, k + T.long()[b]


def get_valid_node_idxs(T: torch.Tensor, taus: torch.Tensor, B: int):
    """(batch, node) indices of every valid node 0 .. T[b]+tau[b]-1 (reference util.py:211-231)."""
    return _ragged_arange(T + taus)


def get_batch_offsets(T: torch.Tensor):
    """Exclusive / inclusive cumulative node counts (reference util.py:234-240)."""
    batch_ends = T.cumsum(dim=0)
    return batch_ends - T, batch_ends


def get_causal_edges_one_batch(t, tau, window=None):
    """All (sink, source) pairs with source < sink and sink among the tau new nodes
    (reference util.py:242-263)."""
    dev = t.device if isinstance(t, torch.Tensor) else None
    t, tau = int(t), int(tau)
    sink_len = torch.arange(t, t + tau, device=dev)            # a sink s has s candidate sources
    sink, src = _ragged_arange(sink_len)
    sink = sink + t
    if window is not None:
        keep = src >= max(0, t - int(window))
        sink, src = sink[keep], src[keep]
    return torch.stack([sink, src])


def get_causal_edges(T, taus, window=None):
    """Batched causal pairs as rows (batch, sink, source) (reference util.py:270-282)."""
    out = []
    for b in range(T.numel()):
        e = get_causal_edges_one_batch(T[b], taus[b], window=window).to(T.device)
        out.append(torch.cat([torch.full((1, e.shape[1]), b, dtype=torch.long, device=T.device), e]))
    return torch.cat(out, dim=-1)


def flatten_adj(adj, T, taus, B):
    """COO [B, MAX, MAX] -> flat edges [2, E] offset per graph, weights, batch ids
    (reference util.py:287-304)."""
    batch_starts, _ = get_batch_offsets(T + taus)
    adj = adj.coalesce()
    batch_idx = adj.indices()[0]
    return adj.indices()[1:] + batch_starts[batch_idx], adj.values(), batch_idx


def unflatten_adj(edges, weights, batch_idx, T, taus, B, max_edges):
    """Inverse of flatten_adj (reference util.py:307-319)."""
    batch_starts, _ = get_batch_offsets(T + taus)
    local = edges - batch_starts[batch_idx]
    return torch.sparse_coo_tensor(indices=torch.stack([batch_idx, local[0], local[1]]), values=weights,
                                   size=(B, max_edges, max_edges))


def pack_hidden(hidden, B, max_edges: int, edge_fill: int = -1, weight_fill: float = 1.0):
    """COO adjacency -> fixed-size [B, 2, max_edges] edge list for RLlib (reference util.py:323-353).  CUDA tensors: one
    kernel (gcm_pack_edges, csrc/gcm_pack.cu); CPU tensors: the same thing with torch indexing (helper API, not the hot
    path)."""
    nodes, adj, T = hidden
    adj = adj.coalesce()
    idx, val = adj.indices(), adj.values()
    if adj.is_cuda:
        from gcm import _cabi
        dense_edges = torch.empty((B, 2, max_edges), device=adj.device, dtype=torch.long)
        dense_weights = torch.empty((B, 1, max_edges), device=adj.device, dtype=torch.float)
        counts = torch.empty(B, device=adj.device, dtype=torch.int32)
        idx_c, val_c = idx.contiguous(), val.to(torch.float32).contiguous()
        _cabi.check(_cabi.lib().gcm_pack_edges(idx_c.data_ptr(), val_c.data_ptr(), idx_c.shape[1], B, max_edges,
                                               int(edge_fill), float(weight_fill), dense_edges.data_ptr(),
                                               dense_weights.data_ptr(), counts.data_ptr(),
                                               _cabi.stream_ptr(adj.device)), "gcm_pack_edges")
        most = int(counts.max()) if B else 0
        assert most < max_edges or idx_c.shape[1] == 0, f"Cannot pack {most} edges into {max_edges}, increase max edges"
        return nodes, dense_edges, dense_weights, T
    dense_edges = torch.full((B, 2, max_edges), edge_fill, device=adj.device, dtype=torch.long)
    dense_weights = torch.full((B, 1, max_edges), weight_fill, device=adj.device, dtype=torch.float)
    counts = torch.bincount(idx[0], minlength=B)
    assert int(counts.max()) < max_edges if idx.numel() else True, (
        f"Cannot pack {int(counts.max()) if idx.numel() else 0} edges into {max_edges}, increase max edges")
    starts = torch.cumsum(counts, 0) - counts
    slot = torch.arange(idx.shape[1], device=adj.device) - starts[idx[0]]   # coalesced => sorted by batch
    dense_edges[idx[0], 0, slot] = idx[1]
    dense_edges[idx[0], 1, slot] = idx[2]
    dense_weights[idx[0], 0, slot] = val
    return nodes, dense_edges, dense_weights, T


def unpack_hidden(hidden, B):
    """Fixed-size edge list -> COO adjacency (reference util.py:355-382).  CUDA tensors: gcm_count_valid_edges +
    gcm_unpack_edges (one host read of the edge total, as the reference's nonzero() has)."""
    nodes, edges, weights, T = hidden
    if edges.is_cuda:
        from gcm import _cabi
        lib = _cabi.lib()
        stream = _cabi.stream_ptr(edges.device)
        Bn, _, M = edges.shape
        e_c = edges.to(torch.long).contiguous()
        w_c = weights.to(torch.float32).contiguous()
        offsets = torch.zeros(Bn + 1, device=edges.device, dtype=torch.long)
        _cabi.check(lib.gcm_count_valid_edges(e_c.data_ptr(), Bn, M, offsets[1:].data_ptr(), stream),
                    "gcm_count_valid_edges")
        torch.cumsum(offsets[1:], 0, out=offsets[1:])
        E = int(offsets[-1])
        coo = torch.empty(3, E, device=edges.device, dtype=torch.long)
        vals = torch.empty(E, device=edges.device, dtype=torch.float32)
        _cabi.check(lib.gcm_unpack_edges(e_c.data_ptr(), w_c.data_ptr(), Bn, M, offsets.data_ptr(), E, coo.data_ptr(),
                                         vals.data_ptr(), stream), "gcm_unpack_edges")
        adj = torch.sparse_coo_tensor(indices=coo, values=vals, size=(B, nodes.shape[1], nodes.shape[1]))
        return nodes, adj, T
    batch_idx, edge_idx = (edges[:, 0] >= 0).nonzero().T.unbind()
    adj_idx = torch.stack([batch_idx, edges[batch_idx, 0, edge_idx], edges[batch_idx, 1, edge_idx]])
    adj = torch.sparse_coo_tensor(indices=adj_idx, values=weights[batch_idx, 0, edge_idx],
                                  size=(B, nodes.shape[1], nodes.shape[1]))
    return nodes, adj, T


def flatten_nodes(nodes: torch.Tensor, T: torch.Tensor, taus: torch.Tensor, B: int):
    """Valid rows of [B, N, F] -> [sum(T+tau), F] plus the flat ids of the new nodes
    (reference util.py:426-452)."""
    batch_offsets, _ = get_batch_offsets(T + taus)
    b, k = get_valid_node_idxs(T, taus, B)
    ob, ok = _ragged_arange(taus)
    return nodes[b, k], batch_offsets[ob] + T.long()[ob] + ok


def diff_or(tensors: List[torch.Tensor]):
    """Differentiable OR of {0,1} tensors (reference util.py:455-465)."""
    res = torch.zeros_like(tensors[0])
    for t in tensors:
        res = res + t - res * t
    return res


def idxs_up_to_including_num_nodes(nodes: torch.Tensor, num_nodes: torch.Tensor):
    """(batch, node) indices with node <= num_nodes[batch] (reference util.py:478-498)."""
    N = nodes.shape[1]
    hit = torch.arange(N, device=nodes.device).unsqueeze(0) <= num_nodes.unsqueeze(1)
    bn = torch.nonzero(hit)
    return bn[:, 0], bn[:, 1]


def idxs_up_to_num_nodes(adj: torch.Tensor, num_nodes: torch.Tensor):
    """(batch, past node, current node) with past < num_nodes[batch] (reference util.py:501-522)."""
    N = adj.shape[-1]
    hit = torch.arange(N, device=adj.device).unsqueeze(0) < num_nodes.unsqueeze(1)
    bn = torch.nonzero(hit)
    return bn[:, 0], bn[:, 1], num_nodes[bn[:, 0]]
