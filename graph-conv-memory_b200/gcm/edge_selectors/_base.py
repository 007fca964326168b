"""Shared plumbing of the dense edge selectors.

Every selector keeps the reference's module interface
`forward(nodes, adj_mats, edge_weights, num_nodes, B) -> (adj_mats, edge_weights)`
(edge_selectors/temporal.py:90-94, dense.py:11-23, distance.py:18-39 of the reference) and
additionally describes itself through `fused_spec()`, which `DenseGCM` uses to run the
selector inside the fused step kernel on the bit-packed adjacency.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from gcm import _cabi


class FusedSelectorSpec:
    """Host-side mirror of `gcm_selector` (include/gcm_b200.h)."""

    __slots__ = ("kind", "direction", "hops", "max_distance", "a_slice", "b_slice", "dist_param")

    def __init__(self, kind, direction=0, hops=(), max_distance=0.0, a_slice=None, b_slice=None,
                 dist_param=None):
        self.kind = kind
        self.direction = direction
        self.hops = tuple(int(h) for h in hops)
        self.max_distance = float(max_distance)
        self.a_slice = a_slice
        self.b_slice = b_slice
        self.dist_param = dist_param

    def key(self):
        return (self.kind, self.direction, self.hops)

    def to_c(self, F: int, dist: Optional[torch.Tensor] = None) -> _cabi.SelectorC:
        s = _cabi.SelectorC()
        s.kind = self.kind
        s.direction = self.direction
        if len(self.hops) > _cabi.GCM_MAX_HOPS:
            raise _cabi.GcmLibraryError(f"at most {_cabi.GCM_MAX_HOPS} hops per selector")
        s.n_hops = len(self.hops)
        for i, h in enumerate(self.hops):
            s.hops[i] = h
        s.max_distance = self.max_distance
        s.a_step = s.b_step = 1
        if self.kind == _cabi.SEL_SPATIAL:
            a0, a1, a_st = self.a_slice.indices(F)
            b0, b1, b_st = (self.b_slice if self.b_slice is not None else self.a_slice).indices(F)
            if a_st < 1 or b_st < 1:
                raise _cabi.GcmLibraryError("SpatialEdge: negative slice steps are not supported")
            la = len(range(a0, a1, a_st))
            lb = len(range(b0, b1, b_st))
            if la != lb:
                raise RuntimeError(
                    f"SpatialEdge: pose slices select {la} and {lb} features; they must match"
                )
            s.a_start, s.a_step, s.b_start, s.b_step, s.slice_len = a0, a_st, b0, b_st, la
        s.dist_param = self.dist_param.data_ptr() if self.dist_param is not None else None
        s.dist = dist.data_ptr() if dist is not None else None
        return s


def run_dense(spec: FusedSelectorSpec, nodes: torch.Tensor, adj_mats: torch.Tensor,
              num_nodes: torch.Tensor) -> None:
    """ORs the selector's edges into the dense `adj_mats` [B,N,N] in place (gcm_select_dense)."""
    _cabi.require_cuda(adj_mats, type(spec).__name__)
    if adj_mats.dtype != torch.float32 or not adj_mats.is_contiguous():
        raise _cabi.GcmLibraryError("edge selectors expect a contiguous float32 adj_mats")
    B, N, _ = adj_mats.shape
    nodes_c = nodes.detach().contiguous().float()
    F = nodes_c.shape[-1]
    nn = num_nodes.to(device=adj_mats.device, dtype=torch.long).contiguous()
    lib = _cabi.lib()
    dist = None
    if spec.kind == _cabi.SEL_EUCLIDEAN:
        bidx = torch.arange(B, device=nodes_c.device)
        cur = nodes_c[bidx, nn.clamp(max=N - 1)].contiguous()
        dist = torch.empty(B, N, device=nodes_c.device, dtype=torch.float32)
        st = _cabi.DenseStateC(nodes_c.data_ptr(), nodes_c.data_ptr(), nodes_c.data_ptr(), B, N, N, F,
                               (N + 31) // 32)
        _cabi.check(lib.gcm_euclid_batchmean(C.byref(st), cur.data_ptr(), B,
                                             _cabi.ptr(spec.dist_param), dist.data_ptr(),
                                             _cabi.stream_ptr(nodes_c.device)), "gcm_euclid_batchmean")
    sel = spec.to_c(F, dist)
    _cabi.check(lib.gcm_select_dense(nodes_c.data_ptr(), adj_mats.data_ptr(), nn.data_ptr(), B, N, F,
                                     C.byref(sel), _cabi.stream_ptr(adj_mats.device)), "gcm_select_dense")
