"""SpatialRadiusEdge (sparse) — reference sparse_edge_selectors/spatial.py:65-115, causal branch:
every pair (sink in the new nodes, source < sink) of the same graph whose positions
`nodes[..., position_slice]` are closer than `radius` (L2, strict).  SpatialKNNEdge (spatial.py:12-63)
needs torch_cluster's knn and is outside the hot path."""
import torch

from gcm import sparse_ops


class SpatialRadiusEdge(torch.nn.Module):
    def __init__(self, position_slice, radius=0.25, causal=True):
        super().__init__()
        self.radius = radius
        self.position_slice = position_slice
        self.causal = causal
        if not causal:
            raise NotImplementedError("only causal radius edges are part of the hot path")

    def fused_spec(self):
        return ("spatial_radius", self.position_slice, float(self.radius))

    def forward(self, nodes, T, taus, B):
        T = T.to(nodes.device).long().contiguous()
        taus = taus.to(nodes.device).long().contiguous()
        new_off = sparse_ops._excl_cumsum(taus)
        n_new, tmax = (int(v) for v in torch.stack([taus.sum(), taus.max()]).tolist())
        edges = sparse_ops.build_edges(nodes, T, taus, new_off, n_new, tmax, (),
                                       (self.position_slice, float(self.radius)))
        return torch.sparse_coo_tensor(indices=edges, values=torch.ones(edges.shape[1], device=nodes.device),
                                       size=(B, nodes.shape[1], nodes.shape[1]), is_coalesced=True)


class SpatialKNNEdge(torch.nn.Module):
    def __init__(self, position_slice, k, causal=True):
        super().__init__()
        raise NotImplementedError("SpatialKNNEdge depends on torch_cluster.knn; outside the B200 hot path")
