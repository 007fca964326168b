"""TemporalEdge (sparse) — reference sparse_edge_selectors/temporal.py:11-63.

For every new node s in [T_b, T_b + tau_b) and hop h: edge (batch b, sink s, source s - h), kept when
source >= 0 and sink > 0.  Returns a torch sparse COO adjacency with index rows (batch, sink, source)
and unit values, like the reference (which uses a nominal size of (B, 1e5, 1e5))."""
from typing import List

import torch

from gcm import sparse_ops


class TemporalEdge(torch.nn.Module):
    """Add temporal edges to the edge list"""

    def __init__(self, hops: List[int] = [1]):
        super().__init__()
        self.hops = torch.tensor(hops)

    def fused_spec(self):
        return ("temporal", tuple(int(h) for h in self.hops.tolist()))

    def forward(self, nodes, T, taus, B):
        T = T.to(nodes.device).long().contiguous()
        taus = taus.to(nodes.device).long().contiguous()
        new_off = sparse_ops._excl_cumsum(taus)
        n_new, tmax = (int(v) for v in torch.stack([taus.sum(), taus.max()]).tolist())
        edges = sparse_ops.build_edges(nodes, T, taus, new_off, n_new, tmax, self.fused_spec()[1], None)
        return torch.sparse_coo_tensor(indices=edges, values=torch.ones(edges.shape[1], device=nodes.device),
                                       size=(B, int(1e5), int(1e5)), is_coalesced=True)
