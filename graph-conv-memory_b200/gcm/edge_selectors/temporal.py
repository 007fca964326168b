"""TemporalBackedge — same constructor and call interface as the reference's
edge_selectors/temporal.py:17-94 (deterministic branch, :72-88).

adj[i, j] = 1 means information flows j -> i (row = sink, column = source)."""
from typing import List

import torch

from gcm import _cabi
from gcm.edge_selectors._base import FusedSelectorSpec, run_dense


class TemporalBackedge(torch.nn.Module):
    """Add temporal directional back edges, e.g. node_t <- node_{t-1} for hops=[1]."""

    def __init__(self, hops: List[int] = [1], direction="forward", learned=False, learning_window=10,
                 deterministic=False, num_samples=3):
        super().__init__()
        self.hops = list(hops)
        assert direction in ["forward", "backward", "both"]
        self.direction = direction
        self.learned = learned
        if learned:
            # reference temporal.py:51-70: gumbel / sparsemax window over past offsets.  Research-only
            # branch (depends on the un-imported `sparsemax`, util.py:5,36); outside the hot path.
            raise NotImplementedError(
                "TemporalBackedge(learned=True) is outside the B200 hot path (SURVEY.md §2 row 2)"
            )

    def fused_spec(self):
        return FusedSelectorSpec(_cabi.SEL_TEMPORAL, _cabi.DIR[self.direction], self.hops)

    def forward(self, nodes, adj_mats, edge_weights, num_nodes, B):
        run_dense(self.fused_spec(), nodes, adj_mats, num_nodes)
        return adj_mats, edge_weights
