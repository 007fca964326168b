"""Distance-selected edges — reference edge_selectors/distance.py:4-81.

The current node t is connected to every EARLIER node j < t whose distance is strictly below
`max_distance` (no self edge, no unfilled slots; distance.py:27-35).  Faithful to the reference's
definitions:
  EuclideanEdge  d[b,j] = mean over ALL batch elements p of ||cur_p - nodes[b,j]||_2   (:48-49)
  CosineEdge     d[b,j] = cosine_similarity(cur_b, nodes[b,j]); edge when similarity < max_distance (:59-61)
  SpatialEdge    d[b,j] = ||cur_b[a_pose_slice] - nodes[b,j][b_pose_slice]||_2          (:77-81)
learned=True divides the nodes by a learnable scalar and thresholds at 1.0 (:13-16, :21-22).
"""
import torch

from gcm import _cabi
from gcm.edge_selectors._base import FusedSelectorSpec, run_dense


class Distance(torch.nn.Module):
    """Base class for edges based on the similarity between latent representations."""

    _kind = _cabi.SEL_NONE

    def __init__(self, max_distance, bidirectional=False, learned=False):
        super().__init__()
        self.max_distance = max_distance
        self.bidirectional = bidirectional
        self.learned = learned
        if learned:
            self.dist_param = torch.nn.Parameter(torch.Tensor([max_distance]))
            self.max_distance = 1.0
        if bidirectional:
            # unreachable through the three subclasses of the reference (distance.py:45-46,55-56,69-70)
            raise NotImplementedError("bidirectional distance edges are not part of the hot path")

    def fused_spec(self):
        return FusedSelectorSpec(
            self._kind, max_distance=self.max_distance,
            a_slice=getattr(self, "a_pose_slice", None), b_slice=getattr(self, "b_pose_slice", None),
            dist_param=self.dist_param if self.learned else None,
        )

    def forward(self, nodes, adj_mats, edge_weights, num_nodes, B):
        run_dense(self.fused_spec(), nodes, adj_mats, num_nodes)
        return adj_mats, edge_weights


class EuclideanEdge(Distance):
    _kind = _cabi.SEL_EUCLIDEAN

    def __init__(self, max_distance, learned=False):
        super().__init__(max_distance, learned=learned)


class CosineEdge(Distance):
    _kind = _cabi.SEL_COSINE

    def __init__(self, max_distance, learned=False):
        super().__init__(max_distance, learned=learned)


class SpatialEdge(Distance):
    _kind = _cabi.SEL_SPATIAL

    def __init__(self, max_distance, a_pose_slice, b_pose_slice=None, learned=False):
        super().__init__(max_distance, learned=learned)
        self.a_pose_slice = a_pose_slice
        self.b_pose_slice = b_pose_slice if b_pose_slice else a_pose_slice
