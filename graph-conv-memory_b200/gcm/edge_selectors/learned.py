"""`gcm.edge_selectors.learned.LearnedEdge` -- import path of the reference's learned dense edge selector
(/root/reference/src/gcm/edge_selectors/learned.py:7-125: an MLP over (current, past) node pairs, gumbel-softmax /
straight-through sampling, float adjacency with gradients).  SURVEY.md section 2 row 13 / section 8: stochastic,
non-binary adjacency, built on util.Spardmax whose `sparsemax` import is commented out in the reference -- OUT OF SCOPE of
the B200 hot path.  The name resolves so that `from gcm.edge_selectors.learned import LearnedEdge` in caller code keeps
importing; constructing it says what to use instead.  Any user-written selector module with the reference's call
signature `(nodes, adj_mats, edge_weights, num_nodes, B) -> (adj_mats, edge_weights)` still runs through
DenseGCM._forward_generic (a float / learned adjacency is never bit-packed)."""
import torch


class LearnedEdge(torch.nn.Module):
    def __init__(self, input_size: int = 0, model: torch.nn.Sequential = None, num_edge_samples: int = 5,
                 deterministic: bool = False):
        super().__init__()
        raise NotImplementedError(
            "gcm.edge_selectors.learned.LearnedEdge (learned, stochastic edge priors) is outside the B200 hot path; "
            "use TemporalBackedge / DenseEdge / EuclideanEdge / CosineEdge / SpatialEdge, or pass your own selector "
            "module (it runs through DenseGCM's generic path)")
