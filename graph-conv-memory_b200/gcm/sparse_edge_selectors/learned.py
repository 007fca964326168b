"""`gcm.sparse_edge_selectors.learned.LearnedEdge` -- import path of the reference's learned sparse edge selector
(/root/reference/src/gcm/sparse_edge_selectors/learned.py:12-160: an MLP over the causal edge list, sparse gumbel
softmax with a learned temperature, `stats` dict and gradient-norm hooks).  SURVEY.md section 2 row 14: stochastic,
research-only -- OUT OF SCOPE of the B200 hot path.  The name resolves; constructing it raises.  A user-written sparse
selector with the reference's signature `(nodes, T, taus, B) -> sparse_coo` still runs through SparseGCM's generic
path."""
from typing import Tuple, Union

import torch


class LearnedEdge(torch.nn.Module):
    def __init__(self, input_size: int = 0, model: Union[None, torch.nn.Module] = None, num_edge_samples: int = 5,
                 deterministic: bool = False, window: Union[int, None] = None, log_stats: bool = True,
                 softmax_temp: float = 1.0, learn_softmax_temp: bool = True,
                 temp_bounds: Tuple[float, float] = (0.001, 5), store_grads: bool = True):
        super().__init__()
        raise NotImplementedError(
            "gcm.sparse_edge_selectors.learned.LearnedEdge (learned, stochastic edge priors) is outside the B200 hot "
            "path; use TemporalEdge / SpatialRadiusEdge, or pass your own selector module (generic path)")
