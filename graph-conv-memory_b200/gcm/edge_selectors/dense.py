"""DenseEdge — reference edge_selectors/dense.py:4-23: the new node is connected to and from every
earlier node, including a self loop (complete graph on the valid block)."""
import torch

from gcm import _cabi
from gcm.edge_selectors._base import FusedSelectorSpec, run_dense


class DenseEdge(torch.nn.Module):
    def __init__(self):
        super().__init__()

    def fused_spec(self):
        return FusedSelectorSpec(_cabi.SEL_DENSE)

    def forward(self, nodes, adj_mats, edge_weights, num_nodes, B):
        run_dense(self.fused_spec(), nodes, adj_mats, num_nodes)
        return adj_mats, edge_weights
