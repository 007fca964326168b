"""GNN layers with torch_geometric's interface and parameter names.

The reference's GNNs are user modules built from `torch_geometric.nn.DenseGraphConv` /
`GraphConv` (call sites /root/reference/README.md:52-62, src/gcm/ray_sparse_gcm.py:34-42).
torch_geometric is a third-party dependency that is not part of this image, so the same
layers are provided here with the same constructor arguments, forward signatures and
state_dict keys (`lin_rel.weight`, `lin_rel.bias`, `lin_root.weight`; PyG 1.x kept the
bias on `lin_root` instead -- pass `bias_on="root"`).  `DenseGCM` / `SparseGCM` recognise
either these classes or the real PyG ones (duck-typed on `lin_rel` / `lin_root`) and run
them through the fused CUDA kernels; called directly they evaluate the published
definition with torch ops (used for the one-time numerical validation of the fused plan
and by arbitrary user GNNs on the generic path).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple, Union

import torch


class DenseGraphConv(torch.nn.Module):
    """out = lin_rel(adj @ x) + lin_root(x), aggr='add' (published PyG definition)."""

    def __init__(self, in_channels: int, out_channels: int, aggr: str = "add", bias: bool = True,
                 bias_on: str = "rel"):
        super().__init__()
        assert aggr == "add", "only aggr='add' (the reference's configuration) is supported"
        assert bias_on in ("rel", "root")
        self.in_channels, self.out_channels, self.aggr = in_channels, out_channels, aggr
        self.lin_rel = torch.nn.Linear(in_channels, out_channels, bias=bias and bias_on == "rel")
        self.lin_root = torch.nn.Linear(in_channels, out_channels, bias=bias and bias_on == "root")

    def reset_parameters(self):
        self.lin_rel.reset_parameters()
        self.lin_root.reset_parameters()

    def forward(self, x: torch.Tensor, adj: torch.Tensor, mask: Optional[torch.Tensor] = None):
        x = x.unsqueeze(0) if x.dim() == 2 else x
        adj = adj.unsqueeze(0) if adj.dim() == 2 else adj
        out = self.lin_rel(torch.matmul(adj.to(x.dtype), x)) + self.lin_root(x)
        if mask is not None:
            out = out * mask.view(x.shape[0], x.shape[1], 1).to(x.dtype)
        return out


class GraphConv(torch.nn.Module):
    """out_i = lin_rel(sum_{j->i} w_ji x_j) + lin_root(x_i); edge_index[0]=source, [1]=sink."""

    def __init__(self, in_channels: int, out_channels: int, aggr: str = "add", bias: bool = True,
                 bias_on: str = "rel"):
        super().__init__()
        assert aggr == "add"
        assert bias_on in ("rel", "root")
        self.in_channels, self.out_channels, self.aggr = in_channels, out_channels, aggr
        self.lin_rel = torch.nn.Linear(in_channels, out_channels, bias=bias and bias_on == "rel")
        self.lin_root = torch.nn.Linear(in_channels, out_channels, bias=bias and bias_on == "root")

    def reset_parameters(self):
        self.lin_rel.reset_parameters()
        self.lin_root.reset_parameters()

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor,
                edge_weight: Optional[torch.Tensor] = None):
        if x.is_cuda:
            from gcm import sparse_ops  # CSR segmented gather-reduce kernels

            return sparse_ops.graph_conv(x, edge_index, edge_weight, self.lin_rel.weight,
                                         _one_bias(self), self.lin_root.weight, act="none")
        src, dst = edge_index[0], edge_index[1]
        msg = x[src]
        if edge_weight is not None and edge_weight.numel() > 0:
            msg = msg * edge_weight.view(-1, 1)
        agg = torch.zeros_like(x).index_add(0, dst, msg)
        return self.lin_rel(agg) + self.lin_root(x)


def _one_bias(conv) -> Optional[torch.Tensor]:
    b_rel = getattr(conv.lin_rel, "bias", None)
    b_root = getattr(conv.lin_root, "bias", None)
    if b_rel is not None and b_root is not None:
        return b_rel + b_root
    return b_rel if b_rel is not None else b_root


class Sequential(torch.nn.Module):
    """`torch_geometric.nn.Sequential(input_args, [(module, "a, b -> c"), module, ...])`.

    A bare module consumes the previous step's output and rebinds that output's name."""

    def __init__(self, input_args: str, modules: Sequence[Union[Tuple[Callable, str], Callable]]):
        super().__init__()
        self.input_args = [a.strip() for a in input_args.split(",") if a.strip()]
        self.specs: List[Tuple[Optional[List[str]], Optional[List[str]]]] = []
        self.fns: List[Optional[Callable]] = []
        mods = torch.nn.ModuleList()
        for m in modules:
            if isinstance(m, (tuple, list)):
                fn, desc = m
                lhs, rhs = desc.split("->")
                ins = [a.strip() for a in lhs.split(",") if a.strip()]
                outs = [a.strip() for a in rhs.split(",") if a.strip()]
            else:
                fn, ins, outs = m, None, None
            self.specs.append((ins, outs))
            if isinstance(fn, torch.nn.Module):
                mods.append(fn)
                self.fns.append(None)
            else:
                mods.append(torch.nn.Identity())
                self.fns.append(fn)
        self.mods = mods

    def steps(self):
        for (ins, outs), mod, fn in zip(self.specs, self.mods, self.fns):
            yield ins, outs, (fn if fn is not None else mod)

    def forward(self, *args):
        env = dict(zip(self.input_args, args))
        last, last_outs = None, None
        for ins, outs, f in self.steps():
            if ins is None:
                res = f(*last) if isinstance(last, tuple) else f(last)
                outs_eff = last_outs
            else:
                res = f(*[env[k] for k in ins])
                outs_eff = outs
                last_outs = outs
            last = res
            if outs_eff:
                if isinstance(res, tuple):
                    for k, v in zip(outs_eff, res):
                        env[k] = v
                else:
                    env[outs_eff[0]] = res
        return last
