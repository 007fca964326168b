"""Stand-in for torch_geometric.nn — restates the published PyG layer definitions.

TEST INFRASTRUCTURE ONLY (see package docstring).  Definitions follow the public
PyG documentation:
  DenseGraphConv(aggr='add'): out = lin_rel(adj @ x) + lin_root(x)
  GraphConv(aggr='add'):      out_i = lin_rel(sum_{j->i} w_ji x_j) + lin_root(x_i)
Both carry exactly one bias vector.  `BIAS_LAYOUT` selects where it lives:
  "rel"  -> lin_rel.bias  (PyG >= 2.0; needed by /root/reference/tests/test_sparse_gcm.py:326-330)
  "root" -> lin_root.bias (PyG 1.x;   needed by /root/reference/tests/test_gcm.py:206,264)
"""
import re
import torch

BIAS_LAYOUT = "rel"


def _lins(in_channels, out_channels, bias):
    rel_bias = bias and BIAS_LAYOUT == "rel"
    root_bias = bias and BIAS_LAYOUT == "root"
    return (
        torch.nn.Linear(in_channels, out_channels, bias=rel_bias),
        torch.nn.Linear(in_channels, out_channels, bias=root_bias),
    )


class DenseGraphConv(torch.nn.Module):
    def __init__(self, in_channels, out_channels, aggr="add", bias=True):
        super().__init__()
        assert aggr == "add"
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_rel, self.lin_root = _lins(in_channels, out_channels, bias)

    def forward(self, x, adj, mask=None):
        x = x.unsqueeze(0) if x.dim() == 2 else x
        adj = adj.unsqueeze(0) if adj.dim() == 2 else adj
        out = torch.matmul(adj.to(x.dtype), x)
        out = self.lin_rel(out)
        out = out + self.lin_root(x)
        if mask is not None:
            out = out * mask.view(x.shape[0], x.shape[1], 1).to(x.dtype)
        return out


class DenseGCNConv(torch.nn.Module):
    """out = D^-1/2 (A + I) D^-1/2 X W + b (published DenseGCNConv definition)."""

    def __init__(self, in_channels, out_channels, improved=False, bias=True):
        super().__init__()
        self.lin = torch.nn.Linear(in_channels, out_channels, bias=False)
        self.bias = torch.nn.Parameter(torch.zeros(out_channels)) if bias else None
        self.improved = improved

    def forward(self, x, adj, mask=None, add_loop=True):
        x = x.unsqueeze(0) if x.dim() == 2 else x
        adj = adj.unsqueeze(0) if adj.dim() == 2 else adj
        B, N, _ = adj.size()
        if add_loop:
            adj = adj.clone()
            idx = torch.arange(N, dtype=torch.long, device=adj.device)
            adj[:, idx, idx] = 1 if not self.improved else 2
        out = self.lin(x)
        deg_inv_sqrt = adj.sum(dim=-1).clamp(min=1).pow(-0.5)
        adj = deg_inv_sqrt.unsqueeze(-1) * adj * deg_inv_sqrt.unsqueeze(-2)
        out = torch.matmul(adj, out)
        if self.bias is not None:
            out = out + self.bias
        return out


class GraphConv(torch.nn.Module):
    def __init__(self, in_channels, out_channels, aggr="add", bias=True):
        super().__init__()
        assert aggr == "add"
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_rel, self.lin_root = _lins(in_channels, out_channels, bias)

    def forward(self, x, edge_index, edge_weight=None):
        src, dst = edge_index[0], edge_index[1]
        msg = x[src]
        if edge_weight is not None:
            msg = edge_weight.view(-1, 1) * msg
        agg = torch.zeros(x.shape[0], x.shape[1], dtype=x.dtype, device=x.device)
        agg = agg.index_add(0, dst, msg)
        return self.lin_rel(agg) + self.lin_root(x)


class Sequential(torch.nn.Module):
    """Published PyG `Sequential(input_args, [(module, "a, b -> c"), module, ...])`."""

    def __init__(self, input_args, modules):
        super().__init__()
        self._in = [a.strip() for a in input_args.split(",")]
        self._specs = []
        self._callables = []
        mods = torch.nn.ModuleList()
        for i, m in enumerate(modules):
            if isinstance(m, (tuple, list)):
                fn, desc = m
                lhs, rhs = desc.split("->")
                ins = [a.strip() for a in lhs.split(",") if a.strip()]
                outs = [a.strip() for a in rhs.split(",") if a.strip()]
            else:
                fn, ins, outs = m, None, None
            self._specs.append((ins, outs))
            if isinstance(fn, torch.nn.Module):
                mods.append(fn)
                self._callables.append(None)
            else:
                mods.append(torch.nn.Identity())
                self._callables.append(fn)
        self.mods = mods

    # the reference's tests do `list(self.g.modules())[1]` to reach the layer list
    def forward(self, *args):
        env = dict(zip(self._in, args))
        last = None
        for (ins, outs), mod, fn in zip(self._specs, self.mods, self._callables):
            f = fn if fn is not None else mod
            if ins is None:
                res = f(*last) if isinstance(last, tuple) else f(last)
                last = res
                # bare module: rebind to the previous outputs' names
                if self._last_outs:
                    if isinstance(res, tuple):
                        for k, v in zip(self._last_outs, res):
                            env[k] = v
                    else:
                        env[self._last_outs[0]] = res
            else:
                res = f(*[env[k] for k in ins])
                last = res
                if isinstance(res, tuple):
                    for k, v in zip(outs, res):
                        env[k] = v
                else:
                    env[outs[0]] = res
                self._last_outs = outs
        return last

    _last_outs = None


def knn(*a, **k):  # torch_cluster-backed in PyG; out of scope
    raise NotImplementedError("knn needs torch_cluster; out of scope")
