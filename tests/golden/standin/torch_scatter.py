"""Stand-in for torch_scatter (imported at /root/reference/src/gcm/util.py:4).

TEST INFRASTRUCTURE ONLY; only the learned-edge code (out of scope) calls these.
"""
import torch


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    size = int(index.max()) + 1 if dim_size is None else dim_size
    res = torch.zeros((size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return res.index_add(0, index, src)


def scatter_max(src, index, dim=0, out=None, dim_size=None):
    size = int(index.max()) + 1 if dim_size is None else dim_size
    res = torch.full((size,), float("-inf"), dtype=src.dtype, device=src.device)
    res = res.scatter_reduce(0, index, src, reduce="amax")
    arg = torch.full((size,), src.numel(), dtype=torch.long, device=src.device)
    hit = src == res[index]
    pos = torch.arange(src.numel(), device=src.device)
    arg = arg.scatter_reduce(0, index[hit], pos[hit], reduce="amin")
    return res, arg
