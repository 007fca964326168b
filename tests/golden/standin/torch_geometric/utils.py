"""Stand-in for torch_geometric.utils (coalesce, k_hop_subgraph, dense helpers).

TEST INFRASTRUCTURE ONLY.  Follows the published PyG >= 2.0 semantics.
"""
import torch


def coalesce(edge_index, edge_attr="???", num_nodes=None, reduce="add",
             is_sorted=False, sort_by_row=True):
    """Sort edges lexicographically by (row, col), merge duplicates with `reduce`."""
    had_attr = not (isinstance(edge_attr, str) and edge_attr == "???")
    attrs = edge_attr if had_attr else None
    n = int(edge_index.max()) + 1 if edge_index.numel() else 0
    if num_nodes is not None:
        n = max(n, int(num_nodes))
    key = edge_index[0] * n + edge_index[1]
    uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
    out_index = torch.stack([uniq // max(n, 1), uniq % max(n, 1)])

    def _reduce(a):
        out = torch.zeros((uniq.numel(),) + tuple(a.shape[1:]), dtype=a.dtype, device=a.device)
        if reduce in ("add", "sum"):
            return out.index_add(0, inv, a)
        if reduce == "mean":
            s = out.index_add(0, inv, a)
            c = torch.zeros(uniq.numel(), dtype=a.dtype, device=a.device).index_add(
                0, inv, torch.ones_like(inv, dtype=a.dtype))
            return s / c.view((-1,) + (1,) * (a.dim() - 1))
        if reduce == "min":
            out = torch.full_like(out, torch.iinfo(a.dtype).max if not a.is_floating_point() else float("inf"))
            return out.scatter_reduce(0, inv, a, reduce="amin")
        if reduce == "max":
            out = torch.full_like(out, torch.iinfo(a.dtype).min if not a.is_floating_point() else float("-inf"))
            return out.scatter_reduce(0, inv, a, reduce="amax")
        raise ValueError(reduce)

    if not had_attr:
        return out_index
    if attrs is None:
        return out_index, None
    if isinstance(attrs, (list, tuple)):
        return out_index, [_reduce(a) for a in attrs]
    return out_index, _reduce(attrs)


def k_hop_subgraph(node_idx, num_hops, edge_index, relabel_nodes=False,
                   num_nodes=None, flow="source_to_target", directed=False):
    num_nodes = int(num_nodes) if num_nodes is not None else int(edge_index.max()) + 1
    if flow == "target_to_source":
        row, col = edge_index
    else:
        col, row = edge_index
    node_mask = row.new_empty(num_nodes, dtype=torch.bool)
    if isinstance(node_idx, (int, list, tuple)):
        node_idx = torch.tensor([node_idx], device=row.device).flatten()
    subsets = [node_idx]
    for _ in range(num_hops):
        node_mask.fill_(False)
        node_mask[subsets[-1]] = True
        edge_mask = node_mask[row]
        subsets.append(col[edge_mask])
    subset, inv = torch.cat(subsets).unique(return_inverse=True)
    inv = inv[: node_idx.numel()]
    node_mask.fill_(False)
    node_mask[subset] = True
    if not directed:
        edge_mask = node_mask[row] & node_mask[col]
    edge_index = edge_index[:, edge_mask]
    if relabel_nodes:
        remap = row.new_full((num_nodes,), -1)
        remap[subset] = torch.arange(subset.size(0), device=row.device)
        edge_index = remap[edge_index]
    return subset, edge_index, inv, edge_mask


def to_dense_batch(x, batch=None, fill_value=0.0, max_num_nodes=None):
    B = int(batch.max()) + 1
    counts = torch.bincount(batch, minlength=B)
    N = int(max_num_nodes) if max_num_nodes is not None else int(counts.max())
    start = torch.cumsum(counts, 0) - counts
    pos = torch.arange(batch.numel(), device=x.device) - start[batch]
    out = x.new_full((B, N) + tuple(x.shape[1:]), fill_value)
    out[batch, pos] = x
    mask = torch.zeros(B, N, dtype=torch.bool, device=x.device)
    mask[batch, pos] = True
    return out, mask


def to_dense_adj(edge_index, batch=None, edge_attr=None, max_num_nodes=None):
    B = int(batch.max()) + 1
    counts = torch.bincount(batch, minlength=B)
    start = torch.cumsum(counts, 0) - counts
    N = int(max_num_nodes) if max_num_nodes is not None else int(counts.max())
    b = batch[edge_index[0]]
    r = edge_index[0] - start[b]
    c = edge_index[1] - start[b]
    adj = torch.zeros(B, N, N, device=edge_index.device)
    adj.index_put_((b, r, c), torch.ones(b.numel(), device=edge_index.device), accumulate=True)
    return adj
