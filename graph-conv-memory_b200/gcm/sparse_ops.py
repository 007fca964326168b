"""Host side of the sparse GCM kernels (csrc/gcm_sparse.cu): edge generation, CSR bookkeeping and
the GraphConv autograd function.  torch is used for allocation and for the small index arithmetic
(cumsum / bincount over per-graph or per-row counters); gathers, reductions, the Linear layers and
their gradients run in the CUDA kernels behind the C ABI.  CUDA only, no CPU fallback.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from gcm import _cabi


def _excl_cumsum(v: torch.Tensor) -> torch.Tensor:
    out = torch.zeros(v.numel() + 1, dtype=torch.long, device=v.device)
    torch.cumsum(v, 0, out=out[1:])
    return out


def ragged_arange(lengths: torch.Tensor, total: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(owner, position-within-owner) of every element of a ragged batch with the given lengths."""
    lengths = lengths.long()
    owners = torch.repeat_interleave(torch.arange(lengths.numel(), device=lengths.device), lengths,
                                     output_size=total)
    starts = torch.cumsum(lengths, 0) - lengths
    pos = torch.arange(owners.numel(), device=lengths.device) - starts[owners]
    return owners, pos


def write_flatten(nodes: torch.Tensor, x: torch.Tensor, T: torch.Tensor, taus: torch.Tensor,
                  offsets: torch.Tensor, n_flat: int, want_flat: bool = True):
    """In place: nodes[b, T_b + k] = x[b, k]; returns flat [n_flat, F] (valid rows of every graph)."""
    _cabi.require_cuda(nodes, "SparseGCM nodes")
    B, N, F = nodes.shape
    flat = torch.empty(n_flat, F, device=nodes.device, dtype=torch.float32) if want_flat else None
    _cabi.check(_cabi.lib().gcm_sparse_write_flatten(nodes.data_ptr(), x.data_ptr(), T.data_ptr(), taus.data_ptr(),
                                                     offsets.data_ptr(), B, N, F, x.shape[1], _cabi.ptr(flat),
                                                     _cabi.stream_ptr(nodes.device)), "gcm_sparse_write_flatten")
    return flat


def write_flatten_oop(nodes: torch.Tensor, x: torch.Tensor, T: torch.Tensor, taus: torch.Tensor, offsets: torch.Tensor,
                      n_flat: int, alias_full: bool = False):
    """(nodes_out, flat): nodes_out = nodes with nodes_out[b, T_b + k] = x[b, k], flat = its valid rows; one pass, no clone.
    alias_full: when every graph is full after the write (n_flat == B * N) the flat layout IS nodes_out: return a view of it
    instead of writing a second copy (callers that record autograd keep the two tensors apart)."""
    _cabi.require_cuda(nodes, "SparseGCM nodes")
    B, N, F = nodes.shape
    nodes = nodes.contiguous()
    out = torch.empty_like(nodes)
    alias = alias_full and n_flat == B * N
    flat = None if alias else torch.empty(n_flat, F, device=nodes.device, dtype=torch.float32)
    _cabi.check(_cabi.lib().gcm_sparse_write_flatten_oop(
        nodes.data_ptr(), out.data_ptr(), x.data_ptr(), T.data_ptr(), taus.data_ptr(), offsets.data_ptr(), B, N, F,
        x.shape[1], _cabi.ptr(flat), _cabi.stream_ptr(nodes.device)), "gcm_sparse_write_flatten_oop")
    return out, (out.view(B * N, F) if alias else flat)


class _WriteFlattenFn(torch.autograd.Function):
    """nodes_out = nodes with the new observations written; flat = valid rows of nodes_out."""

    @staticmethod
    def forward(ctx, nodes, x, T, taus, offsets, n_flat):
        out, flat = write_flatten_oop(nodes.detach(), x.detach().contiguous(), T, taus, offsets, n_flat)
        ctx.meta = (T, taus, offsets, nodes.shape, x.shape)
        return out, flat

    @staticmethod
    def backward(ctx, d_nodes_out, d_flat):
        # one pass (gcm_sparse_write_flatten_bwd) instead of ragged index arithmetic over [B,N,F] tensors in torch
        T, taus, offsets, nshape, xshape = ctx.meta
        B, N, F = nshape
        dev = T.device
        want_nodes, want_x = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (want_nodes or want_x):
            return None, None, None, None, None, None
        dn = None if d_nodes_out is None else d_nodes_out.contiguous().float()
        df = None if d_flat is None else d_flat.contiguous().float()
        d_nodes = torch.empty(nshape, device=dev) if want_nodes else None
        d_x = torch.empty(xshape, device=dev) if want_x else None
        _cabi.check(_cabi.lib().gcm_sparse_write_flatten_bwd(
            _cabi.ptr(dn), _cabi.ptr(df), T.data_ptr(), taus.data_ptr(), offsets.data_ptr(), B, N, F, xshape[1],
            _cabi.ptr(d_nodes), _cabi.ptr(d_x), _cabi.stream_ptr(dev)), "gcm_sparse_write_flatten_bwd")
        return d_nodes, d_x, None, None, None, None


HIT_CAP = 256    # sources kept per new node by pass 1 of the edge builder (a 512-byte slot per node, of which only the
                 # first `degree` entries are touched).  64 left 3.5 % of the nodes of cfg5 (degrees average 20 and reach
                 # 217) -- and with them a second search pass of 1.2 ms -- to the overflow path


def build_edges(nodes: torch.Tensor, T: torch.Tensor, taus: torch.Tensor, new_off: torch.Tensor, n_new: int,
                tmax: int, hops: Sequence[int], radius: Optional[Tuple[slice, float]],
                flat_off: Optional[torch.Tensor] = None):
    """Coalesced edges (batch, sink, source) int64 [3, E] of the new nodes (gcm_sparse_build_edges).
    With flat_off ([B+1] exclusive cumsum of T + tau) also returns (edge_off [n_new+1], flat_col [E]): the CSR of
    the new rows over the flat node numbering, written by the same kernel pass."""
    _cabi.require_cuda(nodes, "sparse edge selector")
    B, N, F = nodes.shape
    dev = nodes.device
    if n_new == 0 or (not hops and radius is None):
        empty = torch.zeros(3, 0, dtype=torch.long, device=dev)
        if flat_off is None:
            return empty
        return empty, torch.zeros(n_new + 1, dtype=torch.long, device=dev), empty[0]
    hops_t = torch.tensor(sorted(set(int(h) for h in hops)), dtype=torch.int32)
    hops_c = (_cabi.C.c_int32 * max(len(hops_t), 1))(*hops_t.tolist())
    use_r, p0, pst, pl, rad = 0, 0, 1, 0, 0.0
    if radius is not None:
        sl, rad = radius
        p0, p1, pst = sl.indices(F)
        pl = len(range(p0, p1, pst))
        use_r = 1
        if pst < 1 or pl < 1:
            raise _cabi.GcmLibraryError("SpatialRadiusEdge: empty or negative-step position slice")
    nodes_c = nodes.detach().contiguous()
    lib = _cabi.lib()
    stream = _cabi.stream_ptr(dev)
    deg = torch.empty(n_new, dtype=torch.int32, device=dev)
    args = (nodes_c.data_ptr(), T.data_ptr(), taus.data_ptr(), new_off.data_ptr(), B, N, F, tmax, hops_c,
            len(hops_t), use_r, p0, pst, pl, float(rad))
    # pass 1 also keeps every new node's sources as a short list (uint16, HIT_CAP per node): if no node has more,
    # pass 2 is a plain expansion of those lists instead of a second search
    hits = torch.empty(n_new, HIT_CAP, dtype=torch.int16, device=dev) if N <= 65536 else None
    _cabi.check(lib.gcm_sparse_build_edges(*args, deg.data_ptr(), None, None, 0, None, None, _cabi.ptr(hits), HIT_CAP,
                                           stream), "gcm_sparse_build_edges")
    edge_off = _excl_cumsum(deg)
    E, max_deg = (int(v) for v in torch.stack([edge_off[-1], deg.max().long()]).tolist())
    edges = torch.empty(3, E, dtype=torch.long, device=dev)
    flat_col = torch.empty(E, dtype=torch.long, device=dev) if flat_off is not None else None
    if E and hits is not None:
        _cabi.check(lib.gcm_sparse_expand_edges(T.data_ptr(), taus.data_ptr(), new_off.data_ptr(), B, tmax,
                                                hits.data_ptr(), HIT_CAP, edge_off.data_ptr(), edges.data_ptr(), E,
                                                _cabi.ptr(flat_off), _cabi.ptr(flat_col), stream),
                    "gcm_sparse_expand_edges")
    if E and (hits is None or max_deg > HIT_CAP):
        # nodes with more than HIT_CAP sources (or all of them without lists) are searched a second time
        _cabi.check(lib.gcm_sparse_build_edges(*args, None, edge_off.data_ptr(), edges.data_ptr(), E,
                                               _cabi.ptr(flat_off), _cabi.ptr(flat_col), None,
                                               0 if hits is None else HIT_CAP, stream), "gcm_sparse_build_edges")
    if flat_off is None:
        return edges
    return edges, edge_off, flat_col


class Csr:
    """Edges grouped by sink over the flat node numbering, plus (lazily) the transposed grouping."""

    def __init__(self, rowptr: torch.Tensor, col: torch.Tensor, n: int, node_off: Optional[torch.Tensor] = None,
                 max_nodes: int = 0, sink_local: Optional[torch.Tensor] = None):
        self.rowptr, self.col, self.n = rowptr, col, n
        self.sink_local = sink_local    # [E] every edge's sink as an index inside its graph (optional, speeds transposed())
        # node_off [B+1]: first flat node of every graph, when the graph is block-diagonal with contiguous per-graph
        # edge ranges (what SparseGCM builds); lets transposed() run gcm_sparse_csr_transpose instead of a global sort
        self.node_off, self.max_nodes = node_off, max_nodes
        self._t = {}

    @classmethod
    def from_sorted_edges(cls, flat_sink: torch.Tensor, flat_src: torch.Tensor, n: int) -> "Csr":
        counts = torch.bincount(flat_sink, minlength=n) if flat_sink.numel() else torch.zeros(
            n, dtype=torch.long, device=flat_sink.device)
        return cls(_excl_cumsum(counts), flat_src.contiguous(), n)

    def transposed(self, rows: Optional[torch.Tensor]):
        """(t_rowptr [n+1], t_col [E']) grouping by SOURCE the edges whose sink is an evaluated row;
        t_col = position of that sink among the evaluated rows."""
        key = None if rows is None else rows.data_ptr()
        if key not in self._t and rows is None and self.node_off is not None and 0 < self.max_nodes <= 8192:
            t_rowptr = torch.empty_like(self.rowptr)
            t_col = torch.empty_like(self.col)
            _cabi.check(_cabi.lib().gcm_sparse_csr_transpose(
                self.rowptr.data_ptr(), self.col.data_ptr(), self.node_off.data_ptr(), _cabi.ptr(self.sink_local),
                self.node_off.numel() - 1, self.n,
                t_rowptr.data_ptr(), t_col.data_ptr(), _cabi.stream_ptr(self.col.device)), "gcm_sparse_csr_transpose")
            self._t[key] = (t_rowptr, t_col)
        if key not in self._t:
            dev = self.col.device
            if rows is None:
                deg = self.rowptr[1:] - self.rowptr[:-1]
                local = torch.repeat_interleave(torch.arange(self.n, device=dev), deg, output_size=self.col.numel())
                src = self.col
            else:
                deg = self.rowptr[rows + 1] - self.rowptr[rows]
                owner, pos = ragged_arange(deg)
                src = self.col[self.rowptr[rows][owner] + pos]
                local = owner
            perm = torch.argsort(src, stable=True)
            counts = torch.bincount(src, minlength=self.n) if src.numel() else torch.zeros(
                self.n, dtype=torch.long, device=dev)
            self._t[key] = (_excl_cumsum(counts), local[perm].contiguous())
        return self._t[key]


def _hint_blocks(csr: Csr, rows) -> None:
    """SparseGCM's CSR is block-diagonal over graphs whose rows are contiguous: tell the next forward call, which can then
    stage a graph's rows in shared memory (gcm_sparse_graphconv_hint_blocks)."""
    if rows is None and csr.node_off is not None and csr.max_nodes > 0:
        _cabi.lib().gcm_sparse_graphconv_hint_blocks(csr.node_off.data_ptr(), csr.node_off.numel() - 1, int(csr.max_nodes))


def _kmajor(w_rel: torch.Tensor, w_root: torch.Tensor) -> torch.Tensor:
    return torch.cat([w_rel.detach().t(), w_root.detach().t()], dim=0).contiguous()


_TWO_PASS_ROWS = 8192      # gcm_sparse_tc.cu: two-pass threshold (64 tiles of 128 rows)


class _GraphConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w_rel, bias, w_root, csr: Csr, rows, act: int):
        x = x.contiguous()
        n, Fin = x.shape
        Fout = w_rel.shape[0]
        m = n if rows is None else rows.numel()
        dev = x.device
        need_grad = any(t is not None and t.requires_grad for t in (x, w_rel, bias, w_root))
        # room for the neighbour sums: kept for the backward while recording; large no-grad calls get it as scratch, which
        # lets the library run its two-pass form (k_csr_gather at full occupancy + the streaming product kernel)
        agg = torch.empty(m, Fin, device=dev) if (need_grad or m >= _TWO_PASS_ROWS) else None
        out = torch.empty(m, Fout, device=dev)
        wt = _kmajor(w_rel, w_root)
        b = None if bias is None else bias.detach().contiguous()
        _cabi.lib().gcm_sparse_graphconv_hint_rows(n)
        _hint_blocks(csr, rows)
        _cabi.check(_cabi.lib().gcm_sparse_graphconv_fwd(
            x.data_ptr(), csr.rowptr.data_ptr(), csr.col.data_ptr(), None, _cabi.ptr(rows), m, Fin, Fout,
            wt.data_ptr(), _cabi.ptr(b), act, _cabi.ptr(agg), out.data_ptr(), _cabi.stream_ptr(dev)),
            "gcm_sparse_graphconv_fwd")
        ctx.csr, ctx.rows, ctx.act, ctx.has_bias = csr, rows, act, bias is not None
        ctx.save_for_backward(x, agg if need_grad else None, out, w_rel, w_root)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, agg, out, w_rel, w_root = ctx.saved_tensors
        csr, rows = ctx.csr, ctx.rows
        n, Fin = x.shape
        Fout = w_rel.shape[0]
        m = out.shape[0]
        dev = x.device
        t_rowptr, t_col = csr.transposed(rows)
        # every row evaluated (rows is None): d_x = dz W_root overwrites all n rows before the transposed gather adds to it
        d_x = (torch.empty if rows is None else torch.zeros)(n, Fin, device=dev)
        d_agg = torch.empty(m, Fin, device=dev)
        d_w_rel = torch.zeros_like(w_rel)
        d_w_root = torch.zeros_like(w_root)
        d_b = torch.zeros(Fout, device=dev) if ctx.has_bias else None
        dz = w_rel_t = w_root_t = ws = None
        if rows is None:
            dz = torch.empty(m, Fout, device=dev)
            w_rel_t = w_rel.detach().t().contiguous()
            w_root_t = w_root.detach().t().contiguous()
            if Fin % 16 == 0 and Fout % 16 == 0:
                ws = torch.empty(int(_cabi.lib().gcm_outer_reduce_tc_workspace(m)), device=dev)
        _cabi.check(_cabi.lib().gcm_sparse_graphconv_bwd(
            x.data_ptr(), agg.data_ptr(), out.data_ptr(), d_out.contiguous().data_ptr(), _cabi.ptr(rows), m, n,
            t_rowptr.data_ptr(), t_col.data_ptr(), None, Fin, Fout, w_rel.detach().contiguous().data_ptr(),
            w_root.detach().contiguous().data_ptr(), ctx.act, d_agg.data_ptr(), d_x.data_ptr(),
            d_w_rel.data_ptr(), d_w_root.data_ptr(), _cabi.ptr(d_b), _cabi.ptr(dz), _cabi.ptr(w_rel_t),
            _cabi.ptr(w_root_t), _cabi.ptr(ws), _cabi.stream_ptr(dev)),
            "gcm_sparse_graphconv_bwd")
        return d_x, d_w_rel, d_b, d_w_root, None, None, None


def graph_conv_csr(x, csr: Csr, rows, w_rel, bias, w_root, act: str = "none", edge_mask=None):
    """edge_mask float32 [E] in CSR order (0 / 1): the k-hop subgraph restriction of sparse_gcm.py:182-199 (an edge counts
    only while its source belongs to the subgraph).  Forward only: a masked call must not record autograd."""
    if edge_mask is not None:
        assert not (torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (x, w_rel, bias, w_root)))
        x = x.contiguous()
        m = x.shape[0] if rows is None else rows.numel()
        out = torch.empty(m, w_rel.shape[0], device=x.device)
        scratch = torch.empty(m, x.shape[1], device=x.device) if m >= _TWO_PASS_ROWS else None
        wt = _kmajor(w_rel, w_root)
        b = None if bias is None else bias.detach().contiguous()
        _cabi.lib().gcm_sparse_graphconv_hint_rows(x.shape[0])
        _hint_blocks(csr, rows)
        _cabi.check(_cabi.lib().gcm_sparse_graphconv_fwd(
            x.data_ptr(), csr.rowptr.data_ptr(), csr.col.data_ptr(), edge_mask.data_ptr(), _cabi.ptr(rows), m, x.shape[1],
            w_rel.shape[0], wt.data_ptr(), _cabi.ptr(b), _cabi.ACT[act], _cabi.ptr(scratch), out.data_ptr(),
            _cabi.stream_ptr(x.device)), "gcm_sparse_graphconv_fwd")
        return out
    return _GraphConvFn.apply(x, w_rel, bias, w_root, csr, rows, _cabi.ACT[act])


def graph_conv(x, edge_index, edge_weight, w_rel, bias, w_root, act: str = "none"):
    """GraphConv on a PyG-style edge_index [2, E] (row 0 = source, row 1 = sink).  Edge weights other
    than 1 are not part of the hot path (the reference forces them to 1, sparse_gcm.py:160-164)."""
    _cabi.require_cuda(x, "GraphConv")
    if edge_weight is not None and edge_weight.numel() and not bool((edge_weight == 1).all()):
        raise _cabi.GcmLibraryError("GraphConv kernels support unit edge weights only")
    n = x.shape[0]
    src, dst = edge_index[0], edge_index[1]
    order = torch.argsort(dst * n + src) if src.numel() else src
    csr = Csr.from_sorted_edges(dst[order], src[order], n)
    return graph_conv_csr(x, csr, None, w_rel, bias, w_root, act)
