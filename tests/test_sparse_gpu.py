"""GPU tier: the sparse GCM path (edge generation, CSR GraphConv fwd/bwd through the C ABI) against the
golden vectors of the unmodified reference and against the oracle on seeded inputs.  Edge lists,
node slots and T bit-exact; outputs and gradients within 1e-5 relative (fp32)."""
import pytest
import torch

import gcm_oracle as oracle
from helpers import (load_golden, make_dense_gnn, make_selector, make_sparse_gnn, make_sparse_selector,
                     named_grads, rel_err, sparse_cases)

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("name", sparse_cases())
@pytest.mark.parametrize("style", ["readme", "sequential"])
def test_sparse_matches_reference_golden(name, style):
    from gcm.sparse_gcm import SparseGCM

    g = load_golden(name)
    dev = torch.device("cuda:0")
    gnn, convs = make_sparse_gnn(g["F"], g["H"], g["params"], g["acts"], style)
    mod = SparseGCM(gnn.to(dev), edge_selectors=make_sparse_selector(g["spec"]),
                    aux_edge_selectors=make_sparse_selector(g["aux"]), graph_size=g["N"], max_hops=g["max_hops"])
    assert mod.fused_plan() is not None
    hidden = None
    xs, outs = [], []
    for (x, taus), want in zip(g["calls"], g["outs"]):
        xd = x.to(dev).requires_grad_("d_x" in g)
        mx, hidden = mod(xd, taus.to(dev), hidden)
        assert rel_err(mx, want) < TOL, name
        xs.append(xd)
        outs.append(mx)
    nodes, adj, T = hidden
    assert torch.equal(nodes.detach().cpu(), g["final_nodes"])
    assert adj.is_sparse and tuple(adj.shape) == (g["B"], g["N"], g["N"])
    assert torch.equal(adj.coalesce().indices().cpu(), g["final_edges"])       # (batch, sink, source), coalesced
    assert torch.equal(adj.coalesce().values().cpu(), torch.ones(g["final_edges"].shape[1]))
    assert torch.equal(T.cpu(), g["final_T"]) and T.dtype == torch.long
    if "d_x" in g:
        sum((o * w.to(dev)).sum() for o, w in zip(outs, g["loss_w"])).backward()
        for xd, want in zip(xs, g["d_x"]):
            assert rel_err(xd.grad, want) < 5 * TOL, name
        got = named_grads(convs)
        for k, v in g["d_params"].items():
            assert rel_err(got[k], v) < 5 * TOL, (name, k)


SEEDED = [
    # B, N, F, H, calls (list of taus), temporal hops, radius
    (5, 64, 64, 64, [[64, 40, 1, 17, 33]], (1,), 0.25),             # BASELINE cfg 5 shape, reduced N/B
    (3, 96, 16, 32, [[30, 20, 10], [25, 30, 5], [40, 46, 81]], (1, 2, 4), None),
    (2, 300, 8, 24, [[300, 150]], (), 0.4),
    (4, 20, 100, 128, [[5, 5, 5, 5], [1, 2, 3, 4]], (2,), 0.6),     # widest supported layer
    (3, 12, 6, 7, [[0, 3, 12]], (1,), None),                        # an empty update in the batch
]


@pytest.mark.parametrize("B,N,F,H,calls,hops,radius", SEEDED)
def test_sparse_matches_oracle_seeded(B, N, F, H, calls, hops, radius):
    from gcm.sparse_gcm import SparseGCM

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(77 + N + F)
    p = oracle.make_params(F, H)
    spec = [("temporal", hops)] if hops else None
    aux = [("spatial_radius", slice(0, 2), radius)] if radius else None
    gnn, convs = make_sparse_gnn(F, H, p, ("tanh", "tanh"))
    mod = SparseGCM(gnn.to(dev), edge_selectors=make_sparse_selector(spec), aux_edge_selectors=make_sparse_selector(aux),
                    graph_size=N)
    p_or = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    hidden, o_hidden = None, None
    pairs, loss, o_loss = [], 0, 0
    for taus in calls:
        taus = torch.tensor(taus)
        tmax = int(taus.max())
        x = torch.randn(B, tmax, F, generator=gen) * 0.5
        x[..., 0:2] = torch.cumsum(0.1 * torch.randn(B, tmax, 2, generator=gen), dim=1)
        for b in range(B):
            x[b, int(taus[b]):] = 0
        w = torch.randn(B, tmax, H, generator=gen)
        xd = x.to(dev).requires_grad_(True)
        xo = x.clone().requires_grad_(True)
        mx, hidden = mod(xd, taus.to(dev), hidden)
        omx, o_hidden = oracle.sparse_gcm_forward(xo, taus, o_hidden, spec, p_or, graph_size=N, aux_selectors=aux)
        assert rel_err(mx, omx) < TOL
        assert torch.equal(hidden[0].detach().cpu(), o_hidden[0].detach())
        assert torch.equal(hidden[1].coalesce().indices().cpu(), o_hidden[1])
        assert torch.equal(hidden[2].cpu(), o_hidden[2])
        loss = loss + (mx * w.to(dev)).sum()
        o_loss = o_loss + (omx * w).sum()
        pairs.append((xd, xo))
    loss.backward()
    o_loss.backward()
    for xd, xo in pairs:
        assert rel_err(xd.grad, xo.grad) < 5 * TOL
    got = named_grads(convs)
    for k in got:
        assert rel_err(got[k], p_or[k].grad) < 5 * TOL, k


def test_dense_and_sparse_agree():
    """Reference tests/test_sparse_gcm.py:427-462 (TestDenseVsSparse.test_temporal_edges): the same weights in
    DenseGraphConv and GraphConv stacks give the same beliefs, nodes and edge sets."""
    from gcm.gcm import DenseGCM
    from gcm.sparse_gcm import SparseGCM

    dev = torch.device("cuda:0")
    F, B, ts, N = 3, 3, 8, 8
    p = oracle.make_params(F, F)
    dgnn, _ = make_dense_gnn(F, F, p, ("none", "none"), style="sequential")
    sgnn, _ = make_sparse_gnn(F, F, p, ("none", "none"), style="sequential")
    sgnn.load_state_dict(dgnn.state_dict())                      # same state_dict keys, like PyG >= 2.0
    dense = DenseGCM(dgnn.to(dev), edge_selectors=make_selector([("temporal", (1, 2), "forward")]), graph_size=N)
    sparse = SparseGCM(sgnn.to(dev), edge_selectors=make_sparse_selector([("temporal", (1, 2))]), graph_size=N)
    obs = torch.arange(B * ts * F, dtype=torch.float32).reshape(B, ts, F).to(dev) * 0.01
    with torch.no_grad():
        d_outs, d_hidden = [], None
        for i in range(ts):
            o, d_hidden = dense(obs[:, i].contiguous(), d_hidden)
            d_outs.append(o)
        d_outs = torch.stack(d_outs, dim=1)
        s_outs, s_hidden = sparse(obs, torch.full((B,), ts, device=dev), None)
        step_outs, step_hidden = [], None
        for i in range(ts):
            o, step_hidden = sparse(obs[:, i:i + 1].contiguous(), torch.ones(B, dtype=torch.long, device=dev), step_hidden)
            step_outs.append(o)
        step_outs = torch.cat(step_outs, dim=1)
    assert rel_err(s_outs, d_outs) < TOL and rel_err(step_outs, d_outs) < TOL
    d_nodes, d_adj, _, _ = d_hidden
    assert torch.equal(d_nodes, s_hidden[0]) and torch.equal(d_nodes, step_hidden[0])
    assert torch.equal(d_adj.nonzero().T, s_hidden[1].coalesce().indices())
    assert torch.equal(d_adj.nonzero().T, step_hidden[1].coalesce().indices())


def test_sparse_overflow_raises():
    from gcm.sparse_gcm import SparseGCM

    dev = torch.device("cuda:0")
    p = oracle.make_params(4, 4)
    gnn, _ = make_sparse_gnn(4, 4, p, ("tanh", "tanh"))
    mod = SparseGCM(gnn.to(dev), edge_selectors=make_sparse_selector([("temporal", (1,))]), graph_size=5)
    with pytest.raises(Exception, match="Overflow"):
        mod(torch.randn(2, 6, 4, device=dev), torch.tensor([6, 2], device=dev), None)


def test_sparse_selectors_standalone_return_coo():
    from gcm.sparse_edge_selectors.spatial import SpatialRadiusEdge
    from gcm.sparse_edge_selectors.temporal import TemporalEdge

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(9)
    B, N, F = 3, 10, 4
    nodes = torch.randn(B, N, F, generator=gen) * 0.3
    T, taus = torch.tensor([2, 0, 5]), torch.tensor([3, 4, 5])
    want_t = oracle.temporal_edges_sparse(T, taus, (1, 3))
    got = TemporalEdge([1, 3])(nodes.to(dev), T.to(dev), taus.to(dev), B)
    assert got.is_sparse and torch.equal(got.coalesce().indices().cpu(), oracle.coalesce_edges(want_t))
    want_r = oracle.spatial_radius_edges_sparse(nodes, T, taus, slice(0, 2), 0.5)
    got = SpatialRadiusEdge(slice(0, 2), 0.5)(nodes.to(dev), T.to(dev), taus.to(dev), B)
    assert torch.equal(got.coalesce().indices().cpu(), oracle.coalesce_edges(want_r))


@pytest.mark.parametrize("B,N,PL,radius,walk", [(6, 700, 2, 0.25, True), (3, 4096, 2, 0.25, True), (4, 512, 3, 0.6, False),
                                                (5, 300, 1, 0.02, False), (3, 260, 2, 1.0e-3, True), (2, 1024, 2, 50.0, False)])
def test_spatial_hash_edges_equal_all_pairs_edges(B, N, PL, radius, walk):
    """gcm_sparse_build_edges has two kernels for the radius selector (include/gcm_b200.h: gcm_edge_builder):
    the all-pairs test and the spatial hash.  They must emit identical coalesced edge lists: clustered random
    walks (BASELINE cfg5's observations), uniform clouds, 1-D / 3-D positions, ragged T / tau, a radius that
    catches nothing and one that catches everything, with temporal hops merged in."""
    from gcm import _cabi, sparse_ops

    dev = torch.device("cuda:0")
    lib = _cabi.lib()
    gen = torch.Generator().manual_seed(N * 7 + PL)
    F = 8
    nodes = torch.randn(B, N, F, generator=gen)
    if walk:
        nodes[..., 1:1 + PL] = torch.cumsum(0.1 * torch.randn(B, N, PL, generator=gen), dim=1)
    else:
        nodes[..., 1:1 + PL] = torch.rand(B, N, PL, generator=gen) * 4 - 2
    nodes[0, 5, 1] = float("nan")                       # NaN positions never match (NaN < r is false)
    T = torch.randint(0, N // 3, (B,), generator=gen)
    taus = torch.tensor([int(torch.randint(1, N - int(t) + 1, (1,), generator=gen)) for t in T])
    taus[0] = N - int(T[0])                              # one graph filled to the brim
    nodes, T, taus = nodes.to(dev), T.to(dev), taus.to(dev)
    new_off = sparse_ops._excl_cumsum(taus)
    n_new, tmax = int(taus.sum()), int(taus.max())
    got = {}
    saved_cap = sparse_ops.HIT_CAP
    try:
        for name, which in (("pairs", _cabi.EB_PAIRS), ("hash", _cabi.EB_HASH)):
            _cabi.check(lib.gcm_set_edge_builder(which), "gcm_set_edge_builder")
            # HIT_CAP = 1: some node has more sources than the per-node list holds -> pass 2 searches again (the
            # original two-pass scheme); HIT_CAP = 4096: pass 2 only expands the lists written by pass 1
            for cap in (1, 4096):
                sparse_ops.HIT_CAP = cap
                flat_off = sparse_ops._excl_cumsum(T + taus)
                e, edge_off, flat_col = sparse_ops.build_edges(nodes, T, taus, new_off, n_new, tmax, (1, 3),
                                                               (slice(1, 1 + PL), radius), flat_off)
                got[(name, cap)] = e
                want_kernel = {1: "k_sparse_edges_hash" if name == "hash" else "k_sparse_edges",
                               4096: "k_sparse_edges_expand"}[cap]
                assert lib.gcm_last_kernel().decode() == want_kernel      # cap 1: expansion + a second search
                assert torch.equal(flat_col, flat_off[e[0]] + e[2])                 # CSR columns over the flat numbering
                assert int(edge_off[-1]) == e.shape[1]
    finally:
        lib.gcm_set_edge_builder(_cabi.EB_AUTO)
        sparse_ops.HIT_CAP = saved_cap
    assert got[("pairs", 1)].shape[1] > 0
    for k in got:
        assert torch.equal(got[("pairs", 1)], got[k]), k


@pytest.mark.parametrize("N,taus,radius,cap,min_edges", [
    (300, [300, 1, 257, 0, 64, 299, 128], 0.3, 0, 1000),
    # > 8192 edges per graph with the smallest row buffer: several source ranges per graph in k_csr_transpose_smem
    (1500, [1500, 0, 1499, 700], 0.8, 8192, 3 * 8192)])
def test_csr_transpose_kernel_equals_the_global_sort(N, taus, radius, cap, min_edges):
    """gcm_sparse_csr_transpose (per-graph counting sort + per-row sort of the sinks, in shared memory when the builder's
    sink list is at hand) must give exactly the grouping the stable global argsort gives: same row pointers, sinks
    ascending within every source.  Ragged graphs, one empty."""
    from gcm import _cabi, sparse_ops

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(5)
    B, F = len(taus), 8
    nodes = torch.randn(B, N, F, generator=gen)
    nodes[..., 0:2] = torch.cumsum(0.1 * torch.randn(B, N, 2, generator=gen), dim=1)
    T = torch.zeros(B, dtype=torch.long)
    taus = torch.tensor(taus)
    nodes, T, taus = nodes.to(dev), T.to(dev), taus.to(dev)
    new_off = sparse_ops._excl_cumsum(taus)
    offsets = sparse_ops._excl_cumsum(T + taus)
    n = int(taus.sum())
    edges, edge_off, flat_col = sparse_ops.build_edges(nodes, T, taus, new_off, n, int(taus.max()), (1, 2),
                                                       (slice(0, 2), radius), offsets)
    assert edges.shape[1] > min_edges
    try:
        _cabi.check(_cabi.lib().gcm_set_csr_transpose_cap(cap), "gcm_set_csr_transpose_cap")
        fast = sparse_ops.Csr(edge_off, flat_col, n, node_off=offsets, max_nodes=N).transposed(None)
        assert _cabi.lib().gcm_last_kernel().decode() == "k_csr_transpose"
        fast2 = sparse_ops.Csr(edge_off, flat_col, n, node_off=offsets, max_nodes=N,
                               sink_local=edges[1].contiguous()).transposed(None)
        assert _cabi.lib().gcm_last_kernel().decode() == "k_csr_transpose_smem"
    finally:
        _cabi.lib().gcm_set_csr_transpose_cap(0)
    slow = sparse_ops.Csr(edge_off, flat_col, n).transposed(None)
    for got in (fast, fast2):
        assert torch.equal(got[0], slow[0]) and torch.equal(got[1], slow[1])


@pytest.mark.parametrize("name", ["pack_ragged", "pack_wide"])
def test_pack_unpack_kernels_match_reference_golden(name):
    """RLlib state wire format: gcm_pack_edges / gcm_unpack_edges behind util.pack_hidden /
    util.unpack_hidden against the fixtures written by the unmodified reference (util.py:323-382), bit for bit."""
    from helpers import load_golden
    from gcm import _cabi, util

    dev = torch.device("cuda:0")
    g = load_golden(name)
    adj = torch.sparse_coo_tensor(g["indices"], g["values"], size=(g["B"], g["N"], g["N"])).to(dev)
    _, edges, weights, _ = util.pack_hidden((g["nodes"].to(dev), adj, g["T"].to(dev)), g["B"], g["max_edges"])
    assert _cabi.lib().gcm_last_kernel().decode() == "k_pack_edges"
    assert edges.dtype == torch.long and torch.equal(edges.cpu(), g["edges"]) and torch.equal(weights.cpu(), g["weights"])
    _, adj2, _ = util.unpack_hidden((g["nodes"].to(dev), edges, weights, g["T"].to(dev)), g["B"])
    assert _cabi.lib().gcm_last_kernel().decode() == "k_unpack_edges"
    assert torch.equal(adj2._indices().cpu(), g["unpacked_indices"]) and torch.equal(adj2._values().cpu(), g["unpacked_values"])
    with pytest.raises(AssertionError, match="Cannot pack"):
        util.pack_hidden((g["nodes"].to(dev), adj, g["T"].to(dev)), g["B"], 3)


def test_pack_unpack_round_trip_at_scale_and_through_sparse_gcm():
    """B = 1024 graphs, ragged edge counts up to 200: pack -> unpack is the identity on the coalesced COO, agrees with the
    oracle's per-graph loops on a sample of graphs, and a SparseGCM hidden state survives the trip RaySparseGCM.forward
    makes around every call (ray_sparse_gcm.py:195-213): the next step's output is unchanged."""
    import gcm_oracle as oracle
    from gcm import util

    dev = torch.device("cuda:0")
    B, N, M = 1024, 64, 256
    gen = torch.Generator().manual_seed(12)
    n_e = torch.randint(0, 200, (B,), generator=gen)
    b_idx = torch.repeat_interleave(torch.arange(B), n_e)
    flat = torch.cat([torch.randperm(N * N, generator=gen)[:int(k)] for k in n_e])
    idx = torch.stack([b_idx, flat // N, flat % N])
    vals = torch.rand(idx.shape[1], generator=gen)
    adj = torch.sparse_coo_tensor(idx, vals, size=(B, N, N)).coalesce()
    nodes, T = torch.zeros(B, N, 2), torch.zeros(B, dtype=torch.long)
    _, edges, weights, _ = util.pack_hidden((nodes.to(dev), adj.to(dev), T.to(dev)), B, M)
    _, back, _ = util.unpack_hidden((nodes.to(dev), edges, weights, T.to(dev)), B)
    assert torch.equal(back._indices().cpu(), adj.indices()) and torch.equal(back._values().cpu(), adj.values())
    sample = [0, 1, 511, 1023]
    sub = torch.sparse_coo_tensor(torch.cat([torch.stack([torch.full_like(adj.indices()[0][adj.indices()[0] == b], i),
                                                          adj.indices()[1][adj.indices()[0] == b],
                                                          adj.indices()[2][adj.indices()[0] == b]]) for i, b in enumerate(sample)], 1),
                                  torch.cat([adj.values()[adj.indices()[0] == b] for b in sample]), size=(len(sample), N, N))
    _, e_o, w_o, _ = oracle.pack_hidden((nodes[sample], sub, T[sample]), len(sample), M)
    assert torch.equal(edges[sample].cpu(), e_o) and torch.equal(weights[sample].cpu(), w_o)
    # through SparseGCM: step, pack + unpack the hidden state, step again == stepping without the trip
    from helpers import make_sparse_gnn, make_sparse_selector
    from gcm.sparse_gcm import SparseGCM
    p = oracle.make_params(6, 8)
    outs = []
    for trip in (False, True):
        gnn, _ = make_sparse_gnn(6, 8, p, ("tanh", "tanh"))
        mod = SparseGCM(gnn.to(dev), edge_selectors=make_sparse_selector([("temporal", (1, 2))]), graph_size=16)
        g2 = torch.Generator().manual_seed(3)
        x1, x2 = torch.randn(5, 4, 6, generator=g2).to(dev), torch.randn(5, 3, 6, generator=g2).to(dev)
        taus1, taus2 = torch.tensor([4, 2, 3, 4, 1], device=dev), torch.tensor([3, 3, 1, 2, 3], device=dev)
        with torch.no_grad():
            _, hid = mod(x1, taus1, None)
            if trip:
                hid = util.unpack_hidden(util.pack_hidden(hid, 5, 64), 5)
            o, hid2 = mod(x2, taus2, hid)
        outs.append((o, hid2[1].coalesce().indices()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("F,H,act,use_rows,masked,n", [(64, 64, "tanh", False, False, 128 * 9 + 37),
                                                        (32, 32, "relu", False, False, 128 * 9 + 37),
                                                        (64, 32, "none", True, False, 128 * 9 + 37),
                                                        (32, 64, "tanh", True, True, 128 * 9 + 37),
                                                        # >= 8192 evaluated rows: the two-pass form (k_csr_gather + product)
                                                        (64, 64, "tanh", False, False, 128 * 70 + 37),
                                                        (32, 48, "relu", True, False, 128 * 150 + 5),
                                                        (64, 64, "none", False, True, 128 * 70 + 37)])
def test_graphconv_tensor_core_kernel_matches_cuda_core_kernel(F, H, act, use_rows, masked, n):
    """k_graphconv_fwd_tc (tcgen05, 3xTF32, SS-form MMAs on a padded K-major tile) against k_graphconv_fwd (CUDA cores) and
    the definition out = act(W_rel sum_j w_j x_j + b + W_root x_i), on a random block-diagonal causal graph: all rows /
    a row subset, unit weights / a 0-1 edge mask, a ragged tail tile."""
    from gcm import _cabi, sparse_ops

    dev = torch.device("cuda:0")
    lib = _cabi.lib()
    gen = torch.Generator().manual_seed(F + H)
    deg = torch.randint(0, 40, (n,), generator=gen)
    deg[n - 3] = 75                               # more than two 32-edge chunks
    deg[0] = 0
    sink = torch.repeat_interleave(torch.arange(n), deg)
    src = (torch.rand(sink.numel(), generator=gen) * sink.float()).long().clamp(max=n - 1)     # source < sink
    order = torch.argsort(sink * n + src)
    sink, src = sink[order], src[order]
    csr = sparse_ops.Csr.from_sorted_edges(sink.to(dev), src.to(dev), n)
    x = torch.randn(n, F, generator=gen).to(dev)
    w_rel = (torch.randn(H, F, generator=gen) / F ** 0.5).to(dev)
    w_root = (torch.randn(H, F, generator=gen) / F ** 0.5).to(dev)
    bias = torch.randn(H, generator=gen).to(dev)
    rows = torch.randperm(n, generator=gen)[: (128 * 5 + 11 if n < 8192 else 128 * 66 + 11)].sort().values.to(dev) if use_rows else None
    mask = (torch.rand(sink.numel(), generator=gen) < 0.7).float().to(dev) if masked else None
    outs = {}
    try:
        for which in (_cabi.GC_CUDA_CORES, _cabi.GC_TC):
            lib.gcm_set_graphconv_kernel(which)
            with torch.no_grad():
                outs[which] = sparse_ops.graph_conv_csr(x, csr, rows, w_rel, bias, w_root, act, edge_mask=mask)
            want = "k_graphconv_fwd_tc" if which == _cabi.GC_TC else "k_graphconv_fwd"
            assert lib.gcm_last_kernel().decode() == want
    finally:
        lib.gcm_set_graphconv_kernel(_cabi.GC_AUTO)
    w = torch.ones(sink.numel(), device=dev) if mask is None else mask
    agg = torch.zeros(n, F, device=dev, dtype=torch.float64).index_add_(0, sink.to(dev), (x[src.to(dev)] * w[:, None]).double())
    ref = agg @ w_rel.double().t() + bias.double() + x.double() @ w_root.double().t()
    ref = {"tanh": torch.tanh, "relu": torch.relu, "none": lambda t: t}[act](ref)
    if rows is not None:
        ref = ref[rows]
    a, b = outs[_cabi.GC_CUDA_CORES], outs[_cabi.GC_TC]
    scale = max(1.0, float(ref.abs().max()))
    assert float((b.double() - ref).abs().max()) / scale < 2e-5
    assert float((a - b).abs().max()) / scale < 1e-5
    if mask is None:
        # recording: the tensor-core forward also hands the aggregation to the backward; same gradients from both kernels
        grads = {}
        try:
            for which in (_cabi.GC_CUDA_CORES, _cabi.GC_TC):
                lib.gcm_set_graphconv_kernel(which)
                xg = x.clone().requires_grad_(True)
                wr, wo, bb = (t.clone().requires_grad_(True) for t in (w_rel, w_root, bias))
                out = sparse_ops.graph_conv_csr(xg, csr, rows, wr, bb, wo, act)
                (out * torch.linspace(-1, 1, out.numel(), device=dev).view_as(out)).sum().backward()
                grads[which] = (xg.grad, wr.grad, wo.grad, bb.grad)
        finally:
            lib.gcm_set_graphconv_kernel(_cabi.GC_AUTO)
        for ga, gb in zip(grads[_cabi.GC_CUDA_CORES], grads[_cabi.GC_TC]):
            # sums of ~1200 signed terms of size ~1: the two forwards differ by ~1e-6, which the cancellation amplifies
            assert float((ga - gb).abs().max()) / max(1.0, float(ga.abs().max())) < 1e-4


@pytest.mark.parametrize("F,H,act,masked", [(64, 64, "tanh", False), (32, 48, "relu", True)])
def test_graphconv_block_local_kernel_matches_the_others(F, H, act, masked):
    """k_graphconv_fwd_blk (a graph's rows staged in shared memory by one bulk copy, gathers from shared memory, M = 64
    tiles) on a block-diagonal graph with ragged blocks -- an empty graph, one row, sizes that are not multiples of the
    tile -- against the CUDA-core kernel and the definition; forward values, the aggregation handed to the backward (via
    the gradients)."""
    from gcm import _cabi, sparse_ops

    dev = torch.device("cuda:0")
    lib = _cabi.lib()
    gen = torch.Generator().manual_seed(3 * F + H)
    sizes = torch.tensor([200, 0, 1, 64, 65, 256, 130, 17, 255, 128] * 4)
    node_off = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(sizes, 0)])
    n = int(node_off[-1])
    gid = torch.repeat_interleave(torch.arange(sizes.numel()), sizes)
    local = torch.arange(n) - node_off[gid]
    deg = torch.minimum(torch.randint(0, 45, (n,), generator=gen), local)          # at most `local` distinct earlier rows
    sink = torch.repeat_interleave(torch.arange(n), deg)
    src = node_off[gid[sink]] + (torch.rand(sink.numel(), generator=gen) * local[sink].float()).long()   # same graph, earlier
    order = torch.argsort(sink * n + src)
    sink, src = sink[order], src[order]
    csr = sparse_ops.Csr.from_sorted_edges(sink.to(dev), src.to(dev), n)
    csr.node_off, csr.max_nodes = node_off.to(dev), int(sizes.max())
    x = torch.randn(n, F, generator=gen).to(dev)
    w_rel = (torch.randn(H, F, generator=gen) / F ** 0.5).to(dev)
    w_root = (torch.randn(H, F, generator=gen) / F ** 0.5).to(dev)
    bias = torch.randn(H, generator=gen).to(dev)
    mask = (torch.rand(sink.numel(), generator=gen) < 0.7).float().to(dev) if masked else None
    outs, grads = {}, {}
    try:
        for which in (_cabi.GC_CUDA_CORES, _cabi.GC_TC):
            lib.gcm_set_graphconv_kernel(which)
            with torch.no_grad():
                outs[which] = sparse_ops.graph_conv_csr(x, csr, None, w_rel, bias, w_root, act, edge_mask=mask)
            want = "k_graphconv_fwd_blk" if which == _cabi.GC_TC else "k_graphconv_fwd"
            assert lib.gcm_last_kernel().decode() == want
            if mask is None:
                xg = x.clone().requires_grad_(True)
                wr, wo, bb = (t.clone().requires_grad_(True) for t in (w_rel, w_root, bias))
                out = sparse_ops.graph_conv_csr(xg, csr, None, wr, bb, wo, act)
                (out * torch.linspace(-1, 1, out.numel(), device=dev).view_as(out)).sum().backward()
                grads[which] = (xg.grad, wr.grad, wo.grad, bb.grad)
    finally:
        lib.gcm_set_graphconv_kernel(_cabi.GC_AUTO)
    w = torch.ones(sink.numel(), device=dev) if mask is None else mask
    agg = torch.zeros(n, F, device=dev, dtype=torch.float64).index_add_(0, sink.to(dev), (x[src.to(dev)] * w[:, None]).double())
    ref = agg @ w_rel.double().t() + bias.double() + x.double() @ w_root.double().t()
    ref = {"tanh": torch.tanh, "relu": torch.relu, "none": lambda t: t}[act](ref)
    a, b = outs[_cabi.GC_CUDA_CORES], outs[_cabi.GC_TC]
    scale = max(1.0, float(ref.abs().max()))
    assert float((b.double() - ref).abs().max()) / scale < 2e-5
    assert float((a - b).abs().max()) / scale < 1e-5
    if mask is None:
        for ga, gb in zip(grads[_cabi.GC_CUDA_CORES], grads[_cabi.GC_TC]):
            assert float((ga - gb).abs().max()) / max(1.0, float(ga.abs().max())) < 1e-4


@pytest.mark.parametrize("F,H,act,masked", [(64, 64, "tanh", False), (32, 48, "relu", False), (64, 32, "none", True)])
def test_graphconv_two_pass_on_ragged_blocks_matches_the_others(F, H, act, masked):
    """The two-pass forward (k_csr_gather + the streaming product kernel) and the backward (shared-memory transposition,
    transposed gather, paired weight-gradient reduction) on a block-diagonal graph of >= 8192 rows with ragged blocks --
    empty, one row, sizes that are not multiples of any tile -- and sources before AND after the sink, against the
    CUDA-core kernel and the float64 definition: forward values and gradients."""
    from gcm import _cabi, sparse_ops

    dev = torch.device("cuda:0")
    lib = _cabi.lib()
    gen = torch.Generator().manual_seed(5 * F + H)
    sizes = torch.tensor([1000, 0, 1, 513, 700, 384, 385, 2049, 1500, 900, 1300, 256, 767])
    node_off = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(sizes, 0)])
    n = int(node_off[-1])
    assert n >= 8192
    gid = torch.repeat_interleave(torch.arange(sizes.numel()), sizes)
    deg = torch.randint(0, 45, (n,), generator=gen)
    deg[node_off[3] + 7] = 200                                   # several 32-edge chunks in one row
    deg[sizes[gid] == 1] = 0
    sink = torch.repeat_interleave(torch.arange(n), deg)
    src = node_off[gid[sink]] + (torch.rand(sink.numel(), generator=gen) * sizes[gid[sink]].float()).long().clamp(max=10 ** 9)
    src = torch.minimum(src, node_off[gid[sink] + 1] - 1)
    key = torch.unique(sink * n + src)                           # sorted by (sink, source), duplicates merged
    sink, src = key // n, key % n
    csr = sparse_ops.Csr.from_sorted_edges(sink.to(dev), src.to(dev), n)
    csr.node_off, csr.max_nodes = node_off.to(dev), int(sizes.max())
    x = torch.randn(n, F, generator=gen).to(dev)
    w_rel = (torch.randn(H, F, generator=gen) / F ** 0.5).to(dev)
    w_root = (torch.randn(H, F, generator=gen) / F ** 0.5).to(dev)
    bias = torch.randn(H, generator=gen).to(dev)
    mask = (torch.rand(sink.numel(), generator=gen) < 0.7).float().to(dev) if masked else None
    outs, grads = {}, {}
    try:
        for which in (_cabi.GC_CUDA_CORES, _cabi.GC_TC):
            lib.gcm_set_graphconv_kernel(which)
            with torch.no_grad():
                outs[which] = sparse_ops.graph_conv_csr(x, csr, None, w_rel, bias, w_root, act, edge_mask=mask)
            if mask is None:
                xg = x.clone().requires_grad_(True)
                wr, wo, bb = (t.clone().requires_grad_(True) for t in (w_rel, w_root, bias))
                out = sparse_ops.graph_conv_csr(xg, csr, None, wr, bb, wo, act)
                (out * torch.linspace(-1, 1, out.numel(), device=dev).view_as(out)).sum().backward()
                grads[which] = (xg.grad, wr.grad, wo.grad, bb.grad)
    finally:
        lib.gcm_set_graphconv_kernel(_cabi.GC_AUTO)
    w = torch.ones(sink.numel(), device=dev) if mask is None else mask
    agg = torch.zeros(n, F, device=dev, dtype=torch.float64).index_add_(0, sink.to(dev), (x[src.to(dev)] * w[:, None]).double())
    ref = agg @ w_rel.double().t() + bias.double() + x.double() @ w_root.double().t()
    ref = {"tanh": torch.tanh, "relu": torch.relu, "none": lambda t: t}[act](ref)
    a, b = outs[_cabi.GC_CUDA_CORES], outs[_cabi.GC_TC]
    scale = max(1.0, float(ref.abs().max()))
    assert float((b.double() - ref).abs().max()) / scale < 2e-5
    assert float((a - b).abs().max()) / scale < 1e-5
    if mask is None:
        # dL/dx against the definition: d_x = (dz W_root) + A^T (dz W_rel), in float64
        xd = x.double().requires_grad_(True)
        aggd = torch.zeros(n, F, device=dev, dtype=torch.float64).index_add(0, sink.to(dev), xd[src.to(dev)])
        outd = aggd @ w_rel.double().t() + bias.double() + xd @ w_root.double().t()
        outd = {"tanh": torch.tanh, "relu": torch.relu, "none": lambda t: t}[act](outd)
        (outd * torch.linspace(-1, 1, outd.numel(), device=dev, dtype=torch.float64).view_as(outd)).sum().backward()
        gx = grads[_cabi.GC_TC][0]
        assert float((gx.double() - xd.grad).abs().max()) / max(1.0, float(xd.grad.abs().max())) < 2e-5
        for ga, gb in zip(grads[_cabi.GC_CUDA_CORES], grads[_cabi.GC_TC]):
            assert float((ga - gb).abs().max()) / max(1.0, float(ga.abs().max())) < 1e-4


def test_finite_check_kernel_finds_nan_and_inf_anywhere():
    """gcm_any_nonfinite (the one-pass form of sparse_gcm.py:203's assert) against torch.isfinite: clean data, one NaN /
    +inf / -inf at the start, in the middle, in the unaligned tail."""
    from gcm.sparse_gcm import _all_finite

    dev = torch.device("cuda:0")
    for n in (1, 3, 4, 1027, 1 << 20, (1 << 20) + 3):
        x = torch.randn(n, device=dev)
        assert _all_finite(x)
        for pos in {0, n // 2, n - 1}:
            for bad in (float("nan"), float("inf"), float("-inf")):
                y = x.clone()
                y[pos] = bad
                assert not _all_finite(y), (n, pos, bad)
