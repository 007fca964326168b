"""GPU tier, BASELINE.json's FULL sizes.  The oracle cannot run 65536 graphs, so parity at these sizes goes through
size-independent properties of the domain:

* graphs of a batch are independent (gcm.py:274-314 index everything by b): the belief stream and the final state of
  any graph must equal what the CPU oracle computes for THAT graph alone, wherever it sits in the batch / tile / CTA;
* two graphs fed the same observations must agree bit for bit;
* weight gradients are sums over graphs: grad(full batch) = grad(first half) + grad(second half);
* the coalesced edge list of the sparse path is strictly sorted by (batch, sink, source), causal, and is the CSR
  the GraphConv kernels consume.
"""
import pytest
import torch

import gcm_oracle as oracle
from helpers import make_dense_gnn, make_selector, make_sparse_gnn, make_sparse_selector, named_grads, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _probe(B):
    """graphs spread over warps / tiles / CTAs of every kernel mapping, first and last included"""
    return sorted({0, 1, 31, 32, 127, 128, B // 3, B // 2 - 1, B // 2, B - 129, B - 2, B - 1})


def test_cfg2_full_size_rollout_matches_oracle_on_probe_graphs():
    """cfg2: DenseGCM N=128 F=H=32 TemporalBackedge([1,2,4]), 65536 graphs, 140 steps (12 of them wrapping; the row-cache
    kernel from step 5 on).  Probe graphs against the oracle every step; graph i + B/2 is fed graph i's observations."""
    from gcm import _cabi
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    B, N, F, H, T = 65536, 128, 32, 32, 140
    spec = [("temporal", (1, 2, 4), "forward")]
    p = oracle.make_params(F, H)
    gnn, _ = make_dense_gnn(F, H, p, ("tanh", "tanh"))
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
    probe = _probe(B)
    gen = torch.Generator(device=dev).manual_seed(1002)
    hidden, o_hidden = None, None
    kernels = set()
    with torch.no_grad():
        for t in range(T):
            obs = torch.randn(B, F, device=dev, generator=gen)
            obs[B // 2:B // 2 + 1000] = obs[:1000]                      # duplicated graphs
            belief, hidden = mod(obs, hidden)
            kernels.add(_cabi.lib().gcm_last_kernel().decode())
            ref, o_hidden = oracle.dense_gcm_step(obs[probe].cpu(), o_hidden, spec, p, graph_size=N)
            assert rel_err(belief[probe], ref) < TOL, t
            assert torch.equal(belief[B // 2:B // 2 + 1000], belief[:1000]), t
    assert "k_step_temporal_hc" in kernels
    nodes, adj, _, num_nodes = hidden
    idx = torch.tensor(probe, device=dev)
    assert torch.equal(nodes[idx].cpu(), o_hidden[0]) and torch.equal(adj[idx].cpu(), o_hidden[1])
    assert torch.equal(num_nodes[idx].cpu(), o_hidden[3]) and int(num_nodes.min()) == N


def test_cfg2_full_size_bptt_window_gradients_of_probe_graphs_match_fp64_oracle():
    """cfg2 chain at full size (65536 graphs, N=128, F=H=32, hops 1/2/4): the graphs are filled by a no-grad sequence
    call (136 steps, wrapping), then a BPTT window of T=64 is recorded through forward_sequence (multi-step cached-row
    kernel forward, window-level backward of gcm.temporal).  The loss only looks at probe graphs, so the full-batch
    weight gradients and dL/dobs must equal the fp64 oracle's autograd on those graphs alone (1e-5 relative, budgeted
    against the fp32 oracle's own distance from fp64)."""
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    B, N, F, H, T, T0 = 65536, 128, 32, 32, 64, 136
    spec = [("temporal", (1, 2, 4), "forward")]
    acts = ("tanh", "tanh")
    p = oracle.make_params(F, H)
    gen = torch.Generator(device=dev).manual_seed(1005)
    fill = torch.randn(T0, B, F, device=dev, generator=gen) * 0.5
    obs = (torch.randn(T, B, F, device=dev, generator=gen) * 0.5).requires_grad_(True)
    probe = _probe(B)
    w = torch.zeros(T, B, H, device=dev)
    w[:, probe] = torch.randn(T, len(probe), H, device=dev, generator=gen)
    gnn, convs = make_dense_gnn(F, H, p, acts)
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
    mod.bptt_capacity = T
    with torch.no_grad():
        _, hidden = mod.forward_sequence(fill, None, time_major=True)
    beliefs, hidden = mod.forward_sequence(obs, hidden.detach(), time_major=True)
    assert getattr(hidden.token, "_gcm_tw", False)
    (beliefs * w).sum().backward()
    res = {}
    for dt in (torch.float64, torch.float32):
        pp = {k: v.to(dt).clone().requires_grad_(True) for k, v in p.items()}
        with torch.no_grad():
            _, hid = oracle.dense_gcm_rollout(fill[:, probe].to(dt).cpu(), None, spec, {k: v.detach() for k, v in pp.items()},
                                              acts, graph_size=N)
        o = obs[:, probe].detach().to(dt).cpu().requires_grad_(True)
        ref, _ = oracle.dense_gcm_rollout(o, hid, spec, pp, acts, graph_size=N)
        (ref * w[:, probe].to(dt).cpu()).sum().backward()
        res[dt] = (ref.detach(), o.grad, {k: v.grad for k, v in pp.items()})
    r64, r32 = res[torch.float64], res[torch.float32]
    assert rel_err(beliefs[:, probe], r64[0]) < TOL + rel_err(r32[0], r64[0])
    assert rel_err(obs.grad[:, probe], r64[1]) < TOL + rel_err(r32[1], r64[1])
    got = named_grads(convs)
    for k in got:
        assert rel_err(got[k], r64[2][k]) < TOL + rel_err(r32[2][k], r64[2][k]), (k, rel_err(got[k], r64[2][k]))
        assert float(got[k].abs().max()) > 0
    rest = torch.ones(B, dtype=torch.bool, device=dev)
    rest[probe] = False
    assert float(obs.grad[:, rest].abs().max()) == 0.0


def test_cfg3_full_size_bptt_window_probe_graphs_and_gradient_additivity():
    """cfg3: DenseEdge N=256 F=H=128, 16384 graphs, BPTT over T=64 from a state pre-filled with 192 nodes, bfloat16
    per-node cache (2e-2): beliefs of probe graphs against the fp64 oracle, and the six weight gradients of the full
    batch against the sum of the two half batches (the reductions run over 1 M / 5.2 M rows on the tensor cores)."""
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    B, N, F, H, T = 16384, 256, 128, 128, 64
    spec = [("dense",)]
    acts = ("tanh", "tanh")
    p = oracle.make_params(F, H)
    gen = torch.Generator(device=dev).manual_seed(1003)
    obs = 0.5 * torch.randn(T, B, F, device=dev, generator=gen)
    nodes0 = 0.5 * torch.randn(B, N, F, device=dev, generator=gen)
    nodes0[:, N - T:] = 0
    nn0 = torch.full((B,), N - T, dtype=torch.long, device=dev)
    w = torch.randn(T, 1, H, device=dev, generator=gen) / B              # loss weights: a mean over the batch

    def window(sl):
        gnn, convs = make_dense_gnn(F, H, p, acts)
        mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
        mod.bptt_capacity = T
        mod.compute_dtype = torch.bfloat16
        n = sl.stop - sl.start
        adj0 = torch.zeros(n, N, N, device=dev)
        adj0[:, : N - T, : N - T] = 1
        hidden = (nodes0[sl], adj0, torch.zeros(0, device=dev), nn0[sl])
        outs = []
        for t in range(T):
            belief, hidden = mod(obs[t, sl], hidden)
            outs.append(belief)
        outs = torch.stack(outs)
        (outs * w).sum().backward()
        return outs.detach(), {k: v.clone() for k, v in named_grads(convs).items()}

    outs, g_full = window(slice(0, B))
    _, g_a = window(slice(0, B // 2))
    _, g_b = window(slice(B // 2, B))
    for k in g_full:
        assert rel_err(g_a[k] + g_b[k], g_full[k]) < 1e-3 * max(1.0, float(g_full[k].abs().max())) + 1e-6, k
        assert float(g_full[k].abs().max()) > 0
    probe = [0, 127, 128, B // 2, B - 1]
    adj0 = torch.zeros(len(probe), N, N, dtype=torch.float64)
    adj0[:, : N - T, : N - T] = 1
    ref, _ = oracle.dense_gcm_rollout(obs[:, probe].double().cpu(), (nodes0[probe].double().cpu(), adj0,
                                                                    torch.zeros(0, dtype=torch.float64), nn0[probe].cpu()),
                                      spec, {k: v.double() for k, v in p.items()}, acts, graph_size=N)
    assert rel_err(outs[:, probe], ref) < 2e-2


def test_cfg3_full_size_bf16_gradients_of_probe_graphs_match_fp64_oracle():
    """cfg3 at full size (16384 graphs, N=256, F=H=128, T=64, bfloat16 per-node cache): the loss only looks at the beliefs
    of a few probe graphs, so the six weight gradients and dL/dobs of the FULL-batch run (every kernel of the window at
    its full grid; the other graphs contribute exact zeros) must equal the fp64 oracle's autograd on those graphs alone
    -- within BASELINE.json's 2e-2 for the bf16 configuration."""
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    B, N, F, H, T = 16384, 256, 128, 128, 64
    spec = [("dense",)]
    acts = ("tanh", "tanh")
    p = oracle.make_params(F, H)
    gen = torch.Generator(device=dev).manual_seed(1004)
    obs = (0.5 * torch.randn(T, B, F, device=dev, generator=gen)).requires_grad_(True)
    nodes0 = 0.5 * torch.randn(B, N, F, device=dev, generator=gen)
    nodes0[:, N - T:] = 0
    nn0 = torch.full((B,), N - T, dtype=torch.long, device=dev)
    probe = [0, 127, 128, B // 2, B - 1]
    w = torch.zeros(T, B, H, device=dev)
    w[:, probe] = torch.randn(T, len(probe), H, device=dev, generator=gen)
    gnn, convs = make_dense_gnn(F, H, p, acts)
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
    mod.bptt_capacity = T
    mod.compute_dtype = torch.bfloat16
    adj0 = torch.zeros(B, N, N, device=dev)
    adj0[:, : N - T, : N - T] = 1
    hidden = (nodes0, adj0, torch.zeros(0, device=dev), nn0)
    outs = []
    for t in range(T):
        belief, hidden = mod(obs[t], hidden)
        outs.append(belief)
    (torch.stack(outs) * w).sum().backward()
    # fp64 oracle on the probe graphs alone
    o = obs[:, probe].detach().double().cpu().requires_grad_(True)
    pp = {k: v.double().clone().requires_grad_(True) for k, v in p.items()}
    a0 = torch.zeros(len(probe), N, N, dtype=torch.float64)
    a0[:, : N - T, : N - T] = 1
    ref, _ = oracle.dense_gcm_rollout(o, (nodes0[probe].double().cpu(), a0, torch.zeros(0, dtype=torch.float64),
                                          nn0[probe].cpu()), spec, pp, acts, graph_size=N)
    (ref * w[:, probe].double().cpu()).sum().backward()
    got = named_grads(convs)
    for k in got:
        assert rel_err(got[k], pp[k].grad) < 2e-2, (k, rel_err(got[k], pp[k].grad))
        assert float(got[k].abs().max()) > 0
    assert rel_err(obs.grad[:, probe], o.grad) < 2e-2
    rest = torch.ones(B, dtype=torch.bool, device=dev)
    rest[probe] = False
    assert float(obs.grad[:, rest].abs().max()) == 0.0          # graphs are independent: no gradient leaks across them


def test_cfg4_full_size_cosine_rollout_matches_oracle_on_probe_graphs():
    """cfg4: CosineEdge(0.5) N=512 F=H=64, 4096 graphs, clustered observations (wide margin around the threshold), 530
    steps (18 of them evicting): beliefs every 10th step and the final adjacency of probe graphs against the oracle."""
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    B, N, F, H, T = 4096, 512, 64, 64, 530
    spec = [("cosine", 0.5)]
    p = oracle.make_params(F, H)
    gnn, _ = make_dense_gnn(F, H, p, ("tanh", "tanh"))
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
    gen = torch.Generator().manual_seed(1004)
    centres = torch.randn(16, F, generator=gen)
    sched = torch.randint(0, 16, (T,), generator=gen)
    probe = [0, 1, 255, 256, B // 2, B - 1]
    hidden, o_hidden = None, None
    gdev = torch.Generator(device=dev).manual_seed(7)
    with torch.no_grad():
        for t in range(T):
            obs = centres[sched[t]].to(dev) + 0.05 * torch.randn(B, F, device=dev, generator=gdev)
            belief, hidden = mod(obs, hidden)
            ref, o_hidden = oracle.dense_gcm_step(obs[probe].cpu(), o_hidden, spec, p, graph_size=N)
            if t % 10 == 0 or t >= N - 2:
                assert rel_err(belief[probe], ref) < 5 * TOL, t
    nodes, adj, _, num_nodes = hidden
    idx = torch.tensor(probe, device=dev)
    assert torch.equal(adj[idx].cpu(), o_hidden[1]) and torch.equal(nodes[idx].cpu(), o_hidden[0])
    assert torch.equal(num_nodes[idx].cpu(), o_hidden[3])


def test_cfg5_full_size_sparse_forward_edges_and_probe_graphs():
    """cfg5: SparseGCM TemporalEdge([1]) + SpatialRadiusEdge(0.25), 1024 graphs x 4096 nodes all at once.  The edge
    list is strictly sorted by (batch, sink, source) and causal; graphs 0 and B-1 (edges and outputs) against the
    oracle run on those two graphs alone."""
    from gcm.sparse_gcm import SparseGCM

    dev = torch.device("cuda:0")
    B, N, F, H = 1024, 4096, 64, 64
    p = oracle.make_params(F, H)
    spec, aux = [("temporal", (1,))], [("spatial_radius", slice(0, 2), 0.25)]
    gnn, _ = make_sparse_gnn(F, H, p, ("tanh", "tanh"))
    mod = SparseGCM(gnn.to(dev), edge_selectors=make_sparse_selector(spec), aux_edge_selectors=make_sparse_selector(aux),
                    graph_size=N)
    gen = torch.Generator(device=dev).manual_seed(1005)
    x = torch.randn(B, N, F, device=dev, generator=gen)
    x[..., 0:2] = torch.cumsum(0.1 * torch.randn(B, N, 2, device=dev, generator=gen), dim=1)
    taus = torch.full((B,), N, dtype=torch.long, device=dev)
    with torch.no_grad():
        mx, (nodes, adj, Tn) = mod(x, taus, None)
    e = adj.coalesce().indices()
    key = (e[0] * N + e[1]) * N + e[2]
    assert bool((key[1:] > key[:-1]).all())                              # sorted, no duplicates
    assert bool((e[2] < e[1]).all()) and int(e[0].max()) == B - 1       # causal
    assert torch.equal(Tn, taus) and torch.equal(nodes, x)
    for b in (0, B - 1):
        omx, (on, oe, oT) = oracle.sparse_gcm_forward(x[b:b + 1].cpu(), taus[:1].cpu(), None, spec, p, graph_size=N,
                                                     aux_selectors=aux)
        eb = e[:, e[0] == b].cpu()
        eb[0] = 0
        assert torch.equal(eb, oe), b
        assert rel_err(mx[b:b + 1], omx) < 5 * TOL, b
