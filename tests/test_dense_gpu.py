"""GPU tier: the fused CUDA path (through the C ABI) against the golden vectors of the reference and
against the oracle on seeded inputs.  Tolerances are BASELINE.json's: adjacency / node slots /
num_nodes bit-exact, beliefs within 1e-5 relative (fp32)."""
import pytest
import torch

import gcm_oracle as oracle
from helpers import (dense_cases, load_golden, make_dense_gnn, make_preprocessor, make_selector, preproc_cases,
                     rel_err)

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _run_case(g, style="readme", keep_tuple_every=None):
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    gnn, _ = make_dense_gnn(g["F"], g["H"], g["params"], g["acts"], style)
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(g["spec"]), graph_size=g["N"])
    if mod.edge_selectors is not None:
        mod.edge_selectors.to(dev)
    assert mod.fused_plan() is not None, "configuration should take the fused path"
    hidden = None if g["init"] is None else tuple(t.to(dev) for t in g["init"])
    with torch.no_grad():
        for t in range(g["T"]):
            belief, hidden = mod(g["obs"][t].to(dev), hidden)
            assert rel_err(belief, g["beliefs"][t]) < TOL, (g["name"], t)
            if t in g["snaps"]:
                for a, b in zip(hidden, g["snaps"][t]):
                    assert torch.equal(a.cpu().float(), b.float()), (g["name"], t)
            if keep_tuple_every and t % keep_tuple_every == 0:
                hidden = tuple(hidden)          # force the ingest path (what RayDenseGCM does)
    assert mod.fused_plan() is not None and mod.fused_plan().validated
    nodes, adj, weights, num_nodes = hidden
    assert torch.equal(nodes.cpu(), g["final"][0])
    assert torch.equal(adj.cpu(), g["final"][1].float())
    assert torch.equal(num_nodes.cpu(), g["final"][3])
    assert num_nodes.dtype == torch.long and adj.dtype == torch.float32


@pytest.mark.parametrize("name", dense_cases())
def test_fused_matches_reference_golden(name):
    _run_case(load_golden(name))


@pytest.mark.parametrize("name", ["dense_temporal124_fwd", "dense_denseedge_wrap", "dense_temporal13_both",
                                  "dense_cosine"])
def test_fused_sequential_gnn_and_reingest(name):
    """Same cases with a gcm.nn.Sequential GNN and the hidden state round-tripped through plain
    tensors every 3 steps (materialize -> ingest)."""
    _run_case(load_golden(name), style="sequential", keep_tuple_every=3)


SEEDED = [
    # B, N, F, H, T, spec
    (33, 128, 32, 32, 140, [("temporal", (1, 2, 4), "forward")]),      # BASELINE cfg 2 shape, wraps
    (16, 128, 8, 32, 20, [("temporal", (1,), "forward")]),             # BASELINE cfg 1 (README)
    (5, 40, 16, 32, 50, [("temporal", (1, 3), "both")]),
    (4, 33, 20, 24, 70, [("temporal", (2, 5), "backward"), ("temporal", (1,), "forward")]),
    (3, 64, 48, 40, 70, [("dense",)]),
    # DenseEdge chained with other selectors is still the all-ones block: the DenseEdge kernels serve it
    (3, 24, 16, 32, 40, [("temporal", (1, 2), "forward"), ("dense",)]),
    (3, 20, 16, 16, 30, [("dense",), ("temporal", (1, 3), "both"), ("cosine", 0.5)]),
    (2, 256, 128, 128, 12, [("dense",)]),                              # BASELINE cfg 3 shape (fp32 path)
    (6, 96, 64, 64, 100, [("cosine", 0.5)]),                           # BASELINE cfg 4 shape, reduced N
    (6, 96, 64, 64, 100, [("euclidean", 2.0)]),
    # B >= 128: the cross-batch mean distance runs on the tensor cores (k_euclid_tc, gcm/fused.py: euclid_batchmean);
    # edge set -> belief end to end through DenseGCM, window wraps, threshold margin asserted below
    (256, 24, 32, 32, 40, [("euclidean", 2.0)]),
    (384, 20, 64, 64, 30, [("euclidean", 2.0)]),
    (4, 70, 12, 16, 80, [("spatial", 1.0, slice(0, 2), None)]),
    (3, 1024, 8, 8, 5, [("temporal", (1, 512), "forward")]),           # maximum graph_size
    (2, 20, 256, 256, 4, [("dense",)]),                                # maximum feature width
    (3, 12, 5, 7, 30, []),                                             # no selector at all
]


def _clustered(gen, T, B, F, K=8, noise=0.03):
    centres = torch.randn(K, F, generator=gen) * 1.5
    sched = torch.randint(0, K, (T,), generator=gen)
    return (centres[sched].unsqueeze(1).expand(T, B, F) + noise * torch.randn(T, B, F, generator=gen)).contiguous()


@pytest.mark.parametrize("B,N,F,H,T,spec", SEEDED)
def test_fused_matches_oracle_seeded(B, N, F, H, T, spec):
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(1234 + N + F)
    distance = bool(spec) and spec[0][0] in ("cosine", "euclidean", "spatial")
    obs = _clustered(gen, T, B, F) if distance else torch.randn(T, B, F, generator=gen)
    p = oracle.make_params(F, H)
    gnn, _ = make_dense_gnn(F, H, p, ("tanh", "tanh"))
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
    hidden, o_hidden, o64 = None, None, None
    p64 = {k: v.double() for k, v in p.items()}
    tc_euclid = bool(spec) and spec[0][0] == "euclidean" and B >= 128
    seen = set()
    if tc_euclid:
        # spy on the library call that computes the distances: the tensor-core entry must be the one that runs
        from gcm import _cabi, fused
        assert fused.EUCLID_TC and F % 16 == 0 and F <= 64
        real = _cabi.lib().gcm_euclid_batchmean_tc

        def spy(*a):
            seen.add("tc")
            return real(*a)
        _cabi.lib().gcm_euclid_batchmean_tc = spy
    with torch.no_grad():
        for t in range(T):
            belief, hidden = mod(obs[t].to(dev), hidden)
            ref, o_hidden = oracle.dense_gcm_step(obs[t], o_hidden, spec, p, graph_size=N)
            ref64, o64 = oracle.dense_gcm_step(obs[t].double(), o64, spec, p64, graph_size=N)
            # 1e-5 relative to the reference, budgeted against fp64: the fp32
            # reference's own rounding distance from the exact result is not charged to the kernel
            assert rel_err(belief, ref64) < TOL + rel_err(ref, ref64), (t, rel_err(belief, ref))
    if tc_euclid:
        _cabi.lib().gcm_euclid_batchmean_tc = real
        assert seen == {"tc"}, "the tensor-core distance kernel should have served this batch size"
    if len(spec) > 1 and any(s[0] == "dense" for s in spec):
        assert mod.fused_plan().ones and hidden.claim().dense_ok and hidden.claim().masks_stale, \
            "a chain that contains DenseEdge should run on the DenseEdge (implicit all-ones) kernels"
    if distance:
        # the edge set depends on a float comparison: the inputs must keep a margin
        kind = spec[0][0]
        nodes_o, _, _, nn_o = o_hidden
        _, d = oracle.distance_edges(nodes_o, torch.zeros(B, N, N), (nn_o - 1).clamp(min=0), kind, spec[0][1],
                                     *(spec[0][2:4] if kind == "spatial" else ()), return_dists=True)
        assert float((d - spec[0][1]).abs().min()) > 1e-3
    nodes, adj, weights, num_nodes = hidden
    assert torch.equal(nodes.cpu(), o_hidden[0])
    assert torch.equal(adj.cpu(), o_hidden[1])
    assert torch.equal(num_nodes.cpu(), o_hidden[3])


def test_selectors_standalone_dense_forward():
    """The selectors' own forward(nodes, adj_mats, edge_weights, num_nodes, B) on dense tensors."""
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(5)
    B, N, F = 7, 19, 6
    nodes = _clustered(gen, N, B, F).permute(1, 0, 2).contiguous()
    num_nodes = torch.randint(0, N, (B,), generator=gen)
    base = (torch.rand(B, N, N, generator=gen) < 0.1).float()
    for spec in ([("temporal", (1, 3), "both")], [("temporal", (2,), "backward")], [("dense",)],
                 [("cosine", 0.4)], [("euclidean", 2.5)], [("spatial", 0.8, slice(1, 4), slice(0, 3))]):
        want = oracle.apply_selectors(nodes, base.clone(), num_nodes, spec)
        sel = make_selector(spec).to(dev)
        adj = base.clone().to(dev)
        got, w = sel(nodes.to(dev), adj, torch.zeros(0, device=dev), num_nodes.to(dev), B)
        assert got.data_ptr() == adj.data_ptr()                       # in place, like the reference
        assert torch.equal(got.cpu(), want), spec


def test_stale_hidden_raises_and_snapshot_is_stable():
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    p = oracle.make_params(4, 8)
    gnn, _ = make_dense_gnn(4, 8, p, ("tanh", "tanh"))
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector([("temporal", (1,), "forward")]), graph_size=6)
    x = torch.randn(3, 4, device=dev)
    with torch.no_grad():
        _, h1 = mod(x, None)
        pinned = tuple(h1)
        nodes_before = pinned[0].clone()
        _, h2 = mod(x + 1, h1)
        assert torch.equal(pinned[0], nodes_before)       # materialised tensors never change
        assert tuple(h1)[0] is pinned[0]                  # cached snapshot stays readable
        _, h3 = mod(x, h2)
        with pytest.raises(RuntimeError, match="stale"):
            tuple(h2)
        with pytest.raises(RuntimeError, match="stale"):
            mod(x, h2)


def test_nan_flag_raises_reference_message():
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    p = oracle.make_params(4, 8)
    gnn, _ = make_dense_gnn(4, 8, p, ("relu", "none"))
    mod = DenseGCM(gnn.to(dev), graph_size=6)
    x = torch.full((2, 4), float("nan"), device=dev)
    with torch.no_grad():
        _, h = mod(x, None)
    with pytest.raises(AssertionError, match="Got NaN in returned memory"):
        tuple(h)


@pytest.mark.parametrize("variant", ["hc", "tc", "win", "rows"])
@pytest.mark.parametrize("B,N,F,T,hops", [(70, 128, 32, 135, (1, 2, 4)), (33, 24, 8, 30, (1,)), (200, 16, 16, 20, (1, 3)),
                                          (130, 8, 32, 20, (1, 4))])
def test_every_temporal_kernel_variant_matches_oracle(variant, B, N, F, T, hops):
    """The pure-temporal step has four kernels (include/gcm_b200.h: gcm_temporal_kernel); each one is forced
    in turn and must reproduce the oracle, including the wrap (T > N) and batches that are not tile multiples.
    The last shape has N - 1 < 2 max_hop, where cached layer-1 rows would be wrong: hc must decline."""
    from gcm import _cabi
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    lib = _cabi.lib()
    which = {"hc": _cabi.TK_HC, "tc": _cabi.TK_TC, "win": _cabi.TK_WIN, "rows": _cabi.TK_ROWS}[variant]
    expect = {"hc": "k_step_temporal_hc" if N - 1 >= 2 * max(hops) else "k_step_temporal_tc",
              "tc": "k_step_temporal_tc", "win": "k_step_temporal_win", "rows": "k_step_temporal"}[variant]
    spec = [("temporal", hops, "forward")]
    gen = torch.Generator().manual_seed(99 + N + F)
    obs = torch.randn(T, B, F, generator=gen)
    p = oracle.make_params(F, 32)
    gnn, _ = make_dense_gnn(F, 32, p, ("tanh", "tanh"))
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
    _cabi.check(lib.gcm_set_temporal_kernel(which), "gcm_set_temporal_kernel")
    try:
        hidden, o_hidden, o64 = None, None, None
        p64 = {k: v.double() for k, v in p.items()}
        with torch.no_grad():
            for t in range(T):
                belief, hidden = mod(obs[t].to(dev), hidden)
                assert t == 0 or lib.gcm_last_kernel().decode() == expect   # step 0 ends with the plan validation
                ref, o_hidden = oracle.dense_gcm_step(obs[t], o_hidden, spec, p, graph_size=N)
                ref64, o64 = oracle.dense_gcm_step(obs[t].double(), o64, spec, p64, graph_size=N)
                assert rel_err(belief, ref64) < TOL + rel_err(ref, ref64), (t, rel_err(belief, ref))
        nodes, adj, weights, num_nodes = hidden
        assert torch.equal(nodes.cpu(), o_hidden[0])
        assert torch.equal(adj.cpu(), o_hidden[1])
        assert torch.equal(num_nodes.cpu(), o_hidden[3])
    finally:
        lib.gcm_set_temporal_kernel(_cabi.TK_AUTO)


def test_row_cache_follows_weight_updates_and_reingest():
    """The layer-1 row cache is only read while every row it would use was written under the current weights:
    after an in-place weight update the recomputing kernel runs for max_hop steps, then the cached-row kernel
    resumes; a state that went through materialise -> ingest (what RayDenseGCM does on every call) is recognised as this
    chain's own adjacency and keeps the fast kernels (cache refilled first); an ingested adjacency that is NOT the chain's
    pattern takes the general kernel.  Every step is checked against the oracle evaluated with the weights of that
    step."""
    from gcm import _cabi
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    lib = _cabi.lib()
    B, N, F, T, hops = 45, 32, 32, 60, (1, 2, 4)
    spec = [("temporal", hops, "forward")]
    gen = torch.Generator().manual_seed(4242)
    obs = torch.randn(T, B, F, generator=gen)
    p = oracle.make_params(F, 32)
    gnn, convs = make_dense_gnn(F, 32, p, ("tanh", "tanh"))
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
    hidden, o_hidden = None, None
    names = []
    with torch.no_grad():
        for t in range(T):
            if t in (10, 40):       # "optimizer step": every weight changes in place
                for k in p:
                    p[k] = p[k] + 0.05 * torch.randn(p[k].shape, generator=gen)
                convs[0].lin_rel.weight.copy_(p["w_rel1"]); convs[0].lin_rel.bias.copy_(p["b1"])
                convs[0].lin_root.weight.copy_(p["w_root1"])
                convs[1].lin_rel.weight.copy_(p["w_rel2"]); convs[1].lin_rel.bias.copy_(p["b2"])
                convs[1].lin_root.weight.copy_(p["w_root2"])
            if t == 50:
                hidden = tuple(hidden)
            belief, hidden = mod(obs[t].to(dev), hidden)
            names.append(lib.gcm_last_kernel().decode() if t else "k_step_temporal_hc")   # step 0 ends with the plan validation
            ref, o_hidden = oracle.dense_gcm_step(obs[t], o_hidden, spec, p, graph_size=N)
            assert rel_err(belief, ref) < 2e-5, (t, names[-1], rel_err(belief, ref))
    assert names[:10] == ["k_step_temporal_hc"] * 10
    assert names[10:14] == ["k_step_temporal_tc"] * 4 and names[14] == "k_step_temporal_hc"
    assert names[40:44] == ["k_step_temporal_tc"] * 4 and names[44] == "k_step_temporal_hc"
    assert names[50:54] == ["k_step_temporal_tc"] * 4 and all(n == "k_step_temporal_hc" for n in names[54:])
    nodes, adj, weights, num_nodes = hidden
    assert torch.equal(nodes.cpu(), o_hidden[0]) and torch.equal(adj.cpu(), o_hidden[1])
    # one extra edge that the chain would not have written: no longer the chain's pattern -> the general kernel
    adj2 = adj.clone()
    adj2[:, 5, 0] = 1.0
    o_hidden = (o_hidden[0], adj2.cpu(), o_hidden[2], o_hidden[3])
    with torch.no_grad():
        belief, hidden = mod(obs[0].to(dev), (nodes, adj2, weights, num_nodes))
        ref, o_hidden = oracle.dense_gcm_step(obs[0], o_hidden, spec, p, graph_size=N)
    assert lib.gcm_last_kernel().decode() == "k_step_general" and rel_err(belief, ref) < 2e-5
    assert torch.equal(tuple(hidden)[1].cpu(), o_hidden[1])


@pytest.mark.parametrize("spec", [[("cosine", 0.5)], [("spatial", 1.0, slice(0, 2), None)], [("euclidean", 2.0)]])
def test_distance_selector_cache_path_and_its_retirement(spec):
    """A rollout with exactly one distance selector runs on the per-node pre-activation cache
    (csrc/gcm_dense_zc.cu), including the eviction correction once the window is full; an in-place weight update
    retires the cache for that state and the general kernel takes over.  Every step against the oracle."""
    from gcm import _cabi
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    lib = _cabi.lib()
    B, N, F, H, T = 5, 40, 24, 20, 130
    gen = torch.Generator().manual_seed(2024)
    obs = _clustered(gen, T, B, F)
    p = oracle.make_params(F, H)
    gnn, convs = make_dense_gnn(F, H, p, ("tanh", "tanh"))
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
    hidden, o_hidden, o64 = None, None, None
    names = []
    with torch.no_grad():
        for t in range(T):
            if t == 100:
                p = {k: v + 0.05 * torch.randn(v.shape, generator=gen) for k, v in p.items()}
                convs[0].lin_rel.weight.copy_(p["w_rel1"]); convs[0].lin_rel.bias.copy_(p["b1"])
                convs[0].lin_root.weight.copy_(p["w_root1"])
                convs[1].lin_rel.weight.copy_(p["w_rel2"]); convs[1].lin_rel.bias.copy_(p["b2"])
                convs[1].lin_root.weight.copy_(p["w_root2"])
            belief, hidden = mod(obs[t].to(dev), hidden)
            names.append(lib.gcm_last_kernel().decode())
            ref, o_hidden = oracle.dense_gcm_step(obs[t], o_hidden, spec, p, graph_size=N)
            ref64, o64 = oracle.dense_gcm_step(obs[t].double(), o64, spec, {k: v.double() for k, v in p.items()},
                                               graph_size=N)
            assert rel_err(belief, ref64) < TOL + rel_err(ref, ref64), (t, names[-1], rel_err(belief, ref))
    assert all(n == "k_step_dist_zc" for n in names[1:100])
    assert all(n == "k_step_general" for n in names[100:])
    nodes, adj, weights, num_nodes = hidden
    assert torch.equal(nodes.cpu(), o_hidden[0]) and torch.equal(adj.cpu(), o_hidden[1])


@pytest.mark.parametrize("spec", [[("temporal", (1, 2, 4), "forward")], [("dense",)], [("temporal", (1,), "both")]])
def test_rowwise_preprocessor_runs_fused_and_matches_the_reference_step(spec):
    """DenseGCM(preprocessor=Linear) -- what RayDenseGCM builds (ray_gcm.py:118,133-136) -- on the fused rollout path
    (SURVEY 8(f) rank 2): the reference preprocesses all N rows every step (gcm.py:290-291), here only the new
    observation.  Beliefs every step and the caller's view of m_t (RAW nodes, adjacency, num_nodes) against the oracle
    run on preprocessed observations; window wraps; a caller-supplied tuple; an in-place update of the preprocessor
    weights in the middle (every stored row must get its new image)."""
    from gcm import _cabi
    from gcm.gcm import DenseGCM
    from gcm.state import DenseHidden

    dev = torch.device("cuda:0")
    B, N, F_raw, F, H, T = 6, 12, 10, 32, 32, 40
    gen = torch.Generator().manual_seed(99)
    p = oracle.make_params(F, H)
    gnn, _ = make_dense_gnn(F, H, p, ("tanh", "tanh"))
    pre = torch.nn.Sequential(torch.nn.Linear(F_raw, F), torch.nn.Tanh())
    mod = DenseGCM(gnn.to(dev), preprocessor=pre.to(dev), edge_selectors=make_selector(spec), graph_size=N)
    assert mod.fused_plan() is not None and mod.fused_plan().pre
    obs = torch.randn(T, B, F_raw, generator=gen)
    hidden, o_hidden, raw_nodes = None, None, torch.zeros(B, N, F_raw)
    names = []
    with torch.no_grad():
        for t in range(T):
            if t == 25:
                pre[0].weight.mul_(0.5)
                pre[0].bias.add_(0.1)
                # the oracle's stored rows are images under the OLD weights: rebuild them from the raw observations
                o_hidden = (pre(raw_nodes.to(dev)).cpu() * (torch.arange(N).view(1, N, 1) < o_hidden[3].view(B, 1, 1)),
                            o_hidden[1], o_hidden[2], o_hidden[3])
            if t == 30:
                hidden = tuple(hidden)                                  # a plain reference-layout tuple comes back in
                assert torch.equal(hidden[0].cpu(), raw_nodes)
            belief, hidden = mod(obs[t].to(dev), hidden)
            names.append(_cabi.lib().gcm_last_kernel().decode())
            assert isinstance(hidden, DenseHidden)
            y = pre(obs[t].to(dev)).cpu()
            ref, o_hidden = oracle.dense_gcm_step(y, o_hidden, spec, p, graph_size=N)
            assert rel_err(belief, ref) < 5 * TOL, (t, names[-1])
            # the reference's own bookkeeping of the raw rows (gcm.py:262-274, 323-355)
            full = int(o_hidden[3][0]) == N and t >= N
            if full:
                raw_nodes = torch.cat([raw_nodes[:, 1:], obs[t].unsqueeze(1)], dim=1)
            else:
                raw_nodes[:, t] = obs[t]
    assert names[-1] == "k_log_write"
    nodes, adj, _, num_nodes = hidden
    assert torch.equal(nodes.cpu(), raw_nodes) and torch.equal(adj.cpu(), o_hidden[1])
    assert torch.equal(num_nodes.cpu(), o_hidden[3])
    # autograd: a forward-only temporal chain records on the window-level backward (tests/test_temporal_bwd_gpu.py pins its
    # gradients to the reference), anything else takes the generic path; either way the belief carries a graph
    x = obs[0].to(dev).requires_grad_(True)
    belief, h2 = mod(x, hidden)
    fused_grad = spec[0][0] == "temporal" and spec[0][2] == "forward"
    assert isinstance(h2, DenseHidden) == fused_grad and belief.requires_grad
    belief.sum().backward()
    assert x.grad is not None and pre[0].weight.grad is not None


@pytest.mark.parametrize("B,C,F,learned", [(300, 9, 64, False), (129, 5, 32, True), (1024, 6, 16, False)])
def test_euclid_batchmean_tensor_core_kernel_matches_the_cuda_core_kernel(B, C, F, learned):
    """EuclideanEdge's cross-batch mean distance (distance.py:48-49) through tcgen05 (|n|^2 + |c|^2 - 2 n.c in 3xTF32)
    against the difference-squared CUDA-core kernel and against float64: clustered rows (small distances, where the
    matmul form cancels), a batch that is not a multiple of the 128-observation tile, the learned scale."""
    from gcm import fused
    from gcm.state import DenseState

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(B + F)
    centres = 2.0 * torch.randn(5, F, generator=gen)
    st = DenseState(B, C, F, dev)
    st.nodes.copy_((centres[torch.randint(0, 5, (B, C), generator=gen)] + 0.05 * torch.randn(B, C, F, generator=gen)).to(dev))
    cur = (centres[torch.randint(0, 5, (B,), generator=gen)] + 0.05 * torch.randn(B, F, generator=gen)).to(dev)
    param = torch.tensor([-1.7], device=dev) if learned else None
    out = {}
    try:
        for tc_on in (False, True):
            fused.EUCLID_TC = tc_on
            d = torch.empty(B, C, device=dev)
            fused.euclid_batchmean(st, cur, param, d)
            out[tc_on] = d
    finally:
        fused.EUCLID_TC = True
    ref = torch.cdist(st.nodes.double().view(1, B * C, F), cur.double().view(1, B, F), compute_mode="donot_use_mm_for_euclid_dist")
    ref = ref.mean(-1).view(B, C) / (1.7 if learned else 1.0)
    for k, d in out.items():
        assert rel_err(d, ref) < 2e-5, k


@pytest.mark.parametrize("name", preproc_cases())
def test_preprocessor_path_matches_reference_golden(name):
    """The fused preprocessor path against golden vectors of the unmodified reference DenseGCM(preprocessor=...)."""
    from gcm.gcm import DenseGCM
    from gcm.state import DenseHidden

    g = load_golden(name)
    dev = torch.device("cuda:0")
    gnn, _ = make_dense_gnn(g["F"], g["H"], g["params"], ("tanh", "tanh"))
    mod = DenseGCM(gnn.to(dev), preprocessor=make_preprocessor(g).to(dev), edge_selectors=make_selector(g["spec"]),
                   graph_size=g["N"])
    hidden = None
    with torch.no_grad():
        for t in range(g["T"]):
            belief, hidden = mod(g["obs"][t].to(dev), hidden)
            assert isinstance(hidden, DenseHidden)
            assert rel_err(belief, g["beliefs"][t]) < 5 * TOL, (name, t)
    nodes, adj, _, num_nodes = hidden
    assert torch.equal(nodes.cpu(), g["final"][0]) and torch.equal(adj.cpu(), g["final"][1].float())
    assert torch.equal(num_nodes.cpu(), g["final"][3])


def test_user_gnn_with_a_different_forward_is_not_fused():
    """A user module whose children look like the README GNN (two DenseGraphConv + one activation) but whose forward()
    uses the adjacency differently (here: transposed) must NOT be replaced by the fused two-layer formula: the one-time
    validation runs the module on a synthetic state WITH edges (the first live step of a rollout has none), the plan is
    dropped with a warning and every step goes through the module itself."""
    import warnings

    from gcm.gcm import DenseGCM
    from gcm.nn import DenseGraphConv

    dev = torch.device("cuda:0")
    F, H, N, B, T = 8, 16, 10, 3, 14

    class Odd(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gc0 = DenseGraphConv(F, H)
            self.gc1 = DenseGraphConv(H, H)
            self.act = torch.nn.Tanh()

        def forward(self, x, adj, weights, B, N):
            x = self.act(self.gc0(x, adj.transpose(1, 2)))
            return self.act(self.gc1(x, adj.transpose(1, 2)))

    torch.manual_seed(3)
    gnn = Odd().to(dev)
    mod = DenseGCM(gnn, edge_selectors=make_selector([("temporal", (1, 2), "forward")]), graph_size=N)
    assert mod.fused_plan() is not None                       # the structure matches ...
    gen = torch.Generator().manual_seed(8)
    obs = torch.randn(T, B, F, generator=gen).to(dev)
    hidden = None
    with torch.no_grad(), warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        for t in range(T):
            belief, hidden = mod(obs[t], hidden)
            # ... but the semantics are the module's: recompute with torch on the materialised state
            nodes, adj, _, nn = tuple(hidden)
            want = gnn(nodes, adj, torch.zeros(0, device=dev), B, N)[torch.arange(B, device=dev), nn - 1]
            assert rel_err(belief, want) < 1e-5, t
    assert mod._plan is None and any("does not compute one" in str(w.message) for w in caught)


@pytest.mark.parametrize("spec", [[("temporal", (1, 2, 4), "forward")], [("dense",)], [("cosine", 0.5)]])
def test_hidden_clone_is_an_independent_checkpoint(spec):
    """m_t.clone(): two rollouts restarted from the same prepared memory give the same beliefs as one rollout from the
    reference-layout tuple, and stepping one copy leaves the other (and the original) untouched."""
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    B, N, F, H, T0, T1 = 6, 12, 16, 32, 9, 8
    gen = torch.Generator().manual_seed(55)
    obs = _clustered(gen, T0 + T1, B, F).to(dev)
    p = oracle.make_params(F, H)
    gnn, _ = make_dense_gnn(F, H, p, ("tanh", "tanh"))
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
    with torch.no_grad():
        hidden = None
        for t in range(T0):
            _, hidden = mod(obs[t], hidden)
        base = tuple(t.clone() for t in hidden)               # the checkpoint in the reference's layout
        runs = []
        for _ in range(2):
            h = hidden.clone()
            outs = []
            for t in range(T0, T0 + T1):
                o, h = mod(obs[t], h)
                outs.append(o)
            runs.append((torch.stack(outs), tuple(h)))
        for a, b in zip(tuple(hidden), base):
            assert torch.equal(a, b)                          # the original never moved
        h, outs = base, []
        for t in range(T0, T0 + T1):
            o, h = mod(obs[t], h)
            outs.append(o)
        want = torch.stack(outs)
    assert torch.equal(runs[0][0], runs[1][0]) and rel_err(runs[0][0], want) < TOL
    for a, b in zip(runs[0][1], tuple(h)):
        assert torch.equal(a, b)
