"""GPU tier: window-level backward of forward-only TemporalBackedge chains (gcm.temporal, csrc/gcm_temporal_bwd.cu)
against the fp64 oracle's autograd (the reference's BPTT, tests/test_gcm.py:412-439): gradients of the six GNN weight
tensors and of every observation, per-step recording, sequence recording and mixtures, windows that wrap, truncated
BPTT over several windows with weight updates in between.  Tolerance: BASELINE.json's 1e-5 relative on top of the fp32
oracle's own distance from fp64."""
import pytest
import torch

import gcm_oracle as oracle
from helpers import make_dense_gnn, make_selector, named_grads, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _oracle_grads(obs, w, p, spec, N, hidden=None):
    res = {}
    for dt in (torch.float64, torch.float32):
        o = obs.to(dt).clone().requires_grad_(True)
        pp = {k: v.to(dt).clone().requires_grad_(True) for k, v in p.items()}
        hid = None if hidden is None else tuple(h.to(dt) if h.is_floating_point() else h for h in hidden[dt])
        outs, hid = oracle.dense_gcm_rollout(o, hid, spec, pp, ("tanh", "tanh"), graph_size=N)
        (outs * w.to(dt)).sum().backward()
        res[dt] = (o.grad, {k: v.grad for k, v in pp.items()}, outs.detach(), tuple(h.detach() for h in hid))
    return res


@pytest.mark.parametrize("mode", ["step", "seq", "mixed"])
@pytest.mark.parametrize("x_grad", [True, False])
@pytest.mark.parametrize("B,N,F,T,hops", [(9, 16, 32, 24, (1, 2, 4)), (5, 128, 32, 40, (1, 2, 4)), (7, 12, 8, 30, (1,)),
                                          (3, 20, 16, 26, (1, 3)), (300, 16, 32, 21, (1, 2, 4)), (130, 12, 32, 9, (3,)),
                                          (256, 64, 32, 45, (1, 2, 4))])               # whole tiles of 128 rows per position
def test_temporal_window_backward_matches_fp64_oracle(mode, x_grad, B, N, F, T, hops):
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    spec = [("temporal", hops, "forward")]
    gen = torch.Generator().manual_seed(100 + N + F)
    obs = torch.randn(T, B, F, generator=gen) * 0.5
    w = torch.randn(T, B, 32, generator=gen)
    p = oracle.make_params(F, 32)
    res = _oracle_grads(obs, w, p, spec, N)
    gnn, convs = make_dense_gnn(F, 32, p, ("tanh", "tanh"))
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
    x = obs.to(dev).requires_grad_(x_grad)
    xb = x.transpose(0, 1)                       # [B, T, F] view for the sequence entry
    hidden = None
    if mode == "step":
        outs = []
        for t in range(T):
            o, hidden = mod(x[t], hidden)
            outs.append(o)
        outs = torch.stack(outs)
    elif mode == "seq":
        outs, hidden = mod.forward_sequence(xb, hidden)
        outs = outs.transpose(0, 1)
    else:
        a, hidden = mod.forward_sequence(xb[:, :7], hidden)
        b, hidden = mod(x[7], hidden)
        c, hidden = mod.forward_sequence(xb[:, 8:], hidden)
        outs = torch.cat([a.transpose(0, 1), b.unsqueeze(0), c.transpose(0, 1)])
    assert getattr(hidden.token, "_gcm_tw", False), "the window-level backward should have been recorded"
    ref64, ref32 = res[torch.float64], res[torch.float32]
    assert rel_err(outs, ref64[2]) < TOL + rel_err(ref32[2], ref64[2])
    (outs * w.to(dev)).sum().backward()
    if x_grad:
        assert rel_err(x.grad, ref64[0]) < TOL + rel_err(ref32[0], ref64[0])
    got = named_grads(convs)
    for k in got:
        assert rel_err(got[k], ref64[1][k]) < TOL + rel_err(ref32[1][k], ref64[1][k]), k
    nodes, adj, _, num_nodes = hidden
    assert torch.equal(adj.cpu(), ref32[3][1]) and torch.equal(num_nodes.cpu(), ref32[3][3])
    assert torch.equal(nodes.detach().cpu(), ref32[3][0])


@pytest.mark.parametrize("x_grad", [False, True])
def test_fused_window_kernel_serves_f32_shapes_and_matches_the_separate_products(x_grad, monkeypatch):
    """F = H = 32: the root runs ONE fused kernel (gcm_temporal_window_bwd; with dz2 formed inside the shift-sum and the
    operand rows left by the forward kernel when the sequence node is the whole window); GCM_B200_NO_WINDOW_BWD_TC=1
    switches back to the separate products.  Same gradients either way."""
    from gcm import _cabi
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    B, N, F, T, hops = 260, 16, 32, 19, (1, 2, 4)
    spec = [("temporal", hops, "forward")]
    gen = torch.Generator().manual_seed(5)
    obs = (torch.randn(T, B, F, generator=gen) * 0.5).to(dev)
    w = torch.randn(T, B, 32, generator=gen).to(dev)
    p = oracle.make_params(F, 32)
    res = {}
    lib = _cabi.lib()
    real, calls = lib.gcm_temporal_window_bwd, []

    def spy(*a):
        calls.append(a[2])
        return real(*a)
    monkeypatch.setattr(lib, "gcm_temporal_window_bwd", spy)
    for fused in (True, False):
        calls.clear()
        if fused:
            monkeypatch.delenv("GCM_B200_NO_WINDOW_BWD_TC", raising=False)
        else:
            monkeypatch.setenv("GCM_B200_NO_WINDOW_BWD_TC", "1")
        gnn, convs = make_dense_gnn(F, 32, p, ("tanh", "tanh"))
        mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
        with torch.no_grad():                                   # a running rollout first: the window starts mid-log
            _, hidden = mod.forward_sequence(obs[:7].transpose(0, 1), None)
        x = obs.clone().requires_grad_(x_grad)
        outs, hidden = mod.forward_sequence(x.transpose(0, 1), hidden)
        (outs.transpose(0, 1) * w).sum().backward()
        assert calls == ([(T + 4) * B] if fused else []), calls      # one launch over the window's rows + max_hop planes
        res[fused] = (named_grads(convs), None if not x_grad else x.grad.clone())
    for k in res[True][0]:
        assert rel_err(res[True][0][k], res[False][0][k]) < 2e-5, k
    if x_grad:
        assert rel_err(res[True][1], res[False][1]) < 2e-5


@pytest.mark.parametrize("layout", ["sum", "batch_major", "sliced", "bf16"])
def test_window_backward_reads_the_belief_gradient_through_its_strides(layout, monkeypatch):
    """dL/dbelief reaches the sequence node as autograd made it: the expanded scalar of a sum loss (all strides 0; torch's
    mean backward divides AFTER expanding, so a mean loss arrives dense), a
    [B, T, H] tensor (seen time-major: transposed strides), a slice of a wider tensor (rows not 16-byte aligned).  The fused
    window backward reads it in place (gcm_temporal_shift_sum_strided: no materialised broadcast, no transposing copy);
    other dtypes are converted first.  Same gradients as the separate-products path, which works on a contiguous copy."""
    from gcm import _cabi
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    B, N, F, T, hops = 200, 16, 32, 23, (1, 2, 4)
    spec = [("temporal", hops, "forward")]
    gen = torch.Generator().manual_seed(11)
    obs = (torch.randn(T, B, F, generator=gen) * 0.5).to(dev)
    w_bt = torch.randn(B, T, 32, generator=gen).to(dev)
    w_wide = torch.randn(T, B, 33, generator=gen).to(dev)
    p = oracle.make_params(F, 32)
    lib = _cabi.lib()
    real, calls = lib.gcm_temporal_shift_sum_strided, []

    def spy(*a):
        calls.append(a[1:4])
        return real(*a)
    monkeypatch.setattr(lib, "gcm_temporal_shift_sum_strided", spy)
    res = {}
    for fused in (True, False):
        calls.clear()
        if fused:
            monkeypatch.delenv("GCM_B200_NO_WINDOW_BWD_TC", raising=False)
        else:
            monkeypatch.setenv("GCM_B200_NO_WINDOW_BWD_TC", "1")
        gnn, convs = make_dense_gnn(F, 32, p, ("tanh", "tanh"))
        mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
        with torch.no_grad():                                   # a running rollout first: the window is ONE sequence node
            _, hidden = mod.forward_sequence(obs[:7].transpose(0, 1), None)
        x = obs.clone().requires_grad_(True)
        if layout == "batch_major":
            outs, hidden = mod.forward_sequence(x.transpose(0, 1), hidden)                 # [B, T, H]
            (outs * w_bt).sum().backward()
            want = [(32, T * 32, 1)]
        else:
            outs, hidden = mod.forward_sequence(x, hidden, time_major=True)                # [T, B, H]
            if layout == "sum":
                (outs.sum() / outs.numel()).backward()
                want = [(0, 0, 0)]
            elif layout == "sliced":
                (torch.cat([outs, outs.new_zeros(T, B, 1)], dim=2) * w_wide).sum().backward()   # gradient: a slice of [T,B,33]
                want = [(B * 33, 33, 1)]
            else:
                (outs.to(torch.bfloat16) * w_wide[..., :32].to(torch.bfloat16)).sum().backward()
                want = []                                                                  # converted to float32: contiguous
        if fused and layout != "bf16":
            assert calls == want, calls
        elif not fused:
            assert calls == []
        res[fused] = (named_grads(convs), x.grad.clone())
    tol = 2e-5 if layout != "bf16" else 1e-2
    for k in res[True][0]:
        assert rel_err(res[True][0][k], res[False][0][k]) < tol, k
    assert rel_err(res[True][1], res[False][1]) < tol


def test_truncated_bptt_over_windows_with_weight_updates():
    """A running rollout trained window by window (m_t.detach(), SGD step in between): fill without autograd, then three
    windows; every window's gradients against the fp64 oracle doing the same."""
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    B, N, F, T, hops = 6, 16, 32, 10, (1, 2, 4)
    spec = [("temporal", hops, "forward")]
    gen = torch.Generator().manual_seed(9)
    fill = torch.randn(N + 3, B, F, generator=gen) * 0.5
    p = oracle.make_params(F, 32)
    gnn, convs = make_dense_gnn(F, 32, p, ("tanh", "tanh"))
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
    mod.bptt_capacity = T
    opt = torch.optim.SGD(mod.parameters(), lr=0.05)
    with torch.no_grad():
        _, hidden = mod.forward_sequence(fill.to(dev).transpose(0, 1), None)
        o_hid = {}
        for dt in (torch.float64, torch.float32):
            _, o_hid[dt] = oracle.dense_gcm_rollout(fill.to(dt), None, spec, {k: v.to(dt) for k, v in p.items()},
                                                    ("tanh", "tanh"), graph_size=N)
    for wi in range(3):
        obs = torch.randn(T, B, F, generator=gen) * 0.5
        w = torch.randn(T, B, 32, generator=gen)
        res = _oracle_grads(obs, w, p, spec, N, hidden=o_hid)
        opt.zero_grad(set_to_none=True)
        hidden = hidden.detach()
        if wi == 1:
            outs = []
            for t in range(T):
                o, hidden = mod(obs[t].to(dev), hidden)
                outs.append(o)
            outs = torch.stack(outs)
        else:
            outs, hidden = mod.forward_sequence(obs.to(dev).transpose(0, 1), hidden)
            outs = outs.transpose(0, 1)
        assert getattr(hidden.token, "_gcm_tw", False)
        (outs * w.to(dev)).sum().backward()
        ref64, ref32 = res[torch.float64], res[torch.float32]
        assert rel_err(outs, ref64[2]) < TOL + rel_err(ref32[2], ref64[2]), wi
        got = named_grads(convs)
        for k in got:
            assert rel_err(got[k], ref64[1][k]) < TOL + rel_err(ref32[1][k], ref64[1][k]), (wi, k)
        opt.step()
        # the oracle takes the same SGD step (with ITS fp64 gradients) and carries its own hidden state on
        p = {k: (v.double() - 0.05 * ref64[1][k]).float() for k, v in p.items()}
        with torch.no_grad():
            for k, v in named_params(convs).items():
                v.copy_(p[k])
        o_hid = {dt: res[dt][3] for dt in res}
    nodes, adj, _, num_nodes = hidden
    assert torch.equal(adj.cpu(), o_hid[torch.float32][1]) and torch.equal(nodes.detach().cpu(), o_hid[torch.float32][0])


def named_params(convs):
    return {"w_rel1": convs[0].lin_rel.weight, "b1": convs[0].lin_rel.bias, "w_root1": convs[0].lin_root.weight,
            "w_rel2": convs[1].lin_rel.weight, "b2": convs[1].lin_rel.bias, "w_root2": convs[1].lin_root.weight}


def test_window_capacity_truncation_and_unused_beliefs():
    """More recorded steps than the log keeps: the forward keeps going (the reference records arbitrarily long grad-mode
    rollouts; an eval loop without torch.no_grad() is legal) on a fresh chain with a one-time warning, beliefs stay
    right, and backward() through the steps left behind raises instead of returning wrong gradients.  Steps whose
    belief does not reach the loss deliver nothing and cost nothing."""
    import warnings

    from gcm import fused
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    B, N, F, T = 4, 16, 32, 30
    spec = [("temporal", (1, 2), "forward")]
    p = oracle.make_params(F, 32)
    gnn, convs = make_dense_gnn(F, 32, p, ("tanh", "tanh"))
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N, bptt_capacity=8)
    gen = torch.Generator().manual_seed(1)
    obs = torch.randn(T, B, F, generator=gen) * 0.5
    hidden, o_hidden = None, None
    outs = []
    fused._warned_truncated[0] = False
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        for t in range(T):
            o, hidden = mod(obs[t].to(dev), hidden)
            outs.append(o)
            ref, o_hidden = oracle.dense_gcm_step(obs[t], o_hidden, spec, p, graph_size=N)
            assert rel_err(o, ref) < TOL, t
    assert sum("autograd history is cut" in str(w.message) for w in caught) == 1
    with pytest.raises(RuntimeError):
        outs[2].sum().backward()                       # recorded before the cut: its window is gone
    outs[-1].sum().backward()                          # the current chain still works
    assert convs[0].lin_rel.weight.grad is not None
    # only step 5's belief is used
    o = obs.clone().requires_grad_(True)
    pp = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    ref, _ = oracle.dense_gcm_rollout(o[:8], None, spec, pp, ("tanh", "tanh"), graph_size=N)
    ref[5].sum().backward()
    mod.zero_grad(set_to_none=True)
    hidden, outs = None, []
    for t in range(8):
        b, hidden = mod(obs[t].to(dev), hidden)
        outs.append(b)
    outs[5].sum().backward()
    got = named_grads(convs)
    for k in got:
        assert rel_err(got[k], pp[k].grad) < 5 * TOL, k


@pytest.mark.parametrize("mode", ["step", "seq", "handle"])
@pytest.mark.parametrize("name", ["train_preproc_temporal", "train_preproc_temporal_young"])
def test_training_with_preprocessor_matches_reference_golden(name, mode):
    """RayDenseGCM's configuration in TRAINING (Linear preprocessor + TemporalBackedge): a BPTT window recorded on the
    window-level backward, started from the plain tensors RLlib hands back ("step" / "seq") or from a live handle after
    a no-grad rollout ("handle").  Beliefs, dL/dobs, the six GNN gradients and the preprocessor's gradients -- which
    include the contributions through rows written before the window (gcm.py:290-291) -- against the fixtures written
    by the unmodified reference."""
    from helpers import load_golden
    from gcm.gcm import DenseGCM

    g = load_golden(name)
    dev = torch.device("cuda:0")
    gnn, convs = make_dense_gnn(g["F"], g["H"], g["params"], ("tanh", "tanh"))
    pre = torch.nn.Linear(g["F_raw"], g["F"])
    with torch.no_grad():
        pre.weight.copy_(g["pre_weight"])
        pre.bias.copy_(g["pre_bias"])
    mod = DenseGCM(gnn.to(dev), preprocessor=pre.to(dev), edge_selectors=make_selector(g["spec"]), graph_size=g["N"])
    T0, T1 = g["T0"], g["T1"]
    if mode == "handle":
        with torch.no_grad():
            _, hidden = mod.forward_sequence(g["obs"][:T0].to(dev), None, time_major=True)
        hidden = hidden.detach()
    else:
        hidden = tuple(t.to(dev) for t in g["start"])
    x = g["obs"][T0:].to(dev).requires_grad_(True)
    if mode == "step":
        outs = []
        for t in range(T1):
            o, hidden = mod(x[t], hidden)
            outs.append(o)
        outs = torch.stack(outs)
    else:
        outs, hidden = mod.forward_sequence(x, hidden, time_major=True)
    assert getattr(hidden.token, "_gcm_tw", False), "the window-level backward should have been recorded"
    assert rel_err(outs, g["beliefs"]) < 5 * TOL
    (outs * g["loss_w"].to(dev)).sum().backward()
    assert rel_err(x.grad, g["d_obs"]) < 5 * TOL
    got = named_grads(convs)
    for k, v in g["d_params"].items():
        assert rel_err(got[k], v) < 5 * TOL, k
    assert rel_err(pre.weight.grad, g["d_pre_weight"]) < 5 * TOL and rel_err(pre.bias.grad, g["d_pre_bias"]) < 5 * TOL
    final = tuple(hidden)
    assert torch.equal(final[0].cpu(), g["final"][0]) and torch.equal(final[1].cpu(), g["final"][1].float())
    assert torch.equal(final[3].cpu(), g["final"][3])


def test_training_windows_with_preprocessor_updates_match_the_generic_path():
    """Truncated BPTT over several windows with SGD updates of the GNN AND of the Linear preprocessor in between: the fused
    route (own product kernels for the preprocessor, only the last 2 max_hop stored rows re-imaged after an update, fused
    window backward) against the same module on the generic torch route (reference statement order, gcm.py:241-321, every
    stored row through the preprocessor at every step).  Beliefs of every window and the parameters after every update."""
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    B, N, F_raw, F, H, T = 37, 12, 16, 32, 32, 7
    spec = [("temporal", (1, 2, 4), "forward")]
    gen = torch.Generator().manual_seed(11)
    obs = (torch.randn(5, T, B, F_raw, generator=gen) * 0.7).to(dev)
    w = torch.randn(5, T, B, H, generator=gen).to(dev)
    p = oracle.make_params(F, H)
    mods = []
    for fused_route in (True, False):
        gnn, convs = make_dense_gnn(F, H, p, ("tanh", "tanh"))
        pre = torch.nn.Linear(F_raw, F)
        with torch.no_grad():
            g2 = torch.Generator().manual_seed(12)
            pre.weight.copy_(torch.randn(F, F_raw, generator=g2) / F_raw ** 0.5)
            pre.bias.copy_(0.1 * torch.randn(F, generator=g2))
        mod = DenseGCM(gnn.to(dev), preprocessor=pre.to(dev), edge_selectors=make_selector(spec), graph_size=N)
        if not fused_route:
            mod._plan, mod._plan_built = None, True            # no fused plan: DenseGCM._forward_generic
        mods.append((mod, torch.optim.SGD(mod.parameters(), lr=0.01)))
    hidden = [None, None]
    pre0 = [q.detach().clone() for q in mods[0][0].preprocessor.parameters()]
    for win in range(4):
        outs = []
        for i, (mod, opt) in enumerate(mods):
            opt.zero_grad(set_to_none=True)
            h = hidden[i]
            if h is not None:
                h = h.detach() if hasattr(h, "detach") else tuple(t.detach() for t in h)
            if i == 0:
                o, h = mod.forward_sequence(obs[win], h, time_major=True)
            else:
                steps = []
                for t in range(T):
                    ot, h = mod(obs[win, t], h)
                    steps.append(ot)
                o = torch.stack(steps)
            (o * w[win]).sum().backward()
            opt.step()
            hidden[i] = h
            outs.append(o.detach())
        # the updates are large (lr 0.01 on a sum loss: the preprocessor's weights move by > 1e-2 per window), so a stale
        # image of a stored row would show at the 1e-2 level; fp32 rounding of two different evaluation orders, amplified
        # through the updates, stays below 1e-3
        tol = 2e-5 if win == 0 else 1e-3
        assert rel_err(outs[0], outs[1]) < tol, f"beliefs of window {win}"
        for (n0, p0), (n1, p1) in zip(mods[0][0].named_parameters(), mods[1][0].named_parameters()):
            assert n0 == n1 and rel_err(p0.detach(), p1.detach()) < (5e-5 if win == 0 else 1e-3), f"{n0} after window {win}"
        if win == 0:
            moved = max(float((p0.detach() - q0).abs().max()) for p0, q0 in zip(mods[0][0].preprocessor.parameters(), pre0))
            assert moved > 1e-2, "the test needs updates that matter"
    assert getattr(hidden[0].token, "_gcm_tw", False)
