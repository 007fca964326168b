"""GPU tier: BPTT through the fused path (gcm_dense_step_bwd) against the gradients of the
unmodified reference (golden) and of the fp64 oracle (autograd on CPU)."""
import pytest
import torch

import gcm_oracle as oracle
from helpers import dense_cases, load_golden, make_dense_gnn, make_selector, named_grads, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _bptt(mod, convs, obs_dev, loss_w, hidden=None):
    outs = []
    for t in range(obs_dev.shape[0]):
        belief, hidden = mod(obs_dev[t], hidden)
        outs.append(belief)
    outs = torch.stack(outs)
    (outs * loss_w).sum().backward()
    return outs, hidden


@pytest.mark.parametrize("name", [n for n in dense_cases() if "grad" in n])
def test_bptt_matches_reference_golden(name):
    from gcm.gcm import DenseGCM

    g = load_golden(name)
    dev = torch.device("cuda:0")
    gnn, convs = make_dense_gnn(g["F"], g["H"], g["params"], g["acts"])
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(g["spec"]), graph_size=g["N"])
    obs = g["obs"].to(dev).requires_grad_(True)
    outs, _ = _bptt(mod, convs, obs, g["loss_w"].to(dev))
    assert rel_err(outs, g["beliefs"]) < TOL
    assert rel_err(obs.grad, g["d_obs"]) < 5 * TOL, name
    got = named_grads(convs)
    for k, v in g["d_params"].items():
        assert rel_err(got[k], v) < 5 * TOL, (name, k)


SEEDED = [
    (9, 16, 32, 32, 24, [("temporal", (1, 2, 4), "forward")], ("tanh", "tanh")),   # wraps inside the window
    (4, 12, 8, 32, 10, [("temporal", (1,), "both")], ("tanh", "tanh")),
    (3, 10, 20, 12, 14, [("dense",)], ("tanh", "relu")),                           # wraps, complement trick off
    (2, 40, 16, 24, 30, [("dense",)], ("tanh", "tanh")),                           # nR > 16: complement trick on
    (3, 24, 12, 10, 30, [("cosine", 0.5)], ("tanh", "tanh")),
    (3, 12, 16, 16, 18, [("temporal", (1, 2), "forward"), ("dense",)], ("tanh", "tanh")),   # chain with DenseEdge: ones path
    (3, 9, 70, 40, 6, [("temporal", (2,), "backward"), ("temporal", (1,), "forward")], ("relu", "none")),
]


@pytest.mark.parametrize("B,N,F,H,T,spec,acts", SEEDED)
def test_bptt_matches_fp64_oracle(B, N, F, H, T, spec, acts):
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(4321 + N + F)
    if spec[0][0] == "cosine":
        centres = torch.randn(4, F, generator=gen) * 1.5
        obs = centres[torch.randint(0, 4, (T,), generator=gen)].unsqueeze(1).expand(T, B, F) \
            + 0.03 * torch.randn(T, B, F, generator=gen)
        obs = obs.contiguous()
    else:
        obs = torch.randn(T, B, F, generator=gen) * 0.5
    w = torch.randn(T, B, H, generator=gen)
    p = oracle.make_params(F, H)
    # fp64 oracle (autograd) is the ground truth; the fp32 oracle gives the reference's own rounding distance
    res = {}
    for dt in (torch.float64, torch.float32):
        o = obs.to(dt).clone().requires_grad_(True)
        pp = {k: v.to(dt).clone().requires_grad_(True) for k, v in p.items()}
        outs, _ = oracle.dense_gcm_rollout(o, None, spec, pp, acts, graph_size=N)
        (outs * w.to(dt)).sum().backward()
        res[dt] = (o.grad, {k: v.grad for k, v in pp.items()})
    gnn, convs = make_dense_gnn(F, H, p, acts)
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
    x = obs.to(dev).requires_grad_(True)
    _bptt(mod, convs, x, w.to(dev))
    ref64, ref32 = res[torch.float64], res[torch.float32]
    assert rel_err(x.grad, ref64[0]) < TOL + rel_err(ref32[0], ref64[0])
    got = named_grads(convs)
    for k in got:
        assert rel_err(got[k], ref64[1][k]) < TOL + rel_err(ref32[1][k], ref64[1][k]), k


def test_grad_reaches_user_supplied_nodes():
    """Reference tests/test_gcm.py:355-365: gradients flow into a caller-supplied `nodes` tensor."""
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    B, N, F, H = 3, 7, 5, 6
    gen = torch.Generator().manual_seed(3)
    p = oracle.make_params(F, H)
    nodes0 = torch.randn(B, N, F, generator=gen)
    nn0 = torch.tensor([2, 7, 4])
    adj0 = torch.zeros(B, N, N)
    for b in range(B):
        nodes0[b, int(nn0[b]):] = 0
        for i in range(1, int(nn0[b])):
            adj0[b, i, i - 1] = 1
    obs = torch.randn(4, B, F, generator=gen)
    spec = [("temporal", (1, 2), "forward")]
    # oracle
    n_o = nodes0.clone().requires_grad_(True)
    hidden = (n_o, adj0, torch.zeros(0), nn0)
    outs = []
    for t in range(4):
        mx, hidden = oracle.dense_gcm_step(obs[t], hidden, spec, p, graph_size=N)
        outs.append(mx)
    torch.stack(outs).sum().backward()
    # fused
    gnn, convs = make_dense_gnn(F, H, p, ("tanh", "tanh"))
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
    n_f = nodes0.clone().to(dev).requires_grad_(True)
    hidden = (n_f, adj0.to(dev), torch.zeros(0, device=dev), nn0.to(dev))
    outs = []
    for t in range(4):
        mx, hidden = mod(obs[t].to(dev), hidden)
        outs.append(mx)
    torch.stack(outs).sum().backward()
    assert rel_err(n_f.grad, n_o.grad) < 5 * TOL


def test_it_learns_and_detach():
    """Reference tests/test_gcm.py:412-439 ("it learns"): 20 Adam steps reduce the loss; truncated BPTT
    through m_t.detach() keeps working on the same in-place state."""
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    B, N, F, H, T = 5, 10, 11, 11, 4
    p = oracle.make_params(F, H)
    gnn, convs = make_dense_gnn(F, H, p, ("relu", "relu"))
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector([("temporal", (1,), "forward")]), graph_size=N)
    opt = torch.optim.Adam(mod.parameters(), lr=0.005)
    losses = []
    for _ in range(20):
        opt.zero_grad()
        obs, hidden = torch.ones(B, F, device=dev), None
        for t in range(T):
            obs, hidden = mod(obs, hidden)
        loss = torch.norm(obs)
        loss.backward()
        losses.append(float(loss))
        opt.step()
    assert losses[-1] < losses[0]
    hidden = None
    for window in range(3):
        opt.zero_grad()
        tot = 0
        for t in range(6):
            out, hidden = mod(torch.randn(B, F, device=dev), hidden)
            tot = tot + out.pow(2).mean()
        tot.backward()
        opt.step()
        hidden = hidden.detach()
    assert int(tuple(hidden)[3].max()) == N


@pytest.mark.parametrize("T,n0", [(8, 10), (20, 10), (5, 24)])
def test_ones_path_bptt_from_prefilled_dense_state(T, n0):
    """BASELINE cfg3's protocol at a reduced size: a caller-supplied hidden state whose adjacency is DenseEdge's
    all-ones block (ragged counts), T recorded steps (with and without the window wrapping inside the chain),
    gradients of the observations and of all six weight tensors against the fp64 oracle; then a second window
    on the detached state (truncated BPTT).  The DenseEdge-only kernels (csrc/gcm_dense_ones.cu) must be the
    ones that ran."""
    from gcm import _cabi
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    B, N, F, H = 4, 24, 16, 12
    spec = [("dense",)]
    acts = ("tanh", "tanh")
    gen = torch.Generator().manual_seed(77 + T)
    p = oracle.make_params(F, H)
    nn0 = torch.tensor([n0, max(n0 - 3, 0), n0, max(n0 - 1, 0)])
    nodes0 = 0.5 * torch.randn(B, N, F, generator=gen)
    adj0 = torch.zeros(B, N, N)
    for b in range(B):
        nodes0[b, int(nn0[b]):] = 0
        adj0[b, : int(nn0[b]), : int(nn0[b])] = 1
    obs = 0.5 * torch.randn(T, B, F, generator=gen)
    w = torch.randn(T, B, H, generator=gen)
    res = {}
    for dt in (torch.float64, torch.float32):
        o = obs.to(dt).clone().requires_grad_(True)
        pp = {k: v.to(dt).clone().requires_grad_(True) for k, v in p.items()}
        outs, _ = oracle.dense_gcm_rollout(o, (nodes0.to(dt), adj0.to(dt), torch.zeros(0, dtype=dt), nn0.clone()),
                                           spec, pp, acts, graph_size=N)
        (outs * w.to(dt)).sum().backward()
        res[dt] = (outs.detach(), o.grad, {k: v.grad for k, v in pp.items()})
    gnn, convs = make_dense_gnn(F, H, p, acts)
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
    x = obs.to(dev).requires_grad_(True)
    hidden = (nodes0.to(dev), adj0.to(dev), torch.zeros(0, device=dev), nn0.to(dev))
    outs, hidden = _bptt(mod, convs, x, w.to(dev), hidden)
    assert _cabi.lib().gcm_last_kernel().decode() in ("k_outer_reduce", "k_linear2", "k_ones_window_bwd")
    ref64, ref32 = res[torch.float64], res[torch.float32]
    assert rel_err(outs, ref64[0]) < TOL + rel_err(ref32[0], ref64[0])
    assert rel_err(x.grad, ref64[1]) < TOL + rel_err(ref32[1], ref64[1])
    got = named_grads(convs)
    for k in got:
        assert rel_err(got[k], ref64[2][k]) < TOL + rel_err(ref32[2][k], ref64[2][k]), k
    # the state the caller sees is the reference's (adjacency materialised from the implicit all-ones block)
    _, o_hidden = oracle.dense_gcm_rollout(obs, (nodes0, adj0, torch.zeros(0), nn0.clone()), spec, p, acts, graph_size=N)
    nodes, adj, _, num_nodes = hidden
    assert torch.equal(nodes.cpu(), o_hidden[0]) and torch.equal(adj.cpu(), o_hidden[1])
    assert torch.equal(num_nodes.cpu(), o_hidden[3])
    # truncated BPTT: a second window on the detached in-place state
    for q in convs:
        q.zero_grad()
    x2 = obs[:3].to(dev).requires_grad_(True)
    _bptt(mod, convs, x2, w[:3].to(dev), hidden.detach())
    o2 = obs[:3].double().clone().requires_grad_(True)
    pp = {k: v.double().clone().requires_grad_(True) for k, v in p.items()}
    outs2, _ = oracle.dense_gcm_rollout(o2, tuple(t.double() if t.is_floating_point() else t for t in o_hidden), spec, pp,
                                        acts, graph_size=N)
    (outs2 * w[:3].double()).sum().backward()
    assert rel_err(x2.grad, o2.grad) < 5 * TOL
    got = named_grads(convs)
    for k in got:
        assert rel_err(got[k], pp[k].grad) < 5 * TOL, k


@pytest.mark.parametrize("acts,bf16", [(("tanh", "tanh"), False), (("tanh", "tanh"), True), (("relu", "none"), False)])
def test_ones_path_window_backward_wide(acts, bf16):
    """The ones path at a width where every vector lane of its kernels is used (H1 = 64: 16 / 8 threads per cache
    row), ragged pre-filled counts, a chain that wraps, observations with and without grad.  float32 cache:
    1e-5 against the fp64 oracle; bfloat16 cache (`DenseGCM.compute_dtype = torch.bfloat16`, BASELINE cfg3's
    precision): 2e-2."""
    from gcm import _cabi
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    B, N, F, H, T = 6, 40, 32, 64, 12
    spec = [("dense",)]
    gen = torch.Generator().manual_seed(5)
    p = oracle.make_params(F, H)
    nn0 = torch.tensor([33, 0, 40, 17, 39, 5])
    nodes0 = 0.5 * torch.randn(B, N, F, generator=gen)
    adj0 = torch.zeros(B, N, N)
    for b in range(B):
        nodes0[b, int(nn0[b]):] = 0
        adj0[b, : int(nn0[b]), : int(nn0[b])] = 1
    obs = 0.5 * torch.randn(T, B, F, generator=gen)
    w = torch.randn(T, B, H, generator=gen)
    o = obs.double().clone().requires_grad_(True)
    pp = {k: v.double().clone().requires_grad_(True) for k, v in p.items()}
    outs, _ = oracle.dense_gcm_rollout(o, (nodes0.double(), adj0.double(), torch.zeros(0, dtype=torch.float64), nn0.clone()),
                                       spec, pp, acts, graph_size=N)
    (outs * w.double()).sum().backward()
    tol = 2e-2 if bf16 else 5 * TOL
    for x_grad in (True, False):
        gnn, convs = make_dense_gnn(F, H, p, acts)
        mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
        mod.bptt_capacity = T
        if bf16:
            mod.compute_dtype = torch.bfloat16
        x = obs.to(dev).requires_grad_(x_grad)
        hidden = (nodes0.to(dev), adj0.to(dev), torch.zeros(0, device=dev), nn0.to(dev))
        got_outs, hidden = _bptt(mod, convs, x, w.to(dev), hidden)
        assert hidden.claim().rc_bf16 == bf16 and mod._plan is not None and hidden.claim().win is not None
        assert rel_err(got_outs, outs.detach()) < tol
        if x_grad:
            assert rel_err(x.grad, o.grad) < tol
        got = named_grads(convs)
        for k in got:
            assert rel_err(got[k], pp[k].grad) < tol, (k, x_grad)


@pytest.mark.parametrize("bf16", [False, True])
@pytest.mark.parametrize("x_grad", [True, False])
@pytest.mark.parametrize("contig", [False, True])
def test_forward_sequence_matches_step_loop_and_oracle(bf16, x_grad, contig):
    """DenseGCM.forward_sequence (the caller is RayDenseGCM's loop over T, ray_gcm.py:200-202) on a
    DenseEdge state: T steps at once must give what T forward() calls give -- beliefs, the hidden state, and every
    gradient (fp64 oracle: 1e-5 class for the float32 cache, 2e-2 for bfloat16).  Ragged pre-filled counts; the
    window wraps inside the sequence; a second sequence continues the same BPTT chain.  contig: the chunks are contiguous
    tensors of their own, so the second call (live handle) takes ALL its steps through the sequence kernels and returns a
    view of the time-major buffer; otherwise they are slices and every call enters through one forward() step."""
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    B, N, F, H, T1, T2 = 5, 40, 32, 64, 9, 7
    T = T1 + T2
    spec = [("dense",)]
    acts = ("tanh", "tanh")
    gen = torch.Generator().manual_seed(11)
    p = oracle.make_params(F, H)
    nn0 = torch.tensor([33, 0, 40, 17, 38])
    nodes0 = 0.5 * torch.randn(B, N, F, generator=gen)
    adj0 = torch.zeros(B, N, N)
    for b in range(B):
        nodes0[b, int(nn0[b]):] = 0
        adj0[b, : int(nn0[b]), : int(nn0[b])] = 1
    obs = 0.5 * torch.randn(T, B, F, generator=gen)
    w = torch.randn(T, B, H, generator=gen)
    o = obs.double().clone().requires_grad_(True)
    pp = {k: v.double().clone().requires_grad_(True) for k, v in p.items()}
    outs, o_hidden = oracle.dense_gcm_rollout(o, (nodes0.double(), adj0.double(), torch.zeros(0, dtype=torch.float64), nn0.clone()),
                                              spec, pp, acts, graph_size=N)
    (outs * w.double()).sum().backward()
    tol = 2e-2 if bf16 else 5 * TOL
    gnn, convs = make_dense_gnn(F, H, p, acts)
    mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
    mod.bptt_capacity = T
    if bf16:
        mod.compute_dtype = torch.bfloat16
    x = obs.to(dev).transpose(0, 1).contiguous().requires_grad_(x_grad and not contig)            # [B, T, F]
    hidden = (nodes0.to(dev), adj0.to(dev), torch.zeros(0, device=dev), nn0.to(dev))
    if contig:
        x1 = x[:, :T1].detach().contiguous().requires_grad_(x_grad)
        x2 = x[:, T1:].detach().contiguous().requires_grad_(x_grad)
    else:
        x1, x2 = x[:, :T1], x[:, T1:]
    b1, hidden = mod.forward_sequence(x1, hidden)
    assert hidden.claim().win is not None and hidden.claim().steps == T1
    b2, hidden = mod.forward_sequence(x2, hidden)
    if contig:
        assert not b2.is_contiguous() and b2.transpose(0, 1).is_contiguous()        # a view of the [T, B, H] buffer
    got = torch.cat([b1, b2], dim=1)                                                # [B, T, H]
    assert got.shape == (B, T, H)
    assert rel_err(got.transpose(0, 1), outs.detach()) < tol
    (got.transpose(0, 1) * w.to(dev)).sum().backward()
    if x_grad:
        xg = torch.cat([x1.grad, x2.grad], dim=1) if contig else x.grad
        assert rel_err(xg.transpose(0, 1), o.grad) < tol
    grads = named_grads(convs)
    for k in grads:
        assert rel_err(grads[k], pp[k].grad) < tol, k
    nodes, adj, _, num_nodes = hidden
    assert torch.equal(nodes.cpu(), o_hidden[0].float()) and torch.equal(adj.cpu(), o_hidden[1].float())
    assert torch.equal(num_nodes.cpu(), o_hidden[3])


def test_forward_sequence_rollout_ring_and_generic_fallback():
    """No-grad sequences on an in-place ring (C == N): 3 x 30 steps from an empty 24-node DenseEdge graph (the window
    wraps inside the second sequence) against the step loop of a second module; and a TemporalBackedge module, whose
    forward_sequence is the plain loop."""
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    B, N, F, H, T = 7, 24, 16, 32, 30
    p = oracle.make_params(F, H)
    gen = torch.Generator().manual_seed(3)
    obs = (0.5 * torch.randn(B, 3 * T, F, generator=gen)).to(dev)
    for spec in ([("dense",)], [("temporal", (1, 2), "forward")]):
        mods = []
        for _ in range(2):
            gnn, _ = make_dense_gnn(F, H, p, ("tanh", "tanh"))
            mods.append(DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N))
        with torch.no_grad():
            h_seq, h_loop, seq_out, loop_out = None, None, [], []
            for c in range(3):
                chunk = obs[:, c * T:(c + 1) * T]
                out, h_seq = mods[0].forward_sequence(chunk.contiguous() if c == 1 else chunk, h_seq)   # c == 1: whole-sequence entry
                seq_out.append(out)
            for t in range(3 * T):
                out, h_loop = mods[1](obs[:, t], h_loop)
                loop_out.append(out)
        seq_out, loop_out = torch.cat(seq_out, dim=1), torch.stack(loop_out, dim=1)
        assert rel_err(seq_out, loop_out) < TOL
        for a, b in zip(tuple(h_seq), tuple(h_loop)):
            assert torch.equal(a, b)
