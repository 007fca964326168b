"""GPU tier: the rollout entry of temporal chains (gcm.temporal, gcm_dense_rollout_fwd / gcm_dense_rollout_step).

DenseGCM.forward_sequence(x[B,T,F], m_t) must give exactly what the step loop of RayDenseGCM.forward gives
(reference ray_gcm.py:200-202: `for t in range(T): out, hidden = self.gcm(flat[:, t, :], hidden)`): the SAME kernels
run with the same arguments, so beliefs and the materialised state are compared bit for bit with the loop, and
within BASELINE.json's 1e-5 with the oracle."""
import pytest
import torch

import gcm_oracle as oracle
from helpers import make_dense_gnn, make_selector, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _pair(F, H, spec, N, pre_dims=None, seed=7):
    """two DenseGCMs with identical weights: one stepped by the loop, one through forward_sequence"""
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    p = oracle.make_params(F, H)
    mods = []
    pres = []
    for _ in range(2):
        gnn, convs = make_dense_gnn(F, H, p, ("tanh", "tanh"))
        pre = None
        if pre_dims is not None:
            torch.manual_seed(seed)
            pre = torch.nn.Linear(pre_dims, F).to(dev)
        mods.append((DenseGCM(gnn.to(dev), preprocessor=pre, edge_selectors=make_selector(spec), graph_size=N), convs))
        pres.append(pre)
    return p, mods, pres


@pytest.mark.parametrize("B,N,F,hops,chunks", [
    (70, 128, 32, (1, 2, 4), (1, 2, 7, 60, 64, 9)),        # BASELINE cfg2 shape; fills, wraps at 128 inside a call
    (33, 24, 8, (1,), (5, 30, 3)),                          # cfg1-like, wraps
    (200, 16, 16, (1, 3), (2, 2, 40)),
])
def test_temporal_sequence_equals_step_loop(B, N, F, hops, chunks):
    from gcm import _cabi
    from gcm.state import DenseHidden

    dev = torch.device("cuda:0")
    lib = _cabi.lib()
    spec = [("temporal", hops, "forward")]
    p, mods, _ = _pair(F, 32, spec, N)
    (m_loop, _), (m_seq, _) = mods
    T = sum(chunks)
    gen = torch.Generator().manual_seed(31)
    obs = torch.randn(B, T, F, generator=gen).to(dev)
    h_loop, h_seq, o_hidden = None, None, None
    t0 = 0
    with torch.no_grad():
        for ci, n in enumerate(chunks):
            x = obs[:, t0:t0 + n]                      # a strided [B, n, F] view of the caller's tensor
            l0 = lib.gcm_launch_count()
            out_seq, h_seq = m_seq.forward_sequence(x, h_seq)
            launched = lib.gcm_launch_count() - l0
            assert isinstance(h_seq, DenseHidden) and out_seq.shape == (B, n, 32)
            outs = []
            for t in range(n):
                o, h_loop = m_loop(x[:, t], h_loop)
                outs.append(o)
                ref, o_hidden = oracle.dense_gcm_step(x[:, t].cpu(), o_hidden, spec, p, graph_size=N)
                assert rel_err(o, ref) < TOL, (ci, t)
            assert torch.equal(out_seq, torch.stack(outs, dim=1)), (ci, n)
            if ci >= 2 and n >= 2:
                # warm state: every step on the cached-row kernel (one launch walks consecutive steps), no staging copies
                assert 1 <= launched <= n, (launched, n)
                assert lib.gcm_last_kernel().decode() == "k_step_temporal_hc"
            t0 += n
        for a, b in zip(h_seq, h_loop):
            assert torch.equal(a, b)
        nodes, adj, _, num_nodes = h_seq
        assert torch.equal(nodes.cpu(), o_hidden[0]) and torch.equal(adj.cpu(), o_hidden[1])
        assert torch.equal(num_nodes.cpu(), o_hidden[3])
        # time-major layout: [T, B, F] in, [T, B, H] out, same numbers
        xt = obs[:, :6].transpose(0, 1).contiguous()
        o_tm, h_seq = m_seq.forward_sequence(xt, h_seq, time_major=True)
        o_bm, h_loop = m_loop.forward_sequence(obs[:, :6], h_loop)
        assert o_tm.shape == (6, B, 32) and o_tm.is_contiguous() and torch.equal(o_tm.transpose(0, 1), o_bm)
        # the handle returned by a sequence call feeds the per-step fast path and the other way round
        o1, h_seq = m_seq(obs[:, 0].contiguous(), h_seq)
        o2, h_loop = m_loop(obs[:, 0].contiguous(), h_loop)
        assert torch.equal(o1, o2)


def test_temporal_sequence_follows_weight_updates():
    """An in-place weight update between two sequence calls: the C loop refills the row cache on the recomputing
    kernel for max_hop steps, then returns to the cached-row kernel -- as the step loop does."""
    dev = torch.device("cuda:0")
    B, N, F, hops = 45, 32, 32, (1, 2, 4)
    spec = [("temporal", hops, "forward")]
    p, mods, _ = _pair(F, 32, spec, N)
    (m_loop, c_loop), (m_seq, c_seq) = mods
    gen = torch.Generator().manual_seed(5)
    obs = torch.randn(B, 60, F, generator=gen).to(dev)
    h_loop = h_seq = None
    with torch.no_grad():
        for lo, hi in ((0, 20), (20, 40), (40, 60)):
            if lo:
                for convs in (c_loop, c_seq):
                    torch.manual_seed(lo)
                    for c in convs:
                        c.lin_rel.weight.add_(0.05 * torch.randn_like(c.lin_rel.weight))
                        c.lin_root.weight.add_(0.05 * torch.randn_like(c.lin_root.weight))
            out_seq, h_seq = m_seq.forward_sequence(obs[:, lo:hi], h_seq)
            outs = []
            for t in range(lo, hi):
                o, h_loop = m_loop(obs[:, t], h_loop)
                outs.append(o)
            assert torch.equal(out_seq, torch.stack(outs, dim=1)), lo
        for a, b in zip(h_seq, h_loop):
            assert torch.equal(a, b)


@pytest.mark.parametrize("spec", [[("temporal", (1, 2, 4), "forward")], [("dense",)]])
def test_sequence_with_rowwise_preprocessor(spec):
    """RayDenseGCM's configuration (ray_gcm.py:118,133-136: Linear preprocessor, F_raw != F_gnn) through
    forward_sequence: same beliefs and same caller-visible m_t (RAW nodes) as the step loop."""
    dev = torch.device("cuda:0")
    B, N, F_raw, F = 12, 16, 10, 32
    p, mods, pres = _pair(F, 32, spec, N, pre_dims=F_raw)
    (m_loop, _), (m_seq, _) = mods
    assert m_seq.fused_plan() is not None and m_seq.fused_plan().pre
    gen = torch.Generator().manual_seed(77)
    obs = torch.randn(B, 44, F_raw, generator=gen).to(dev)
    h_loop = h_seq = None
    o_hidden = None
    with torch.no_grad():
        for lo, hi in ((0, 3), (3, 30), (30, 44)):
            out_seq, h_seq = m_seq.forward_sequence(obs[:, lo:hi], h_seq)
            outs = []
            for t in range(lo, hi):
                o, h_loop = m_loop(obs[:, t], h_loop)
                outs.append(o)
                ref, o_hidden = oracle.dense_gcm_step(pres[0](obs[:, t]).cpu(), o_hidden, spec, p, graph_size=N)
                assert rel_err(o, ref) < 5 * TOL
            # (not bit for bit: the sequence call maps all T observations through the preprocessor in one torch Linear,
            # the per-step fast path through the library's own Linear kernel)
            assert rel_err(out_seq, torch.stack(outs, dim=1)) < TOL
        a, b = tuple(h_seq), tuple(h_loop)
        assert a[0].shape == (B, N, F_raw)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[3], b[3])
        assert torch.equal(a[0].cpu(), obs[:, 44 - N:44].cpu())            # the window of raw observations


def test_uniform_count_beyond_24_bits():
    """The uniform node count travels as its own int argument: a state whose graphs have seen more than 2^24 nodes
    steps like a young one (the kernels only use the count modulo the log / ring sizes)."""
    from gcm import _cabi

    dev = torch.device("cuda:0")
    B, N, F, hops = 64, 128, 32, (1, 2, 4)
    spec = [("temporal", hops, "forward")]
    p, mods, _ = _pair(F, 32, spec, N)
    (m_a, _), (m_b, _) = mods
    gen = torch.Generator().manual_seed(3)
    obs = torch.randn(B, 150, F, generator=gen).to(dev)
    with torch.no_grad():
        _, h_a = m_a.forward_sequence(obs[:, :140], None)
        _, h_b = m_b.forward_sequence(obs[:, :140], None)
        st = h_b.claim()
        big = 1 << 24                      # a multiple of N and of the ring: the slots stay the same
        st.count.add_(big)
        st.host_count += big
        st.max_count += big
        for t in range(140, 150):
            o_a, h_a = m_a(obs[:, t].contiguous(), h_a)
            o_b, h_b = m_b(obs[:, t].contiguous(), h_b)
            assert torch.equal(o_a, o_b), t
            assert _cabi.lib().gcm_last_kernel().decode() == "k_step_temporal_hc"
        assert int(st.count[0]) == big + 150 and st.host_count == big + 150
        o_a, h_a = m_a.forward_sequence(obs[:, :9], h_a)
        o_b, h_b = m_b.forward_sequence(obs[:, :9], h_b)
        assert torch.equal(o_a, o_b)
        assert torch.equal(tuple(h_a)[0], tuple(h_b)[0]) and torch.equal(tuple(h_a)[1], tuple(h_b)[1])
