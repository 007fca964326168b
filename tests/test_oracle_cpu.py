"""CPU tier: the oracle against (a) the golden vectors generated from the unmodified reference and
(b) the hand-written known-answer targets of the reference's own tests."""
import pytest
import torch

import gcm_oracle as oracle
import helpers
from helpers import dense_cases, load_golden, sparse_cases


@pytest.mark.parametrize("name", dense_cases())
def test_dense_oracle_matches_reference_golden(name):
    g = load_golden(name)
    hidden = None if g["init"] is None else tuple(t.clone() for t in g["init"])
    for t in range(g["T"]):
        mx, hidden = oracle.dense_gcm_step(g["obs"][t], hidden, g["spec"], g["params"], g["acts"],
                                           graph_size=g["N"])
        assert torch.allclose(mx, g["beliefs"][t], rtol=1e-5, atol=1e-6)
        if t in g["snaps"]:
            for a, b in zip(hidden, g["snaps"][t]):
                assert torch.equal(a.float(), b.float())
    for a, b in zip(hidden, g["final"]):
        assert torch.equal(a.float(), b.float())          # node slots, adjacency, num_nodes: bit-exact


@pytest.mark.parametrize("name", [n for n in dense_cases() if "grad" in n])
def test_dense_oracle_grads_match_reference(name):
    g = load_golden(name)
    obs = g["obs"].clone().requires_grad_(True)
    p = {k: v.clone().requires_grad_(True) for k, v in g["params"].items()}
    outs, hidden = oracle.dense_gcm_rollout(obs, None, g["spec"], p, g["acts"], graph_size=g["N"])
    (outs * g["loss_w"]).sum().backward()
    assert torch.allclose(obs.grad, g["d_obs"], rtol=1e-4, atol=1e-6)
    for k, v in g["d_params"].items():
        assert torch.allclose(p[k].grad, v, rtol=1e-4, atol=1e-5), k


@pytest.mark.parametrize("name", sparse_cases())
def test_sparse_oracle_matches_reference_golden(name):
    g = load_golden(name)
    hidden = None
    for (x, taus), want in zip(g["calls"], g["outs"]):
        mx, hidden = oracle.sparse_gcm_forward(x, taus, hidden, g["spec"], g["params"], g["acts"],
                                               graph_size=g["N"], max_hops=g["max_hops"],
                                               aux_selectors=g["aux"])
        assert torch.allclose(mx, want, rtol=1e-5, atol=1e-6)
    assert torch.equal(hidden[0], g["final_nodes"])
    assert torch.equal(hidden[1], g["final_edges"])
    assert torch.equal(hidden[2], g["final_T"])


# ---- known-answer targets restated from the reference's own tests -------------------------------
def _params_identity(F, rel=1.0, root=1.0):
    eye = torch.eye(F)
    z = torch.zeros(F)
    return {"w_rel1": eye * rel, "b1": z, "w_root1": eye * root, "w_rel2": eye * rel, "b2": z,
            "w_root2": eye * root}


def test_kat_wrap_overflow():
    """tests/test_gcm.py:113-184 (TestWrapOverflow): N=7, num_nodes=[1,7]; graph 1 is shifted."""
    B, N, F = 2, 7, 5
    nodes = torch.arange(B * N * F, dtype=torch.float).reshape(B, N, F)
    nodes[:, 0] = 0
    adj = torch.zeros(B, N, N)
    adj[:, 0, :] = 1
    adj[:, :, 0] = 1
    weights = torch.ones(B, N, N)
    weights[:, 0, :] = 5
    weights[:, :, 0] = 5
    obs = torch.ones(B, F) * 5
    p = oracle.make_params(F, F)
    _, (n2, a2, w2, nn2) = oracle.dense_gcm_step(obs, (nodes, adj, weights, torch.tensor([1, 7])), [], p,
                                                ("relu", "none"))
    want_adj = torch.zeros_like(adj)
    want_adj[0, 0, :] = 1
    want_adj[0, :, 0] = 1
    assert torch.equal(a2, want_adj)
    want_w = torch.ones_like(weights)
    want_w[0, 0, :] = 5
    want_w[0, :, 0] = 5
    want_w[1, -1, :] = 0
    want_w[1, :, -1] = 0
    assert torch.equal(w2, want_w)
    assert torch.equal(n2[0, 1], obs[0]) and torch.equal(n2[1, -1], obs[1])
    assert torch.equal(n2[1, 0], torch.arange(8 * 5, 9 * 5, dtype=torch.float))   # old row 1 of graph 1
    assert torch.equal(n2[1, 1], nodes[1, 2])
    assert nn2.tolist() == [2, 7]


def test_kat_direction_and_identity_e2e():
    """tests/test_gcm.py:187-323: adj[i,j]=1 pulls node j into row i; identity weights => out == obs."""
    F, N = 11, 10
    nodes = torch.arange(N * F, dtype=torch.float).reshape(1, N, F)
    adj = torch.zeros(1, N, N)
    adj[:, 0, 3] = 1
    p = _params_identity(F, rel=1.0, root=0.0)
    x = torch.ones(1, F)
    feats = oracle.dense_graph_conv(nodes.clone().index_put((torch.tensor([0]), torch.tensor([0])), x),
                                    adj, p["w_rel1"], p["b1"], p["w_root1"])
    assert torch.equal(feats[0, 0], torch.arange(3 * F, 4 * F, dtype=torch.float))
    # three steps with identity root/rel weights and no edges: belief == observation
    p = _params_identity(F)
    hidden = (torch.zeros(5, N, F), torch.zeros(5, N, N), torch.ones(5, N, N), torch.zeros(5, dtype=torch.long))
    for k in (1.0, 2.0, 3.0):
        obs = k * torch.ones(5, F)
        out, hidden = oracle.dense_gcm_step(obs, hidden, [], p, ("relu", "relu"))
        assert torch.equal(out, obs)
    assert torch.equal(hidden[0][:, 0], torch.ones(5, F))


def test_kat_temporal_far_hops():
    """tests/test_gcm.py:595-617 (TestTemporalEdge.test_far_hops): hops=[4], 10 steps, N=10."""
    B, N, F = 2, 10, 3
    p = oracle.make_params(F, F)
    hidden = (torch.arange(B * N * F, dtype=torch.float).reshape(B, N, F), torch.zeros(B, N, N),
              torch.ones(B, N, N), torch.zeros(B, dtype=torch.long))
    for _ in range(10):
        _, hidden = oracle.dense_gcm_step(torch.ones(B, F), hidden, [("temporal", (4,), "forward")], p)
    want = torch.zeros(B, N, N)
    for i in range(4, 10):
        want[:, i, i - 4] = 1
    assert torch.equal(hidden[1], want)


def test_kat_two_nodes_and_dense_edge():
    """tests/test_gcm.py:581-593 and :784-801."""
    B, N, F = 2, 10, 3
    p = oracle.make_params(F, F)
    h = None
    for _ in range(2):
        _, h = oracle.dense_gcm_step(torch.ones(B, F), h, [("temporal", (1,), "forward")], p, graph_size=N)
    want = torch.zeros(B, N, N)
    want[:, 1, 0] = 1
    assert torch.equal(h[1], want)
    h = None
    for _ in range(2):
        _, h = oracle.dense_gcm_step(torch.zeros(B, F), h, [("dense",)], p, graph_size=N)
    want = torch.zeros(B, N, N)
    want[:, :2, :2] = 1
    assert torch.equal(h[1], want)


def test_kat_distance_edges():
    """tests/test_gcm.py:708-729 (Euclidean zero/one dist) and :1135-1166 (SpatialEdge)."""
    B, N, F = 5, 10, 11
    p = oracle.make_params(F, F)
    base = (torch.zeros(B, N, F), torch.zeros(B, N, N), torch.ones(B, N, N), torch.ones(B, dtype=torch.long))
    _, h = oracle.dense_gcm_step(torch.zeros(B, F), base, [("euclidean", 1)], p)
    want = torch.zeros(B, N, N)
    want[:, 1, 0] = 1
    assert torch.equal(h[1], want)
    _, h = oracle.dense_gcm_step(torch.ones(B, F), base, [("euclidean", 1)], p)
    assert torch.equal(h[1], torch.zeros(B, N, N))
    sl = slice(0, 2)
    nodes = torch.ones(B, N, F)
    nodes[:, 0:2, sl] = 0
    _, h = oracle.dense_gcm_step(torch.zeros(B, F), (nodes, base[1], base[2], base[3]),
                                 [("spatial", 1, sl, None)], p)
    assert torch.equal(h[1], want)
    nodes = torch.zeros(B, N, F)
    nodes[:, 0, sl] = 1
    _, h = oracle.dense_gcm_step(torch.zeros(B, F), (nodes, base[1], base[2], base[3]),
                                 [("spatial", 1, sl, None)], p)
    assert torch.equal(h[1], torch.zeros(B, N, N))


def test_kat_double_edge_chain():
    """tests/test_gcm.py:631-682: chained TemporalBackedge([1]) and ([2]) compose by OR."""
    B, N, F = 5, 10, 11
    p = oracle.make_params(F, F)
    h = None
    for _ in range(4):
        _, h = oracle.dense_gcm_step(torch.zeros(B, F), h,
                                     [("temporal", (1,), "forward"), ("temporal", (2,), "forward")], p,
                                     graph_size=N)
    want = torch.zeros(B, N, N)
    for i in range(1, 4):
        want[:, i, i - 1] = 1
    for i in range(2, 4):
        want[:, i, i - 2] = 1
    assert torch.equal(h[1], want)


@pytest.mark.parametrize("name", helpers.preproc_cases())
def test_oracle_on_preprocessed_observations_reproduces_the_reference_with_a_preprocessor(name):
    """Golden vectors of the UNMODIFIED reference DenseGCM(preprocessor=Linear[, act]) (gcm.py:290-291; what RayDenseGCM
    builds): for a per-row preprocessor the beliefs are those of the plain step fed the preprocessed observations, and
    the hidden state keeps the raw observations -- the equivalence the fused preprocessor path is built on."""
    g = helpers.load_golden(name)
    pre = helpers.make_preprocessor(g)
    hidden = None
    raw = torch.zeros(g["B"], g["N"], g["F_raw"])
    with torch.no_grad():
        for t in range(g["T"]):
            mx, hidden = oracle.dense_gcm_step(pre(g["obs"][t]), hidden, g["spec"], g["params"], ("tanh", "tanh"),
                                               graph_size=g["N"])
            assert torch.allclose(mx, g["beliefs"][t], rtol=1e-5, atol=1e-6), (name, t)
            if t >= g["N"]:
                raw = torch.cat([raw[:, 1:], g["obs"][t].unsqueeze(1)], dim=1)
            else:
                raw[:, t] = g["obs"][t]
    assert torch.equal(raw, g["final"][0])                       # the reference's m_t holds RAW observations
    assert torch.equal(hidden[1].float(), g["final"][1].float()) and torch.equal(hidden[3], g["final"][3])


@pytest.mark.parametrize("name", ["pack_ragged", "pack_wide"])
def test_pack_unpack_oracle_and_cpu_helpers_match_reference_golden(name):
    """util.pack_hidden / unpack_hidden (reference util.py:323-382): the oracle's restatement and the product package's
    torch path for CPU tensors against the fixtures written by the unmodified reference."""
    from gcm import util

    g = load_golden(name)
    adj = torch.sparse_coo_tensor(g["indices"], g["values"], size=(g["B"], g["N"], g["N"]))
    for impl in (oracle, util):
        _, edges, weights, _ = impl.pack_hidden((g["nodes"], adj, g["T"]), g["B"], g["max_edges"])
        assert torch.equal(edges, g["edges"]) and torch.equal(weights, g["weights"]), impl.__name__
        _, adj2, _ = impl.unpack_hidden((g["nodes"], g["edges"], g["weights"], g["T"]), g["B"])
        assert torch.equal(adj2._indices(), g["unpacked_indices"]) and torch.equal(adj2._values(), g["unpacked_values"])
    with pytest.raises(AssertionError, match="Cannot pack"):
        oracle.pack_hidden((g["nodes"], adj, g["T"]), g["B"], 3)
