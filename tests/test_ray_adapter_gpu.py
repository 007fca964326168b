"""GPU tier: the reference's OWN RLlib adapter (RayDenseGCM, /root/reference/src/gcm/ray_gcm.py:21-212) runs unchanged
on top of this package (BASELINE.json north_star): the file is executed from where it lies (the reference
checkout in the build container, baseline/_ref on the GPU box -- never copied into the repository) against stand-ins
for gym / ray / torch_geometric (tests/rllib_stubs.py).  Its forward -- Linear preprocessor, `for t in range(T):
out, hidden = self.gcm(flat[:, t, :], hidden)`, `state = list(hidden)` -- is checked against the oracle."""
import pytest
import torch

import gcm_oracle as oracle
import rllib_stubs
from helpers import make_dense_gnn, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(rllib_stubs.reference_adapter_path() is None, reason="reference ray_gcm.py not available")
@pytest.mark.parametrize("selector", ["temporal", "dense"])
def test_reference_ray_adapter_runs_on_this_package(selector):
    import gcm.gcm
    from gcm.edge_selectors.dense import DenseEdge
    from gcm.edge_selectors.temporal import TemporalBackedge
    from gcm.state import DenseHidden

    ref = rllib_stubs.load_reference_adapter()
    assert ref.DenseGCM is gcm.gcm.DenseGCM, "the adapter must have imported THIS package's DenseGCM"
    dev = torch.device("cuda:0")
    B, T, obs_dim, F, H, N, n_out = 5, 7, 6, 32, 32, 8, 3
    p = oracle.make_params(F, H)
    gnn, _ = make_dense_gnn(F, H, p, ("tanh", "tanh"), "sequential")
    sel = TemporalBackedge([1, 2]) if selector == "temporal" else DenseEdge()
    spec = [("temporal", (1, 2), "forward")] if selector == "temporal" else [("dense",)]
    torch.manual_seed(11)
    model = ref.RayDenseGCM(rllib_stubs.Space(obs_dim), rllib_stubs.Space(2), n_out, {}, "gcm", gnn=gnn,
                            edge_selectors=sel, graph_size=N, gnn_input_size=F, gnn_output_size=H).to(dev)
    assert model.gcm.fused_plan() is not None and model.gcm.fused_plan().pre, "RayDenseGCM's DenseGCM should be fusable"
    state = [s.unsqueeze(0).repeat(B, *([1] * s.dim())).to(dev) for s in model.get_initial_state()]
    assert state[0].shape == (B, N, obs_dim) and state[3].shape == (B,)
    gen = torch.Generator().manual_seed(5)
    pre = model.gcm.preprocessor
    o_hidden, raw = None, []
    with torch.no_grad():
        for call in range(3):                       # 21 steps on N = 8: the window wraps; the state list round-trips
            obs = torch.randn(B, T, obs_dim, generator=gen)
            logits, state = model.forward({"obs_flat": obs.reshape(B * T, obs_dim).to(dev)}, state, torch.full((B,), T))
            assert logits.shape == (B * T, n_out) and isinstance(state, list) and len(state) == 4
            assert all(isinstance(s, torch.Tensor) and not isinstance(s, DenseHidden) for s in state)
            beliefs = []
            for t in range(T):
                y = pre(obs[:, t].to(dev)).cpu()
                b, o_hidden = oracle.dense_gcm_step(y, o_hidden, spec, p, graph_size=N)
                beliefs.append(b)
                raw.append(obs[:, t])
            want = model.logit_branch(torch.stack(beliefs, dim=1).reshape(B * T, H).to(dev))
            assert rel_err(logits, want) < 5e-5, call
            assert rel_err(model.value_function(),
                           model.value_branch(torch.stack(beliefs, dim=1).reshape(B * T, H).to(dev)).squeeze(1)) < 5e-5
            # the state RLlib carries on: RAW observations in window order, the selector's adjacency, num_nodes
            n_seen = len(raw)
            window = torch.stack(raw[max(0, n_seen - N):], dim=1)
            assert torch.equal(state[0][:, : window.shape[1]].cpu(), window)
            assert torch.equal(state[1].cpu(), o_hidden[1]) and torch.equal(state[3].cpu(), o_hidden[3])
    # a training forward of the adapter (autograd on): gradients reach the heads, the GNN and the Linear preprocessor
    obs = torch.randn(B, T, obs_dim, generator=gen)
    logits, state2 = model.forward({"obs_flat": obs.reshape(B * T, obs_dim).to(dev)}, state, torch.full((B,), T))
    (logits.pow(2).sum() + model.value_function().sum()).backward()
    for prm in list(model.gcm.preprocessor.parameters()) + list(model.gcm.gnn.parameters()) + list(model.logit_branch.parameters()):
        assert prm.grad is not None and bool(torch.isfinite(prm.grad).all()) and float(prm.grad.abs().max()) > 0
    assert all(isinstance(s, torch.Tensor) for s in state2) and state2[3].dtype == torch.long
