"""DenseEdge-only fast path of DenseGCM (csrc/gcm_dense_ones.cu; include/gcm_b200.h "ones" section).

On a state whose adjacency is the all-ones block over the valid nodes (built from empty by DenseEdge alone, or
ingested from a caller tuple that has exactly that adjacency) the reference's step (gcm.py:262-321 with the
DenseGraphConv stack of README.md:52-62) reduces to a per-graph running sum, two small projections and ONE
streaming pass over a per-node cache R_i = W_root1 x_i; see the header of the .cu file for the algebra.  This
module owns the host side: the extra state buffers, the autograd functions (one node per step, newest-first
through the token chain of gcm.fused) and the hand-over to the general kernels when a configuration leaves
the path.  CUDA only; every product is a kernel behind the C ABI.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from gcm import _cabi


def _lin2(a1, w1, a2=None, w2=None, bias=None, act=0, out=None, status=None, accumulate=False):
    """out = act(a1 @ w1.T + a2 @ w2.T + bias) through gcm_linear2.  a1 [rows, K1] (row stride = stride(0))."""
    rows, k1 = a1.shape
    ho = w1.shape[0]
    if out is None:
        out = torch.empty(rows, ho, device=a1.device, dtype=torch.float32)
    k2 = 0 if a2 is None else a2.shape[1]
    _cabi.check(_cabi.lib().gcm_linear2(
        a1.data_ptr(), k1, a1.stride(0), w1.data_ptr(),
        None if a2 is None else a2.data_ptr(), k2, 0 if a2 is None else a2.stride(0),
        None if w2 is None else w2.data_ptr(), None if bias is None else bias.data_ptr(), act, rows, ho,
        out.data_ptr(), out.stride(0), None if status is None else status.data_ptr(), int(accumulate),
        _cabi.stream_ptr(a1.device)), "gcm_linear2")
    return out


def _outer(a, x, dw, db=None):
    """dw += a.T @ x, db += a.sum(0) through gcm_outer_reduce."""
    _cabi.check(_cabi.lib().gcm_outer_reduce(a.data_ptr(), a.stride(0), a.shape[1], x.data_ptr(), x.stride(0),
                                             x.shape[1], a.shape[0], dw.data_ptr(),
                                             None if db is None else db.data_ptr(), _cabi.stream_ptr(a.device)),
                "gcm_outer_reduce")


def plan_supports(plan) -> bool:
    g = plan.gnn
    return (bool(plan.sels) and all(s.kind == _cabi.SEL_DENSE for s in plan.sels)
            and max(g.F, g.H1, g.H2) <= 128 and g.F % 4 == 0)


def _weights(plan, dev):
    plan.gnn.packed(dev)
    return plan.gnn._packed[1]


def _root_key(plan):
    w = plan.gnn.conv1.lin_root._parameters["weight"]
    return (w.data_ptr(), w._version)


def prepare(plan, state) -> None:
    """Bring the extra buffers of the path up to date: S from the log, R for every row under the current W_root1."""
    dev = state.device
    w = _weights(plan, dev)
    lib = _cabi.lib()
    if state.xsum is None:
        state.xsum = torch.empty(state.B, state.F, device=dev, dtype=torch.float32)
        _cabi.check(lib.gcm_dense_ones_xsum(state.c_ref(), state.xsum.data_ptr(), _cabi.stream_ptr(dev)),
                    "gcm_dense_ones_xsum")
    key = _root_key(plan)
    if state.rcache is None or state.rc_key != key:
        if state.rcache is None:
            state.rcache = torch.empty(state.B, state.C, plan.gnn.H1, device=dev, dtype=torch.float32)
        _lin2(state.nodes.view(state.B * state.C, state.F), w["w_root1"],
              out=state.rcache.view(state.B * state.C, plan.gnn.H1))
        state.rc_key = key


def _forward_kernels(plan, state, x):
    """One step on the in-place state.  Returns (belief, c, G, h_t)."""
    dev = state.device
    lib = _cabi.lib()
    g = plan.gnn
    w = _weights(plan, dev)
    stream = _cabi.stream_ptr(dev)
    _cabi.check(lib.gcm_dense_ones_update(state.c_ref(), x.data_ptr(), state.xsum.data_ptr(), stream),
                "gcm_dense_ones_update")
    c = _lin2(state.xsum, w["w_rel1"], bias=w["b1"])
    r_t = _lin2(x, w["w_root1"])
    G = torch.empty(state.B, g.H1, device=dev, dtype=torch.float32)
    h_t = torch.empty(state.B, g.H1, device=dev, dtype=torch.float32)
    _cabi.check(lib.gcm_dense_ones_stream_fwd(state.c_ref(), g.H1, _cabi.ACT[g.act1], state.rcache.data_ptr(),
                                              c.data_ptr(), r_t.data_ptr(), G.data_ptr(), h_t.data_ptr(), stream),
                "gcm_dense_ones_stream_fwd")
    belief = _lin2(G, w["w_rel2"], h_t, w["w_root2"], bias=w["b2"], act=_cabi.ACT[g.act2], status=state.status)
    state.masks_stale = True
    state.version += 1
    state.steps += 1
    state.max_count += 1
    if state.host_count is not None:
        state.host_count += 1
    return belief, c, G, h_t


def step_nograd(plan, state, x):
    prepare(plan, state)
    return _forward_kernels(plan, state, x)[0]


class _OnesRootFn(torch.autograd.Function):
    """Start of a recorded chain on the ones path.  Runs LAST in backward: the products of the accumulated
    per-node dL/d(pre-activation) with the node rows are linear, so they are applied here once per chain:
    dW_root1 = sum_{b,i} DZ_i x_i^T."""

    @staticmethod
    def forward(ctx, anchor, w_root1, state):
        ctx.state = state
        return anchor.clone()

    @staticmethod
    def backward(ctx, d_token):
        st = ctx.state
        dw = None
        if st.DZ is not None:
            H1 = st.DZ.shape[-1]
            dw = torch.zeros(H1, st.F, device=st.device, dtype=torch.float32)
            _outer(st.DZ.view(st.B * st.C, H1), st.nodes.view(st.B * st.C, st.F), dw)
            st.DZ.zero_()
            st.ds_run.zero_()
            st.ds_snap.clear()
        return torch.zeros_like(d_token), dw, None


class _OnesStepFn(torch.autograd.Function):
    """One step of the ones path.  Saved: S, c, G, h_t and the belief of the step ([B, F|H] each); the per-node
    work of the backward is recomputed from the R cache."""

    @staticmethod
    def forward(ctx, x, token, plan, state, *params):
        belief, c, G, h_t = _forward_kernels(plan, state, x.detach())
        ctx.plan, ctx.state = plan, state
        ctx.step_index = state.steps
        ctx.pkey = plan.gnn._key
        ctx.packed = plan.gnn._packed
        ctx.save_for_backward(state.xsum.clone(), c, G, h_t, belief)
        return belief, torch.zeros(1, device=state.device)

    @staticmethod
    def backward(ctx, d_belief, d_token):
        plan, st = ctx.plan, ctx.state
        g = plan.gnn
        dev = st.device
        if g.current_key(dev) != ctx.pkey:
            raise RuntimeError("GNN parameters were modified in place between forward and backward")
        S, c, G, h_t, belief = ctx.saved_tensors
        w = ctx.packed[1]
        steps_back = st.steps - ctx.step_index
        if steps_back > st.C - st.N:
            raise RuntimeError(
                f"BPTT window too long for the node log: this step is {steps_back} steps old but the log "
                f"keeps {st.C - st.N} spare rows; raise DenseGCM.bptt_capacity")
        if st.DZ is None:
            st.DZ = torch.zeros(st.B, st.C, g.H1, device=dev, dtype=torch.float32)
            st.ds_run = torch.zeros(st.B, st.F, device=dev, dtype=torch.float32)
        tw = plan.gnn.transposed(dev)
        db_ = d_belief.contiguous().float()
        if g.act2 == "tanh":
            do = db_ * (1.0 - belief * belief)
        elif g.act2 == "relu":
            do = db_ * (belief > 0).to(db_.dtype)
        else:
            do = db_
        grads = {
            "w_rel1": torch.zeros(g.H1, g.F, device=dev), "b1": torch.zeros(g.H1, device=dev),
            "w_rel2": torch.zeros(g.H2, g.H1, device=dev), "w_root2": torch.zeros(g.H2, g.H1, device=dev),
            "b2": torch.zeros(g.H2, device=dev),
        }
        _outer(do, G, grads["w_rel2"], grads["b2"])
        _outer(do, h_t, grads["w_root2"])
        dG = _lin2(do, tw["w_rel2_t"])                       # [B, H1] = do W_rel2
        dh_t = _lin2(do, tw["w_root2_t"])
        dc = torch.empty(st.B, g.H1, device=dev, dtype=torch.float32)
        dz_t = torch.empty(st.B, g.H1, device=dev, dtype=torch.float32)
        _cabi.check(_cabi.lib().gcm_dense_ones_stream_bwd(
            st.c_ref(), steps_back, g.H1, _cabi.ACT[g.act1], st.rcache.data_ptr(), c.data_ptr(), dG.data_ptr(),
            dh_t.data_ptr(), st.DZ.data_ptr(), dc.data_ptr(), dz_t.data_ptr(), _cabi.stream_ptr(dev)),
            "gcm_dense_ones_stream_bwd")
        _outer(dc, S, grads["w_rel1"], grads["b1"])
        # dL/dS of this step reaches every node of its window: running sum over the later steps
        _lin2(dc, tw["w_rel1_t"], out=st.ds_run, accumulate=True)
        if st.max_count > st.N:
            st.ds_snap[ctx.step_index] = st.ds_run.clone()
        d_x = st.ds_run.clone()
        gone = st.ds_snap.get(ctx.step_index + st.N)          # steps after this node left the window
        if gone is not None:
            d_x -= gone
        _lin2(dz_t, tw["w_root1_t"], out=d_x, accumulate=True)
        out = []
        for conv, wr, wo, bb in ((g.conv1, "w_rel1", None, "b1"), (g.conv2, "w_rel2", "w_root2", "b2")):
            out.append(grads[wr])
            if conv.lin_rel.bias is not None:
                out.append(grads[bb])
            out.append(None if wo is None else grads[wo])     # dW_root1 is applied once, by _OnesRootFn
            if conv.lin_root.bias is not None:
                out.append(grads[bb])
        return (d_x, torch.zeros(1, device=dev), None, None, *out)


def step_grad(plan, state, x, token):
    """Recording step.  Returns (belief, token)."""
    prepare(plan, state)
    if token is None:
        anchor = torch.zeros(1, device=state.device, requires_grad=True)
        token = _OnesRootFn.apply(anchor, plan.gnn.conv1.lin_root.weight, state)
        state.chain_start = state.steps
    elif state.steps + 1 - getattr(state, "chain_start", 0) > state.C - state.N + 1:
        raise RuntimeError(
            f"more than {state.C - state.N + 1} recorded steps on one hidden state; raise "
            "DenseGCM.bptt_capacity or cut the graph with m_t.detach()")
    belief, token = _OnesStepFn.apply(x, token, plan, state, *plan.gnn.params())
    return belief, token
