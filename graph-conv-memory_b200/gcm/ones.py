"""DenseEdge-only fast path of DenseGCM (csrc/gcm_dense_ones.cu; include/gcm_b200.h "ones" section).

On a state whose adjacency is the all-ones block over the valid nodes (built from empty by DenseEdge alone, or
ingested from a caller tuple that has exactly that adjacency) the reference's step (gcm.py:262-321 with the
DenseGraphConv stack of README.md:52-62) reduces to a per-graph running sum, two small projections and ONE
streaming pass over a per-node cache (a function of R_i = W_root1 x_i); see the header of the .cu file for the
algebra.  This module owns the host side: the extra state buffers, the per-step buffers of a BPTT window, the
autograd functions and the hand-over to the general kernels when a configuration leaves the path.  CUDA only;
every product is a kernel behind the C ABI.

Autograd: GCM has no recurrence through the belief (the hidden state is the observation log), so once
`backward()` runs, dL/dbelief of every step is known independently of the other steps.  A step's backward node
therefore only does [B, H]-sized work (dG, dc, and dL/dx_k when the observation requires grad); the pass over the
per-node cache and every weight gradient are done ONCE per window by the chain's root node, which autograd runs
last (every step depends on it through the token chain of gcm.fused).
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


def _lin_tc32(a1, w1, a2=None, w2=None, bias=None, act=0, out=None, status=None):
    """out = act(a1 @ w1.T [3xTF32] + a2 @ w2.T [bf16] + bias) through gcm_linear_tc32 (tcgen05)."""
    rows, k1 = a1.shape
    ho = w1.shape[0]
    if out is None:
        out = torch.empty(rows, ho, device=a1.device, dtype=torch.float32)
    _cabi.check(_cabi.lib().gcm_linear_tc32(
        a1.data_ptr(), k1, a1.stride(0), w1.data_ptr(),
        None if a2 is None else a2.data_ptr(), 0 if a2 is None else a2.shape[1], 0 if a2 is None else a2.stride(0),
        None if w2 is None else w2.data_ptr(), None if bias is None else bias.data_ptr(), act, rows, ho,
        out.data_ptr(), out.stride(0), None if status is None else status.data_ptr(),
        _cabi.stream_ptr(a1.device)), "gcm_linear_tc32")
    return out


def _outer(a, x, dw, db=None):
    """dw += a.T @ x, db += a.sum(0) through gcm_outer_reduce."""
    _cabi.check(_cabi.lib().gcm_outer_reduce(a.data_ptr(), a.stride(0), a.shape[1], x.data_ptr(), x.stride(0),
                                             x.shape[1], a.shape[0], dw.data_ptr(),
                                             None if db is None else db.data_ptr(), _cabi.stream_ptr(a.device)),
                "gcm_outer_reduce")


_WS = {}


def _outer_tc32(a, x, dw, db=None):
    """dw += a.T @ x, db += a.sum(0) in 3xTF32 on the tensor cores (fp32-accurate, deterministic)."""
    return _outer_tc(a, x, dw, db, fn="gcm_outer_reduce_tc32")


def _outer_tc(a, x, dw, db=None, fn="gcm_outer_reduce_tc"):
    """dw += a.T @ x, db += a.sum(0) on the bf16 tensor-core path (gcm_outer_reduce_tc; deterministic)."""
    lib = _cabi.lib()
    rows = a.shape[0]
    need = int(lib.gcm_outer_reduce_tc_workspace(rows))
    ws = _WS.get(a.device)
    if ws is None or ws.numel() < need:
        ws = _WS[a.device] = torch.empty(need, device=a.device, dtype=torch.float32)
    _cabi.check(getattr(lib, fn)(a.data_ptr(), a.stride(0), a.shape[1], x.data_ptr(), x.stride(0), x.shape[1],
                                 rows, ws.data_ptr(), dw.data_ptr(), None if db is None else db.data_ptr(),
                                 _cabi.stream_ptr(a.device)), fn)


def plan_supports(plan) -> bool:
    """A selector chain that CONTAINS a DenseEdge builds the all-ones valid block whatever else is chained with it:
    DenseEdge links every new node to every node of the window in both directions, with self loops
    (edge_selectors/dense.py:16-21), and the other selectors only ever OR in edges between valid nodes (SURVEY.md
    checklist item 8), so their edges are already there.  Such chains take this path too (the distance selectors of the
    chain are not even evaluated: they cannot change the adjacency)."""
    g = plan.gnn
    known = (_cabi.SEL_TEMPORAL, _cabi.SEL_DENSE, _cabi.SEL_EUCLIDEAN, _cabi.SEL_COSINE, _cabi.SEL_SPATIAL)
    return (bool(plan.sels) and any(s.kind == _cabi.SEL_DENSE for s in plan.sels)
            and all(s.kind in known for s in plan.sels)
            and max(g.F, g.H1, g.H2) <= 128 and g.F % 4 == 0 and g.H1 % 4 == 0)


def want_bf16(mod, plan) -> bool:
    """bfloat16 per-node cache (BASELINE cfg3's precision: beliefs / gradients within 2e-2 instead of 1e-5):
    `DenseGCM.compute_dtype = torch.bfloat16`, or a surrounding torch.autocast("cuda", dtype=torch.bfloat16)."""
    if plan.gnn.H1 % 8:
        return False
    if getattr(mod, "compute_dtype", None) == torch.bfloat16:
        return True
    return torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16


def _weights(plan, dev):
    plan.gnn.packed(dev)
    return plan.gnn._packed[1]


def _cache_act(plan) -> int:
    """epilogue that turns c / R into what the cache passes read: exp(2 clamp(.)) for tanh, identity otherwise"""
    return _cabi.ACT_EXP2X if plan.gnn.act1 == "tanh" else 0


def prepare(plan, state, bf16: bool = False) -> None:
    """Bring the extra buffers of the path up to date: S from the log, the cache of every row under the current
    W_root1 (and activation / element type)."""
    dev = state.device
    w = _weights(plan, dev)
    lib = _cabi.lib()
    if state.xsum is None:
        state.xsum = torch.empty(state.B, state.F, device=dev, dtype=torch.float32)
        _cabi.check(lib.gcm_dense_ones_xsum(state.c_ref(), state.xsum.data_ptr(), _cabi.stream_ptr(dev)),
                    "gcm_dense_ones_xsum")
    wr = plan.gnn.conv1.lin_root._parameters["weight"]
    key = (wr.data_ptr(), wr._version, plan.gnn.act1, bool(bf16))
    if state.rcache is None or state.rc_key != key:
        H1 = plan.gnn.H1
        rows = state.B * state.C
        if bf16 and state.F % 16 == 0 and H1 % 16 == 0:
            # bf16 tensor cores (tcgen05): the rounding of R_i is independent per node, like the cache's own
            state.rcache = torch.empty(state.B, state.C, H1, device=dev, dtype=torch.bfloat16)
            _cabi.check(lib.gcm_linear_tc(state.nodes.data_ptr(), state.F, state.F, w["w_root1"].data_ptr(), None,
                                          _cache_act(plan), rows, H1, state.rcache.data_ptr(), H1, 1,
                                          _cabi.stream_ptr(dev)), "gcm_linear_tc")
        else:
            lin = _lin_tc32 if (state.F % 16 == 0 and H1 % 16 == 0) else _lin2      # float32 cache: 3xTF32
            full = lin(state.nodes.view(rows, state.F), w["w_root1"], act=_cache_act(plan))
            if bf16:
                state.rcache = torch.empty(state.B, state.C, H1, device=dev, dtype=torch.bfloat16)
                _cabi.check(lib.gcm_to_bf16(full.data_ptr(), state.rcache.data_ptr(), rows * H1,
                                            _cabi.stream_ptr(dev)), "gcm_to_bf16")
            else:
                state.rcache = full.view(state.B, state.C, H1)
        state.rc_key = key
        state.rc_bf16 = bool(bf16)
        state.ones_tmp = None


class _Window:
    """Per-step buffers of one BPTT window, [K, B, .] float32 each, so that the window-level kernels address
    step k at a fixed stride.  Forward: S_k, E_k (= exp(2 c_k) or c_k), G_k, P_k, h_t,k.  Backward: dL/dbelief
    through act2 (do), dG, the own-node term (dzo), dc and - when observations require grad - its suffix sums."""

    FWD = (("S", "F"), ("E", "H1"), ("G", "H1"), ("P", "H1"), ("ht", "H1"))
    BWD = (("do", "H2"), ("dG", "H1"), ("dzo", "H1"), ("dc", "H1"))

    def __init__(self, state, g):
        self.B, self.dev = state.B, state.device
        self.dims = {"F": g.F, "H1": g.H1, "H2": g.H2}
        self.K = 0
        self.Kb = 0
        self.chain_id = 0
        self.chain_start = 0
        self.kmax = -1          # newest-first backward: steps kmax .. k have delivered gradients so far
        self.need_dx = False
        self.dcs = None

    def _alloc(self, K, h):
        return torch.empty(K, self.B, self.dims[h], device=self.dev, dtype=torch.float32)

    def ensure_fwd(self, k, cap):
        if k < self.K:
            return
        K = min(cap, max(64, 2 * self.K, k + 1))
        for name, h in self.FWD:
            new = self._alloc(K, h)
            if self.K:
                new[: self.K].copy_(getattr(self, name))
            setattr(self, name, new)
        self.K = K

    def ensure_bwd(self):
        if self.Kb < self.K:
            for name, h in self.BWD:
                setattr(self, name, self._alloc(self.K, h))
            self.dcs = None
            self.Kb = self.K
        if self.need_dx and self.dcs is None:
            self.dcs = self._alloc(self.K, "H1")


def _tmp(state, g):
    t = getattr(state, "ones_tmp", None)
    if t is None:
        mk = lambda h: torch.empty(state.B, h, device=state.device, dtype=torch.float32)
        t = state.ones_tmp = {"E": mk(g.H1), "q": mk(g.H1), "G": mk(g.H1), "ht": mk(g.H1), "S": mk(g.F)}
    return t


def _forward_kernels(plan, state, x, k: Optional[int] = None):
    """One step on the in-place state; k = index of the step in the recording window (None: not recording)."""
    dev = state.device
    lib = _cabi.lib()
    g = plan.gnn
    w = _weights(plan, dev)
    stream = _cabi.stream_ptr(dev)
    tmp = _tmp(state, g)
    win = state.win if k is not None else None
    if win is not None:
        S, E, G, P, ht = win.S[k], win.E[k], win.G[k], win.P[k], win.ht[k]
    else:
        S, E, G, P, ht = tmp["S"], tmp["E"], tmp["G"], None, tmp["ht"]
    _cabi.check(lib.gcm_dense_ones_update(state.c_ref(), x.data_ptr(), state.xsum.data_ptr(), S.data_ptr(), stream),
                "gcm_dense_ones_update")
    state.xsum = S
    ca = _cache_act(plan)
    tc_dims = g.F % 16 == 0 and g.H1 % 16 == 0 and g.H2 % 16 == 0
    if tc_dims:
        # c = W_rel1 S + b1 is common to every node of the graph: 3xTF32 (fp32-accurate) on the tensor cores
        _lin_tc32(S, w["w_rel1"], bias=w["b1"], act=ca, out=E)
    else:
        _lin2(S, w["w_rel1"], bias=w["b1"], act=ca, out=E)
    if state.rc_bf16 and g.F % 16 == 0 and g.H1 % 16 == 0:
        # the new node's cache row on the bf16 tensor cores, like the rows written by prepare()
        _cabi.check(lib.gcm_linear_tc(x.data_ptr(), g.F, x.stride(0), w["w_root1"].data_ptr(), None, ca, state.B, g.H1,
                                      tmp["q"].data_ptr(), g.H1, 0, stream), "gcm_linear_tc")
    elif tc_dims:
        _lin_tc32(x, w["w_root1"], act=ca, out=tmp["q"])
    else:
        _lin2(x, w["w_root1"], act=ca, out=tmp["q"])
    _cabi.check(lib.gcm_dense_ones_fwd(state.c_ref(), g.H1, _cabi.ACT[g.act1], int(state.rc_bf16),
                                       state.rcache.data_ptr(), E.data_ptr(), tmp["q"].data_ptr(), G.data_ptr(),
                                       None if P is None else P.data_ptr(), ht.data_ptr(), stream),
                "gcm_dense_ones_fwd")
    if tc_dims and state.rc_bf16:
        # G W_rel2^T in 3xTF32 (G sums up to N rows), h_t W_root2^T in bf16 (|h_t| < 1, per-graph rounding only)
        belief = _lin_tc32(G, w["w_rel2"], ht, w["w_root2"], bias=w["b2"], act=_cabi.ACT[g.act2], status=state.status)
    else:
        belief = _lin2(G, w["w_rel2"], ht, w["w_root2"], bias=w["b2"], act=_cabi.ACT[g.act2], status=state.status)
    state.masks_stale = True
    state.fast_ok = False
    state.version += 1
    state.steps += 1
    state.max_count += 1
    if state.host_count is not None:
        state.host_count += 1
    return belief


def time_fwd_kernel(plan, state, iters: int = 20) -> float:
    """Average duration (ms, CUDA events on the current stream) of back-to-back k_ones_fwd launches on the state as it
    is (the kernel re-derives everything from the counters, so repeating it changes nothing).  For bench.py."""
    g = plan.gnn
    tmp = _tmp(state, g)
    lib = _cabi.lib()
    stream = _cabi.stream_ptr(state.device)
    P = torch.empty_like(tmp["G"])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for i in range(iters + 3):
        if i == 3:
            ev[0].record()
        _cabi.check(lib.gcm_dense_ones_fwd(state.c_ref(), g.H1, _cabi.ACT[g.act1], int(state.rc_bf16),
                                           state.rcache.data_ptr(), tmp["E"].data_ptr(), tmp["q"].data_ptr(),
                                           tmp["G"].data_ptr(), P.data_ptr(), tmp["ht"].data_ptr(), stream),
                    "gcm_dense_ones_fwd")
    ev[1].record()
    ev[1].synchronize()
    return ev[0].elapsed_time(ev[1]) / iters


def step_nograd(plan, state, x, bf16: bool = False):
    prepare(plan, state, bf16)
    return _forward_kernels(plan, state, x)


def sequence_supported(plan, state, T: int, bf16: bool) -> bool:
    """Can T steps be taken at once (gcm_dense_ones_window_fwd keeps N + T cache rows in shared memory)?"""
    g = plan.gnn
    if g.H1 % (8 if bf16 else 4):
        return False
    return int(_cabi.lib().gcm_dense_ones_seq_smem(state.N, T, g.H1, int(bool(bf16)))) <= 220 * 1024


def _sequence_kernels(plan, state, x_seq, k0: Optional[int] = None):
    """T steps on the in-place state from x_seq [B, T, F] (contiguous); k0 = index of the first step in the recording
    window (None: not recording).  Returns beliefs [T, B, H2].  Same arithmetic per step as _forward_kernels."""
    dev = state.device
    lib = _cabi.lib()
    g = plan.gnn
    w = _weights(plan, dev)
    stream = _cabi.stream_ptr(dev)
    B, T, F = x_seq.shape
    if k0 is not None:
        win = state.win
        S, E, G, P, ht = (getattr(win, n)[k0:k0 + T] for n in ("S", "E", "G", "P", "ht"))
    else:
        mk = lambda h: torch.empty(T, B, h, device=dev, dtype=torch.float32)
        S, E, G, P, ht = mk(F), mk(g.H1), mk(g.H1), None, mk(g.H1)
    _cabi.check(lib.gcm_dense_ones_seq_update(state.c_ref(), x_seq.data_ptr(), x_seq.stride(0), x_seq.stride(1), T,
                                              state.xsum.data_ptr(), S.data_ptr(), B * F, stream),
                "gcm_dense_ones_seq_update")
    state.xsum = S[T - 1]
    ca = _cache_act(plan)
    tc_dims = g.F % 16 == 0 and g.H1 % 16 == 0 and g.H2 % 16 == 0
    lin_c = _lin_tc32 if tc_dims else _lin2
    lin_c(S.view(T * B, F), w["w_rel1"], bias=w["b1"], act=ca, out=E.view(T * B, g.H1))
    q_new = torch.empty(B * T, g.H1, device=dev, dtype=torch.float32)
    xf = x_seq.view(B * T, F)
    if state.rc_bf16 and tc_dims:
        _cabi.check(lib.gcm_linear_tc(xf.data_ptr(), F, F, w["w_root1"].data_ptr(), None, ca, B * T, g.H1,
                                      q_new.data_ptr(), g.H1, 0, stream), "gcm_linear_tc")
    else:
        _lin2(xf, w["w_root1"], act=ca, out=q_new)
    # chunks of the window small enough for three resident CTAs per SM (cache rows of N + chunk nodes in shared memory),
    # when that is possible; steps are independent given the cache, so the chunks are just consecutive launches
    chunk = T
    smem = lambda t: int(lib.gcm_dense_ones_seq_smem(state.N, t, g.H1, int(state.rc_bf16)))
    if smem(T) > 74 * 1024:
        fit = [t for t in range(min(T - 1, 64), 7, -8) if smem(t) <= 74 * 1024]
        if fit:
            n_chunks = -(-T // fit[0])
            chunk = -(-T // n_chunks)            # balanced: 63 steps -> 32 + 31, not 48 + 15
    for k_start in range(0, T, chunk):
        tc_ = min(chunk, T - k_start)
        _cabi.check(lib.gcm_dense_ones_window_fwd(
            state.c_ref(), g.H1, _cabi.ACT[g.act1], int(state.rc_bf16), state.rcache.data_ptr(), tc_, T - k_start - tc_,
            E[k_start].data_ptr(), q_new.data_ptr() + 4 * k_start * g.H1, T * g.H1, G[k_start].data_ptr(),
            None if P is None else P[k_start].data_ptr(), ht[k_start].data_ptr(), B * g.H1, stream),
            "gcm_dense_ones_window_fwd")
    lin_b = _lin_tc32 if (tc_dims and state.rc_bf16) else _lin2
    beliefs = lin_b(G.view(T * B, g.H1), w["w_rel2"], ht.view(T * B, g.H1), w["w_root2"], bias=w["b2"],
                    act=_cabi.ACT[g.act2], status=state.status)
    state.masks_stale = True
    state.version += 1
    state.steps += T
    state.max_count += T
    if state.host_count is not None:
        state.host_count += T
    return beliefs.view(T, B, g.H2)


def sequence_nograd(plan, state, x_seq, bf16: bool = False):
    prepare(plan, state, bf16)
    return _sequence_kernels(plan, state, x_seq)


def _param_grads(g, grads):
    """gradients in the order of GnnPlan.params()"""
    out = []
    for conv, wr, wo, bb in ((g.conv1, "w_rel1", "w_root1", "b1"), (g.conv2, "w_rel2", "w_root2", "b2")):
        out.append(grads[wr])
        if conv.lin_rel.bias is not None:
            out.append(grads[bb])
        out.append(grads[wo])
        if conv.lin_root.bias is not None:
            out.append(grads[bb])
    return out


def _lin_bwd(st, a, wt, out):
    """out = a @ wt.T for the backward products dL/dG = do W_rel2, dL/dh_t = do W_root2 ([rows, H2] x [H2, H1]);
    bf16 tensor cores when the state computes in bfloat16 (per-sample rounding, gradients only)."""
    rows, k = a.shape
    ho = wt.shape[0]
    if st.rc_bf16 and k % 16 == 0 and ho % 16 == 0:
        _cabi.check(_cabi.lib().gcm_linear_tc(a.data_ptr(), k, a.stride(0), wt.data_ptr(), None, 0, rows, ho,
                                              out.data_ptr(), out.stride(0), 0, _cabi.stream_ptr(a.device)),
                    "gcm_linear_tc")
    elif k % 16 == 0 and ho % 16 == 0:
        _lin_tc32(a, wt, out=out)
    else:
        _lin2(a, wt, out=out)


def _bwd_small(plan, st, k0, k1, have_next):
    """The [B, H]-sized part of the backward of window steps k0 .. k1-1 (stacked rows): dG, dzo, dc (and the suffix
    sums of dc when observations require grad; then k1 == k0 + 1)."""
    g, win = plan.gnn, st.win
    tw = g.transposed(st.device)
    rows = (k1 - k0) * st.B
    do = win.do[k0:k1].view(rows, g.H2)
    dG = win.dG[k0:k1].view(rows, g.H1)
    dzo = win.dzo[k0:k1].view(rows, g.H1)
    _lin_bwd(st, do, tw["w_rel2_t"], dG)                        # dL/dG = do W_rel2
    _lin_bwd(st, do, tw["w_root2_t"], dzo)                      # dL/dh_t through lin_root2
    _cabi.check(_cabi.lib().gcm_dense_ones_dc(
        dG.data_ptr(), dzo.data_ptr(), win.P[k0:k1].data_ptr(), win.ht[k0:k1].data_ptr(), _cabi.ACT[g.act1],
        rows * g.H1, win.dc[k0:k1].data_ptr(), win.dcs[k0 + 1].data_ptr() if have_next else None,
        win.dcs[k0].data_ptr() if win.need_dx else None, _cabi.stream_ptr(st.device)), "gcm_dense_ones_dc")


class _OnesRootFn(torch.autograd.Function):
    """Start of a recorded chain on the ones path.  Runs LAST in backward, when every step of the window has
    delivered dG_k / dzo_k / dc_k: ONE pass over the per-node cache gives DZ for all nodes
    (gcm_dense_ones_window_bwd), and each weight gradient is one reduction over the window's buffers."""

    @staticmethod
    def forward(ctx, anchor, plan, state, chain_id, *params):
        ctx.plan, ctx.state, ctx.chain_id = plan, state, chain_id
        ctx.pkey = plan.gnn.current_key(state.device)
        return anchor.clone()

    @staticmethod
    def backward(ctx, d_token):
        plan, st = ctx.plan, ctx.state
        g, win, dev = plan.gnn, st.win, st.device
        if g.current_key(dev) != ctx.pkey:
            raise RuntimeError("GNN parameters were modified in place between forward and backward")
        if win.chain_id != ctx.chain_id:
            raise RuntimeError("backward through a GCM window after a newer window was recorded on the same state")
        Kc = win.kmax + 1
        win.kmax = -1
        grads = {
            "w_rel1": torch.zeros(g.H1, g.F, device=dev), "w_root1": torch.zeros(g.H1, g.F, device=dev),
            "b1": torch.zeros(g.H1, device=dev),
            "w_rel2": torch.zeros(g.H2, g.H1, device=dev), "w_root2": torch.zeros(g.H2, g.H1, device=dev),
            "b2": torch.zeros(g.H2, device=dev),
        }
        if Kc > 0:
            if not win.need_dx:
                _bwd_small(plan, st, 0, Kc, False)
            steps_total = st.steps - win.chain_start
            if steps_total > st.C - st.N + 1:
                raise RuntimeError(
                    f"BPTT window too long for the node log: {steps_total} steps since the chain started but the log "
                    f"keeps {st.C - st.N} spare rows; raise DenseGCM.bptt_capacity")
            if st.DZ is None:
                st.DZ = torch.empty(st.B, st.C, g.H1, device=dev, dtype=torch.float32)
            _cabi.check(_cabi.lib().gcm_dense_ones_window_bwd(
                st.c_ref(), g.H1, _cabi.ACT[g.act1], int(st.rc_bf16), st.rcache.data_ptr(), steps_total, Kc,
                win.E.data_ptr(), win.dG.data_ptr(), win.dzo.data_ptr(), st.B * g.H1, st.DZ.data_ptr(),
                _cabi.stream_ptr(dev)), "gcm_dense_ones_window_bwd")
            # weight gradients: reductions over (graph, node) and (step, graph) rows; on the bf16 path the products
            # run on the tensor cores (independent rounding per row, fp32 accumulation)
            tc_ok = g.F % 16 == 0 and g.H1 % 16 == 0 and g.H2 % 16 == 0
            outer = (_outer_tc if st.rc_bf16 else _outer_tc32) if tc_ok else _outer
            outer(st.DZ.view(st.B * st.C, g.H1), st.nodes.view(st.B * st.C, st.F), grads["w_root1"])
            rows = Kc * st.B
            do = win.do[:Kc].view(rows, g.H2)
            outer(do, win.G[:Kc].view(rows, g.H1), grads["w_rel2"], grads["b2"])
            outer(do, win.ht[:Kc].view(rows, g.H1), grads["w_root2"])
            outer(win.dc[:Kc].view(rows, g.H1), win.S[:Kc].view(rows, g.F), grads["w_rel1"], grads["b1"])
        return (torch.zeros_like(d_token), None, None, None, *_param_grads(g, grads))


class _OnesStepFn(torch.autograd.Function):
    """One step of the ones path.  Saved: the belief ([B, H2]) and the step's index in the window buffers."""

    @staticmethod
    def forward(ctx, x, token, plan, state, k):
        belief = _forward_kernels(plan, state, x.detach(), k)
        ctx.plan, ctx.state, ctx.k = plan, state, k
        ctx.chain_id = state.win.chain_id
        ctx.save_for_backward(belief)
        return belief, torch.zeros(1, device=state.device)

    @staticmethod
    def backward(ctx, d_belief, d_token):
        plan, st, k = ctx.plan, ctx.state, ctx.k
        g, win, dev = plan.gnn, st.win, st.device
        if win.chain_id != ctx.chain_id:
            raise RuntimeError("backward through a GCM window after a newer window was recorded on the same state")
        (belief,) = ctx.saved_tensors
        lib = _cabi.lib()
        stream = _cabi.stream_ptr(dev)
        win.ensure_bwd()
        db_ = d_belief.contiguous().float()
        _cabi.check(lib.gcm_act_backward(db_.data_ptr(), belief.data_ptr(), _cabi.ACT[g.act2], st.B * g.H2,
                                         win.do[k].data_ptr(), stream), "gcm_act_backward")
        if win.need_dx:
            # dL/dx_k is due now: finish this step's [B, H] work (otherwise the root does it for all steps at once)
            _bwd_small(plan, st, k, k + 1, k < win.kmax)
        tw = g.transposed(dev)
        win.kmax = max(win.kmax, k)
        d_x = None
        if ctx.needs_input_grad[0]:
            Kc = win.kmax + 1
            steps_total = st.steps - win.chain_start
            if steps_total > st.C - st.N + 1:
                raise RuntimeError(
                    f"BPTT window too long for the node log: {steps_total} steps since the chain started but the log "
                    f"keeps {st.C - st.N} spare rows; raise DenseGCM.bptt_capacity")
            dz = torch.empty(st.B, g.H1, device=dev, dtype=torch.float32)
            _cabi.check(lib.gcm_dense_ones_node_bwd(
                st.c_ref(), g.H1, _cabi.ACT[g.act1], int(st.rc_bf16), st.rcache.data_ptr(), steps_total, Kc, k,
                win.E.data_ptr(), win.dG.data_ptr(), win.dzo.data_ptr(), st.B * g.H1, dz.data_ptr(), stream),
                "gcm_dense_ones_node_bwd")
            # dL/dS of a step reaches every node of its window: steps k .. k+N-1 saw node k
            ds = win.dcs[k]
            if k + st.N < Kc:
                ds = ds - win.dcs[k + st.N]
            d_x = _lin2(dz, tw["w_root1_t"], ds, tw["w_rel1_t"])
        return d_x, torch.zeros(1, device=dev), None, None, None


class _OnesSeqFn(torch.autograd.Function):
    """T steps of the ones path taken at once: one node standing for the steps k0 .. k0+T-1 of the window."""

    @staticmethod
    def forward(ctx, x_seq, token, plan, state, k0):
        beliefs = _sequence_kernels(plan, state, x_seq.detach(), k0)
        ctx.plan, ctx.state, ctx.k0, ctx.T = plan, state, k0, x_seq.shape[1]
        ctx.chain_id = state.win.chain_id
        ctx.save_for_backward(beliefs)
        return beliefs, torch.zeros(1, device=state.device)

    @staticmethod
    def backward(ctx, d_beliefs, d_token):
        plan, st, k0, T = ctx.plan, ctx.state, ctx.k0, ctx.T
        g, win, dev = plan.gnn, st.win, st.device
        if win.chain_id != ctx.chain_id:
            raise RuntimeError("backward through a GCM window after a newer window was recorded on the same state")
        (beliefs,) = ctx.saved_tensors
        lib = _cabi.lib()
        stream = _cabi.stream_ptr(dev)
        win.ensure_bwd()
        if (d_beliefs.dtype == torch.float32 and not d_beliefs.is_contiguous() and d_beliefs.stride(2) == 1
                and g.H2 % 4 == 0 and d_beliefs.stride(0) % 4 == 0 and d_beliefs.stride(1) % 4 == 0
                and d_beliefs.data_ptr() % 16 == 0):
            # a [B, T, H] gradient seen time-major: read through its strides, no transposing copy
            _cabi.check(lib.gcm_act_backward_strided(d_beliefs.data_ptr(), d_beliefs.stride(0), d_beliefs.stride(1),
                                                     beliefs.data_ptr(), _cabi.ACT[g.act2], T, st.B, g.H2,
                                                     win.do[k0:k0 + T].data_ptr(), stream), "gcm_act_backward_strided")
        else:
            db_ = d_beliefs.contiguous().float()
            _cabi.check(lib.gcm_act_backward(db_.data_ptr(), beliefs.data_ptr(), _cabi.ACT[g.act2], T * st.B * g.H2,
                                             win.do[k0:k0 + T].data_ptr(), stream), "gcm_act_backward")
        d_x = None
        if win.need_dx:
            # observations require grad: finish the steps newest-first, like T single-step nodes would
            tw = g.transposed(dev)
            steps_total = st.steps - win.chain_start
            if steps_total > st.C - st.N + 1:
                raise RuntimeError(
                    f"BPTT window too long for the node log: {steps_total} steps since the chain started but the log "
                    f"keeps {st.C - st.N} spare rows; raise DenseGCM.bptt_capacity")
            d_x = torch.empty(st.B, T, g.F, device=dev, dtype=torch.float32) if ctx.needs_input_grad[0] else None
            dz = torch.empty(st.B, g.H1, device=dev, dtype=torch.float32)
            for k in range(k0 + T - 1, k0 - 1, -1):
                _bwd_small(plan, st, k, k + 1, k < win.kmax)
                win.kmax = max(win.kmax, k)
                if d_x is None:
                    continue
                Kc = win.kmax + 1
                _cabi.check(lib.gcm_dense_ones_node_bwd(
                    st.c_ref(), g.H1, _cabi.ACT[g.act1], int(st.rc_bf16), st.rcache.data_ptr(), steps_total, Kc, k,
                    win.E.data_ptr(), win.dG.data_ptr(), win.dzo.data_ptr(), st.B * g.H1, dz.data_ptr(), stream),
                    "gcm_dense_ones_node_bwd")
                ds = win.dcs[k]
                if k + st.N < Kc:
                    ds = ds - win.dcs[k + st.N]
                _lin2(dz, tw["w_root1_t"], ds, tw["w_rel1_t"], out=d_x[:, k - k0])
        win.kmax = max(win.kmax, k0 + T - 1)
        return d_x, torch.zeros(1, device=dev), None, None, None


def _chain(plan, state, token, n_steps):
    """Window bookkeeping shared by the recording entries: (token, index of the next step in the window)."""
    win = getattr(state, "win", None)
    if win is None:
        win = state.win = _Window(state, plan.gnn)
    cap = state.C - state.N + 1
    if token is None:
        win.chain_id += 1
        win.chain_start = state.steps
        win.kmax = -1
        win.need_dx = False
        anchor = torch.zeros(1, device=state.device, requires_grad=True)
        token = _OnesRootFn.apply(anchor, plan, state, win.chain_id, *plan.gnn.params())
    k = state.steps - win.chain_start
    if k + n_steps > cap:
        if n_steps > cap:
            raise RuntimeError(
                f"a sequence of {n_steps} recorded steps does not fit the node log ({cap} steps); raise "
                "DenseGCM.bptt_capacity")
        # the log cannot keep more recorded steps: the forward goes on on a fresh chain (see gcm.fused.warn_truncated);
        # backward() through the steps left behind raises
        from gcm import fused
        fused.warn_truncated(cap)
        return _chain(plan, state, None, n_steps)
    win.ensure_fwd(k + n_steps - 1, cap)
    return token, k


def sequence_grad(plan, state, x_seq, token, bf16: bool = False):
    """Recording sequence entry.  Returns (beliefs [T, B, H2], token)."""
    prepare(plan, state, bf16)
    token, k0 = _chain(plan, state, token, x_seq.shape[1])
    state.win.need_dx = state.win.need_dx or x_seq.requires_grad
    beliefs, token = _OnesSeqFn.apply(x_seq, token, plan, state, k0)
    return beliefs, token


def step_grad(plan, state, x, token, bf16: bool = False):
    """Recording step.  Returns (belief, token)."""
    prepare(plan, state, bf16)
    token, k = _chain(plan, state, token, 1)
    state.win.need_dx = state.win.need_dx or x.requires_grad
    belief, token = _OnesStepFn.apply(x, token, plan, state, k)
    return belief, token
