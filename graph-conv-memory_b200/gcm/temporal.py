"""Temporal-chain rollouts of DenseGCM without a Python round trip per kernel argument (and, for sequences, per step).

A forward-only TemporalBackedge chain (edge_selectors/temporal.py:72-88) on a state it built itself runs on the
cached-row kernel (csrc/gcm_dense_fwd_hc.cu).  What the host has to keep between two steps of such a rollout is small:
the uniform node count, how many cached layer-1 rows are valid under the current weights, the weight packs.  This module
keeps it in a `gcm_rollout` descriptor (include/gcm_b200.h) next to the state, so that

* `fast_step`       one DenseGCM.forward step = a 4-argument C call (`gcm_dense_rollout_step`);
* `sequence_nograd` the T steps of `for t in range(T): out, hidden = self.gcm(flat[:, t, :], hidden)` (RayDenseGCM.forward,
                    reference ray_gcm.py:200-202) = ONE C call that enqueues the T step kernels back to back
                    (`gcm_dense_rollout_fwd`), reading the caller's [B, T, F] tensor and writing [B, T, H] in place
                    (strided rows, no per-step copies).

Results are those of the step loop, launch for launch (tests/test_dense_gpu.py::test_temporal_sequence_*).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from gcm import _cabi


class Rollout:
    """Host handle of one gcm_rollout descriptor: (plan, state, weights key) -> C struct."""

    __slots__ = ("c", "ref", "key", "plan", "keep", "step_fn", "seq_fn")

    def __init__(self, plan, state):
        dev = state.device
        gnn_c = plan.gnn.packed(dev)
        self.plan = plan
        self.key = plan.gnn._key
        self.keep = plan.gnn._packed            # the weight packs the descriptor points at
        c = _cabi.RolloutC()
        c.st = state._c
        c.gnn = gnn_c
        sels, n = plan.selectors_c(state.F, None)
        for i in range(n):
            c.sels[i] = sels[i]
        c.n_sels = n
        c.max_hop = plan.max_hop
        c.hcache = state.hcache.data_ptr()
        c.hc_ring = plan.hc_ring
        c.status = state.status.data_ptr()
        c.scratch_obs = None
        c.scratch_belief = None
        self.c = c
        self.ref = C.byref(c)
        lib = _cabi.lib()
        self.step_fn = lib.gcm_dense_rollout_step
        self.seq_fn = lib.gcm_dense_rollout_fwd


def ready(plan, state) -> Optional[Rollout]:
    """The state's descriptor under the CURRENT weights, (re)built when needed; None when the state is not a pure
    temporal state of this plan on the cached-row kernel."""
    if (not plan.hc_ring or state.hcache is None or state.pure_key != plan.temporal_key or state.masks_stale
            or plan.temporal_key is None):
        return None
    plan.gnn.packed(state.device)
    ro = state.rollout
    if ro is None or ro.plan is not plan or ro.key != plan.gnn._key or ro.c.hcache != state.hcache.data_ptr():
        if state.hc_key != plan.gnn._key:
            # new weights: no cached row is valid any more (the C loop then refills the cache on the tc kernel)
            state.hc_key, state.hc_fresh = plan.gnn._key, 0
        ro = state.rollout = Rollout(plan, state)
    return ro


def _sync_in(ro: Rollout, state) -> None:
    c = ro.c
    hc = state.host_count
    c.uniform_count = -1 if hc is None else hc
    c.hc_fresh = state.hc_fresh
    c.weights_stable = 1 if state.hc_fresh >= 1 else 0


def _sync_out(ro: Rollout, state, steps: int) -> None:
    c = ro.c
    state.hc_fresh = c.hc_fresh
    state.fast_ok = c.hc_fresh > 0
    if state.host_count is not None:
        state.host_count = c.uniform_count
    state.max_count += steps
    state.version += 1
    state.steps += steps


def _failed(state) -> None:
    """A C call failed part-way: the host mirrors can no longer be trusted."""
    state.host_count, state.hc_fresh, state.fast_ok, state.rollout = None, 0, False, None
    state.version += 1


def fast_step(plan, state, x: torch.Tensor):
    """The steady-state rollout step (no autograd, cached rows valid): input checks, the weights key, ONE C call.
    Returns None whenever anything is unusual; DenseGCM.forward then takes the general route, which re-derives
    everything (and is the only place that raises)."""
    ro = state.rollout
    if (ro is None or ro.plan is not plan or x.dtype is not torch.float32 or not x.is_cuda or x.dim() != 2
            or x.shape[0] != state.B or x.shape[1] != state.F or not x.is_contiguous()
            or state.pure_key != plan.temporal_key or state.masks_stale or state.hc_fresh < plan.max_hop
            or torch.cuda.is_current_stream_capturing()):
        return None
    dev = state.device
    if not plan.gnn.key_is(ro.key, dev) or (state.hc_key is not ro.key and state.hc_key != ro.key):
        return None                              # weights changed: the general route handles it
    hc = state.host_count
    if hc is not None and hc < plan.max_hop:
        return None
    c = ro.c
    c.uniform_count = -1 if hc is None else hc
    c.hc_fresh = state.hc_fresh
    c.weights_stable = 1
    belief = torch.empty(state.B, plan.gnn.H2, device=dev, dtype=torch.float32)
    rc = ro.step_fn(ro.ref, x.data_ptr(), belief.data_ptr(), _cabi.stream_ptr(dev))
    if rc:
        _failed(state)
        _cabi.check(rc, "gcm_dense_rollout_step")
    _sync_out(ro, state, 1)
    return belief


def sequence_supported(plan, state, x_seq: torch.Tensor) -> bool:
    return (ready(plan, state) is not None and x_seq.dtype is torch.float32 and x_seq.is_cuda and x_seq.dim() == 3
            and x_seq.shape[0] == state.B and x_seq.shape[2] == state.F and x_seq.stride(2) == 1
            and x_seq.stride(0) % 4 == 0 and x_seq.stride(1) % 4 == 0 and x_seq.data_ptr() % 16 == 0
            and not torch.cuda.is_current_stream_capturing())


def sequence_nograd(plan, state, x_seq: torch.Tensor, beliefs: torch.Tensor, rec=None) -> int:
    """T in-place steps from x_seq [B, T, F] (inner stride 1) into beliefs [B, T, H2] (inner stride 1): the step loop of
    ray_gcm.py:200-202 as ONE C call.  Returns the number of kernels launched.  Caller checked sequence_supported.
    rec = [tiled operand buffer, row of the first step, None]: the cached-row kernel also writes the layer-1 operand rows
    of its steps there (gcm_rollout.xrec); rec[2] becomes the index of the first step it wrote."""
    ro = ready(plan, state)
    dev = state.device
    T = x_seq.shape[1]
    c = ro.c
    strided = x_seq.stride(0) != state.F or beliefs.stride(0) != plan.gnn.H2
    if strided and not c.scratch_obs:
        scr = state.__dict__.get("_seq_scratch")
        if scr is None:
            scr = state.__dict__["_seq_scratch"] = (
                torch.empty(state.B, state.F, device=dev, dtype=torch.float32),
                torch.empty(state.B, plan.gnn.H2, device=dev, dtype=torch.float32))
        c.scratch_obs, c.scratch_belief = scr[0].data_ptr(), scr[1].data_ptr()
    _sync_in(ro, state)
    if rec is not None:
        c.xrec, c.xrec_row0 = rec[0].data_ptr(), rec[1]
    rc = ro.seq_fn(ro.ref, x_seq.data_ptr(), x_seq.stride(0), x_seq.stride(1), beliefs.data_ptr(), beliefs.stride(0),
                   beliefs.stride(1), T, _cabi.stream_ptr(dev))
    if rec is not None:
        c.xrec, rec[2] = None, (T if rc else int(c.xrec_from))
    if rc:
        _failed(state)
        _cabi.check(rc, "gcm_dense_rollout_fwd")
    _sync_out(ro, state, T)
    return int(c.launches)


# ------------------------------------------------------------------------------------------------
# BPTT: window-level backward (csrc/gcm_temporal_bwd.cu + the 3xTF32 products of csrc/gcm_tc_gemm.cu)
# ------------------------------------------------------------------------------------------------
# Autograd through gcm.py:262-321 over the steps of a BPTT window (the reference's training loop:
# tests/test_gcm.py:412-439).  GCM has no recurrence through the belief, so when backward() runs every step's
# dL/dbelief is known; each recorded step (or sequence call) only delivers dz2 = dL/dbelief * act2'(belief) into the
# window buffer, and the chain's root node -- which autograd runs last, every step depends on it through the token
# chain -- does the whole window's work as a few row-parallel products over [rows, graphs] (see the .cu header for
# the algebra).  Nothing but the beliefs is saved: layer 1 is recomputed from the node log, which keeps every row
# the window needs because the log has spare capacity while gradients are being recorded (DenseState.C > N).
def grad_supported(plan, state) -> bool:
    """Forward-only hops on the cached-row shape, state built from empty by this chain (so an edge p-s -> p exists iff
    p - s >= 0 and every graph has the same count, mirrored on the host)."""
    return bool(plan.hc_ring and plan.temporal_key is not None and state.pure_key in ((), plan.temporal_key)
                and state.host_count is not None and state.F % 4 == 0 and state.N - 1 >= 2 * plan.max_hop
                and not state.masks_stale)


class _TWindow:
    """Bookkeeping of the BPTT window being recorded on a state (one chain of steps at a time)."""

    def __init__(self):
        self.chain_id = 0
        self.chain_start = 0      # state.steps when the chain started
        self.P0 = 0               # absolute position of the node written by the chain's first step
        self.kmax = -1            # newest-first backward: steps kmax .. k have delivered dz2 so far
        self.dz2 = None           # [K, B, H2] float32, time-major; zero rows = steps whose belief got no gradient
        self.cache = None         # (A1, h, dz1) over the whole window, left by a sequence node for the root
        self.lazy = None          # (dL/dbelief, beliefs), both [K, B, H2] time-major: dz2 of the WHOLE window not yet formed
        self.xrec = None          # [tiled X buffer, row of step 0, first step the forward kernel wrote] of the WHOLE window


def _hops(plan, dev):
    """the chain's hop list as a HOST int32 array (the C entry points read it on the host)"""
    t = plan.__dict__.get("_hops_host")
    if t is None:
        hops = sorted({h for s in plan.sels for h in s.hops})
        t = plan.__dict__["_hops_host"] = (torch.tensor(hops, dtype=torch.int32, device="cpu"), len(hops))
    return t


def _wcat(plan, dev):
    """[W_rel | W_root] packs for the window products, rebuilt when a parameter changed."""
    g = plan.gnn
    g.packed(dev)
    if g.__dict__.get("_wcat_key") != g._key:
        w = g._packed[1]
        g._wcat = {
            "w1": torch.cat([w["w_rel1"], w["w_root1"]], dim=1).contiguous(),               # [H1, 2F]   h  = A1 w1^T
            "w2t": torch.cat([w["w_rel2"].t(), w["w_root2"].t()], dim=1).contiguous(),      # [H1, 2H2]  dh = D2 w2t^T
            "w1t": torch.cat([w["w_rel1"].t(), w["w_root1"].t()], dim=1).contiguous(),      # [F, 2H1]   dx = D1 w1t^T
            "b1": w["b1"],
        }
        g._wcat_key = g._key
    return g._wcat


def _lin(a, w, bias=None, act=0, out=None):
    from gcm import ones
    rows, k = a.shape
    if k % 16 == 0 and w.shape[0] % 16 == 0:
        return ones._lin_tc32(a, w, bias=bias, act=act, out=out)
    return ones._lin2(a, w, bias=bias, act=act, out=out)


def _outer(a, x, dw, db):
    """dw += a^T x, db += column sums of a, for NARROW operands (a [rows, 32], x [rows, 16..64]).
    gcm_outer_reduce_tc32 maps channels to threads (128 per group), so a 32 x 64 product would leave most of its threads
    idle; viewing s consecutive rows as one (x as [rows / s, s * Hx] in the kernel's wide A role, a as [rows / s, s * Ha])
    makes it a 128 x (s * Ha) product whose s diagonal [Hx, Ha] blocks sum to (a^T x)^T -- same bytes read, s^2 / s of
    the MMA work wasted on the off-diagonal blocks, all threads loading."""
    from gcm import ones
    rows, ha = a.shape
    hx = x.shape[1]
    s = min(128 // hx, 128 // ha)
    while s > 1 and rows % s:
        s //= 2
    if s > 1 and hx % 16 == 0 and (s * ha) % 16 == 0 and a.is_contiguous() and x.is_contiguous():
        tmp = torch.zeros(s * hx, s * ha, device=a.device, dtype=torch.float32)
        ones._outer_tc32(x.view(rows // s, s * hx), a.view(rows // s, s * ha), tmp, None)
        acc = tmp[:hx, :ha]
        for i in range(1, s):
            acc = acc + tmp[i * hx:(i + 1) * hx, i * ha:(i + 1) * ha]
        dw += acc.t()
        db += a.sum(0)
    elif hx % 16 == 0:
        ones._outer_tc32(a, x, dw, db)
    else:
        ones._outer(a, x, dw, db)


def _shift_sum(plan, src, src_pos0, sign, out, out_pos0, n_out=None, act_out=None, act=0):
    """n_out given: `out` is a flat buffer in the tiled layout of the fused window kernel (see include/gcm_b200.h);
    act_out given: the source rows are src * act'(act_out)"""
    hops, nh = _hops(plan, src.device)
    n_src, B, H = src.shape
    if src.dtype != torch.float32:
        src = src.float()
    if not src.is_contiguous():
        # dL/dbelief as autograd delivered it -- a [B, T, H] tensor seen time-major, or the expanded scalar of a sum loss --
        # is read through its strides where the kernel can (no transposing copy, no materialised broadcast)
        rc = _cabi.ERR_UNSUPPORTED
        if src.dtype == torch.float32 and n_out is not None:
            rc = _cabi.lib().gcm_temporal_shift_sum_strided(
                src.data_ptr(), src.stride(0), src.stride(1), src.stride(2), src_pos0, n_src, 0, hops.data_ptr(), nh, sign,
                out.data_ptr(), out_pos0, n_out, B, H, 1, None if act_out is None else act_out.data_ptr(), act,
                _cabi.stream_ptr(src.device))
        if rc != _cabi.ERR_UNSUPPORTED:
            _cabi.check(rc, "gcm_temporal_shift_sum_strided")
            return
        src = src.contiguous().float()
    _cabi.check(_cabi.lib().gcm_temporal_shift_sum(src.data_ptr(), src_pos0, n_src, 0, hops.data_ptr(), nh, sign,
                                                   out.data_ptr(), out_pos0, out.shape[0] if n_out is None else n_out, B, H,
                                                   0 if n_out is None else 1,
                                                   None if act_out is None else act_out.data_ptr(), act,
                                                   _cabi.stream_ptr(src.device)),
                "gcm_temporal_shift_sum")


def _rows(plan, st, win, p_lo, p_hi, Kc):
    """Layer-1 operand rows A1 = [sum_s x_{p-s} | x_p], h_p and dz1_p = dL/d(pre-activation of layer 1) for the node
    positions p_lo .. p_hi-1, given the dz2 delivered so far (chain steps 0 .. Kc-1 = positions P0 .. P0+Kc-1)."""
    g, dev = plan.gnn, st.device
    n, B = p_hi - p_lo, st.B
    hops, nh = _hops(plan, dev)
    wc = _wcat(plan, dev)
    lib = _cabi.lib()
    A1 = torch.empty(n, B, 2 * st.F, device=dev, dtype=torch.float32)
    _cabi.check(lib.gcm_temporal_gather(st.c_ref(), hops.data_ptr(), nh, p_lo, n, A1.data_ptr(), 0, _cabi.stream_ptr(dev)),
                "gcm_temporal_gather")
    h = _lin(A1.view(n * B, 2 * st.F), wc["w1"], bias=wc["b1"], act=_cabi.ACT[g.act1]).view(n, B, g.H1)
    D2 = torch.empty(n, B, 2 * g.H2, device=dev, dtype=torch.float32)
    _shift_sum(plan, win.dz2[:Kc], win.P0, +1, D2, p_lo)
    dz1 = _lin(D2.view(n * B, 2 * g.H2), wc["w2t"]).view(n, B, g.H1)
    del D2
    _cabi.check(lib.gcm_act_backward(dz1.data_ptr(), h.data_ptr(), _cabi.ACT[g.act1], n * B * g.H1, dz1.data_ptr(),
                                     _cabi.stream_ptr(dev)), "gcm_act_backward")
    return A1, h, dz1


_WB_WS = {}


def _fused_shape(plan, st) -> bool:
    g = plan.gnn
    return not os.environ.get("GCM_B200_NO_WINDOW_BWD_TC") and st.F == 32 and g.H1 == 32 and g.H2 == 32


def _window_fused(plan, st, win, Kc, want_dz1=False):
    """Both layers' row products and both weight-gradient reductions of the window in ONE kernel
    (csrc/gcm_temporal_bwd_tc.cu) over the operand rows [sum x | x] and [sum dz2 | dz2] of the positions
    P0 - max_hop .. P0 + Kc - 1.  Returns (grads dict, (first position, dz1 rows) or None), or None when the shape is not
    the kernel's (F = H1 = H2 = 32) -- the caller then runs the separate products."""
    g, dev = plan.gnn, st.device
    if not _fused_shape(plan, st):
        return None
    lib = _cabi.lib()
    hops, nh = _hops(plan, dev)
    wc = _wcat(plan, dev)
    p_lo, p_hi = win.P0 - plan.max_hop, win.P0 + Kc
    n, B = p_hi - p_lo, st.B
    rows = n * B
    padded = (rows + 127) // 128 * 128          # the kernel reads whole tiles of 128 rows; rows past the end are zero
    alloc = torch.empty if padded == rows else torch.zeros
    n_gather = n
    if win.xrec is not None and win.xrec[0].numel() == padded * 64 and win.xrec[1] == plan.max_hop * B:
        A1, n_gather = win.xrec[0], plan.max_hop + win.xrec[2]      # the rest was written by the forward kernel
    else:
        A1 = alloc(padded * 64, device=dev, dtype=torch.float32)
    if n_gather > 0:
        _cabi.check(lib.gcm_temporal_gather(st.c_ref(), hops.data_ptr(), nh, p_lo, n_gather, A1.data_ptr(), 1,
                                            _cabi.stream_ptr(dev)), "gcm_temporal_gather")
    D2 = alloc(padded * 64, device=dev, dtype=torch.float32)
    if win.lazy is not None:
        _shift_sum(plan, win.lazy[0], win.P0, +1, D2, p_lo, n_out=n, act_out=win.lazy[1], act=_cabi.ACT[g.act2])
    else:
        _shift_sum(plan, win.dz2[:Kc], win.P0, +1, D2, p_lo, n_out=n)
    ws = _WB_WS.get(dev)
    if ws is None:
        ws = _WB_WS[dev] = torch.empty(int(lib.gcm_temporal_window_bwd_workspace()), device=dev, dtype=torch.float32)
    acc = torch.zeros(2 * 64 * 32 + 64, device=dev, dtype=torch.float32)
    g1, g2 = acc[:2048].view(64, 32), acc[2048:4096].view(64, 32)
    db1, db2 = acc[4096:4128], acc[4128:4160]
    dz1 = torch.empty(n, B, 32, device=dev, dtype=torch.float32) if want_dz1 else None
    _cabi.check(lib.gcm_temporal_window_bwd(A1.data_ptr(), D2.data_ptr(), n * B, wc["w1"].data_ptr(), wc["b1"].data_ptr(),
                                            wc["w2t"].data_ptr(), _cabi.ACT[g.act1], ws.data_ptr(), g1.data_ptr(),
                                            g2.data_ptr(), db1.data_ptr(), db2.data_ptr(),
                                            None if dz1 is None else dz1.data_ptr(), _cabi.stream_ptr(dev)),
                "gcm_temporal_window_bwd")
    grads = {"w_rel1": g1[:32].t().contiguous(), "w_root1": g1[32:].t().contiguous(), "b1": db1,
             "w_rel2": g2[:32].contiguous(), "w_root2": g2[32:].contiguous(), "b2": db2}
    return grads, (None if dz1 is None else (p_lo, dz1))


def _check_log(st, win, plan, margin=2):
    steps_total = st.steps - win.chain_start
    if steps_total + margin * plan.max_hop > st.C:
        raise RuntimeError(
            f"BPTT window too long for the node log: {steps_total} steps since the chain started but the log keeps "
            f"{st.C - st.N} spare rows; raise DenseGCM.bptt_capacity")


def _dx(plan, st, win, k_lo, k_hi, Kc, rows=None):
    """dL/dx of chain steps k_lo .. k_hi-1 ([n, B, F], time-major); valid once steps k_lo .. Kc-1 delivered dz2."""
    g = plan.gnn
    mh = plan.max_hop
    q_lo, q_hi = win.P0 + k_lo, win.P0 + k_hi
    if rows is None:
        z_lo, z_hi = q_lo, min(q_hi + mh, win.P0 + Kc)
        _, _, dz1 = _rows(plan, st, win, z_lo, z_hi, Kc)
    else:
        z_lo, dz1 = rows
    D1 = torch.empty(k_hi - k_lo, st.B, 2 * g.H1, device=st.device, dtype=torch.float32)
    _shift_sum(plan, dz1, z_lo, +1, D1, q_lo)
    return _lin(D1.view(-1, 2 * g.H1), _wcat(plan, st.device)["w1t"]).view(k_hi - k_lo, st.B, st.F)


def _deliver(plan, st, win, k0, n, d_beliefs_tm, beliefs_tm):
    """dz2 of chain steps k0 .. k0+n-1 from dL/dbelief and the beliefs (both [n, B, H2], time-major, contiguous)."""
    g = plan.gnn
    K = st.steps - win.chain_start
    if win.dz2 is None or win.dz2.shape[0] < K:
        win.dz2 = torch.zeros(K, st.B, g.H2, device=st.device, dtype=torch.float32)
    _cabi.check(_cabi.lib().gcm_act_backward(d_beliefs_tm.data_ptr(), beliefs_tm.data_ptr(), _cabi.ACT[g.act2],
                                             n * st.B * g.H2, win.dz2[k0:k0 + n].data_ptr(),
                                             _cabi.stream_ptr(st.device)), "gcm_act_backward")
    win.kmax = max(win.kmax, k0 + n - 1)


class _TRootFn(torch.autograd.Function):
    """Start of a recorded chain of temporal steps.  Runs LAST in backward: every weight gradient of the window."""

    @staticmethod
    def forward(ctx, anchor, plan, state, chain_id, *params):
        ctx.plan, ctx.state, ctx.chain_id = plan, state, chain_id
        ctx.pkey = plan.gnn.current_key(state.device)
        return anchor.clone()

    @staticmethod
    def backward(ctx, d_token):
        from gcm import ones
        plan, st = ctx.plan, ctx.state
        g, win, dev = plan.gnn, st.twin, st.device
        if g.current_key(dev) != ctx.pkey:
            raise RuntimeError("GNN parameters were modified in place between forward and backward")
        if win.chain_id != ctx.chain_id:
            raise RuntimeError("backward through a GCM window after a newer window was recorded on the same state")
        Kc = win.kmax + 1
        F, H1, H2, B, mh = st.F, g.H1, g.H2, st.B, plan.max_hop
        dW1 = torch.zeros(H1, 2 * F, device=dev)
        dW2 = torch.zeros(H2, 2 * H1, device=dev)
        db1, db2 = torch.zeros(H1, device=dev), torch.zeros(H2, device=dev)
        fused = None
        if Kc > 0:
            _check_log(st, win, plan)
        if Kc > 0 and win.cache is None:
            fused = _window_fused(plan, st, win, Kc)
        elif Kc > 0 and isinstance(win.cache[0], str):
            fused = (win.cache[1], None)                 # a sequence node already ran the fused kernel
        if fused is not None:
            win.kmax, win.dz2, win.cache, win.lazy, win.xrec = -1, None, None, None, None
            return (torch.zeros_like(d_token), None, None, None, *ones._param_grads(g, fused[0]))
        if win.lazy is not None:
            _deliver(plan, st, win, 0, Kc, win.lazy[0].contiguous().float(), win.lazy[1])
            win.lazy = None
        if Kc > 0:
            p_lo, p_hi = win.P0 - mh, win.P0 + Kc
            if win.cache is not None:
                A1, h, dz1 = win.cache
            else:
                A1, h, dz1 = _rows(plan, st, win, p_lo, p_hi, Kc)
            n = p_hi - p_lo
            _outer(dz1.view(n * B, H1), A1.view(n * B, 2 * F), dW1, db1)
            del A1, dz1
            A2 = torch.empty(Kc, B, 2 * H1, device=dev, dtype=torch.float32)
            _shift_sum(plan, h, p_lo, -1, A2, win.P0)
            _outer(win.dz2[:Kc].view(Kc * B, H2), A2.view(Kc * B, 2 * H1), dW2, db2)
        win.kmax, win.dz2, win.cache, win.lazy, win.xrec = -1, None, None, None, None
        grads = {"w_rel1": dW1[:, :F].contiguous(), "w_root1": dW1[:, F:].contiguous(), "b1": db1,
                 "w_rel2": dW2[:, :H1].contiguous(), "w_root2": dW2[:, H1:].contiguous(), "b2": db2}
        return (torch.zeros_like(d_token), None, None, None, *ones._param_grads(g, grads))


class _TStepFn(torch.autograd.Function):
    """One recorded step.  Saved: the belief [B, H2]."""

    @staticmethod
    def forward(ctx, x, token, plan, state, k):
        from gcm import fused
        belief = torch.empty(state.B, plan.gnn.H2, device=state.device, dtype=torch.float32)
        fused._launch_fwd(plan, state, x.detach(), belief)
        ctx.plan, ctx.state, ctx.k = plan, state, k
        ctx.chain_id = state.twin.chain_id
        ctx.save_for_backward(belief)
        return belief, torch.zeros(1, device=state.device)

    @staticmethod
    def backward(ctx, d_belief, d_token):
        plan, st, k = ctx.plan, ctx.state, ctx.k
        win = st.twin
        if win.chain_id != ctx.chain_id:
            raise RuntimeError("backward through a GCM window after a newer window was recorded on the same state")
        (belief,) = ctx.saved_tensors
        _deliver(plan, st, win, k, 1, d_belief.contiguous().float(), belief)
        d_x = None
        if ctx.needs_input_grad[0]:
            _check_log(st, win, plan)
            d_x = _dx(plan, st, win, k, k + 1, win.kmax + 1)[0]
        return d_x, torch.zeros(1, device=st.device), None, None, None


class _TSeqFn(torch.autograd.Function):
    """T recorded steps taken at once (DenseGCM.forward_sequence): one node for the steps k0 .. k0+T-1.
    hist (optional, first node of a chain only): the GNN-input rows of the 2 max_hop nodes written BEFORE the chain,
    [B, 2 max_hop, F], as a function of something that requires grad -- a preprocessor's image of the stored raw
    observations (the reference maps every stored row through the preprocessor at every step, gcm.py:290-291, so its
    parameters also collect gradient through rows older than the window).  Forward ignores its values (the node log
    already holds them); backward returns dL/dhist."""

    @staticmethod
    def forward(ctx, x_seq, token, plan, state, k0, hist):
        from gcm import fused
        T = x_seq.shape[1]
        buf = torch.empty(T, state.B, plan.gnn.H2, device=state.device, dtype=torch.float32)   # time-major
        xs = x_seq.detach()
        rec = None
        if sequence_supported(plan, state, xs):
            if k0 == 0 and _fused_shape(plan, state):
                # the forward kernel leaves the layer-1 operand rows of its steps for the fused window backward
                rows = (plan.max_hop + T) * state.B
                padded = (rows + 127) // 128 * 128
                rec = [(torch.empty if padded == rows else torch.zeros)(padded * 64, device=state.device), plan.max_hop * state.B,
                       None]
            sequence_nograd(plan, state, xs, buf.transpose(0, 1), rec)
        else:
            for t in range(T):
                fused._launch_fwd(plan, state, xs[:, t].contiguous(), buf[t])
        ctx.plan, ctx.state, ctx.k0, ctx.T = plan, state, k0, T
        ctx.chain_id = state.twin.chain_id
        ctx.buf = buf
        ctx.xrec = rec
        ctx.has_hist = hist is not None
        return buf.transpose(0, 1), torch.zeros(1, device=state.device)

    @staticmethod
    def backward(ctx, d_beliefs, d_token):
        plan, st, k0, T = ctx.plan, ctx.state, ctx.k0, ctx.T
        win = st.twin
        if win.chain_id != ctx.chain_id:
            raise RuntimeError("backward through a GCM window after a newer window was recorded on the same state")
        d_x = d_hist = None
        want_hist = ctx.has_hist and ctx.needs_input_grad[5]
        if k0 == 0 and T == st.steps - win.chain_start and win.kmax < 0 and _fused_shape(plan, st):
            # this node is the whole window: the fused kernel's shift-sum forms dz2 on the fly, reading dL/dbelief
            # through its strides (time-major VIEW: not copied here)
            win.lazy, win.kmax, win.xrec = (d_beliefs.transpose(0, 1), ctx.buf), T - 1, ctx.xrec
        else:
            _deliver(plan, st, win, k0, T, d_beliefs.transpose(0, 1).contiguous().float(), ctx.buf)
        ctx.buf = ctx.xrec = None
        if ctx.needs_input_grad[0] or want_hist:
            Kc, mh = win.kmax + 1, plan.max_hop
            _check_log(st, win, plan)
            rows = None
            if k0 == 0:
                # the chain's first node runs last: every step has delivered.  One evaluation of the window serves dL/dx
                # here, dL/dhist, and the weight gradients at the root
                fz = _window_fused(plan, st, win, Kc, want_dz1=True)
                if fz is not None:
                    win.cache, rows = ("fused", fz[0]), fz[1]     # the weight gradients are done: the root returns them
                else:
                    if win.lazy is not None:
                        _deliver(plan, st, win, 0, Kc, win.lazy[0].contiguous().float(), win.lazy[1])
                        win.lazy = None
                    full = _rows(plan, st, win, win.P0 - mh, win.P0 + Kc, Kc)
                    win.cache = full
                    rows = (win.P0 - mh, full[2])
            if ctx.needs_input_grad[0]:
                if rows is not None:
                    d_x = _dx(plan, st, win, 0, T, Kc, rows=(win.P0, rows[1][mh:]))
                else:
                    d_x = _dx(plan, st, win, k0, k0 + T, Kc)
                d_x = d_x.transpose(0, 1)
            if want_hist:
                assert k0 == 0
                # rows P0 - 2 max_hop .. P0 - 1: dz1 vanishes below P0 - max_hop, so the cached rows are all that is needed
                d_hist = _dx(plan, st, win, -2 * mh, 0, Kc, rows=rows).transpose(0, 1)
        return d_x, torch.zeros(1, device=st.device), None, None, None, d_hist


def _chain(plan, state, token, n_steps):
    win = state.twin
    if win is None:
        win = state.twin = _TWindow()
    cap = state.C - state.N + 1
    if token is None:
        win.chain_id += 1
        win.chain_start = state.steps
        win.P0 = state.host_count
        win.kmax, win.dz2, win.cache, win.lazy, win.xrec = -1, None, None, None, None
        anchor = torch.zeros(1, device=state.device, requires_grad=True)
        token = _TRootFn.apply(anchor, plan, state, win.chain_id, *plan.gnn.params())
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
    return token, k


def _room(plan, state, token, n_steps, bptt_capacity):
    """A state without enough spare log rows is re-homed (-> a new chain) before the recorded step(s)."""
    from gcm import fused
    if state.C - state.N + 1 < n_steps or state.C - state.N < 1:
        if token is not None:
            fused.warn_truncated(state.C - state.N + 1)
        state = fused.grow_state(state, state.N + max(int(bptt_capacity), n_steps, 1))
        token = None                  # a re-homed log starts a new chain
    return state, token


def step_grad(plan, state, x, token, bptt_capacity):
    """Recording step.  Returns (belief, token, state) -- the state may have been re-homed."""
    state.fast_ok = False
    state, token = _room(plan, state, token, 1, bptt_capacity)
    token, k = _chain(plan, state, token, 1)
    belief, token = _TStepFn.apply(x, token, plan, state, k)
    token._gcm_tw = True
    return belief, token, state


def sequence_grad(plan, state, x_seq, token, bptt_capacity, hist=None):
    """Recording sequence entry.  Returns (beliefs [B, T, H2], token, state).  hist: see _TSeqFn (chain start only)."""
    state.fast_ok = False
    T = x_seq.shape[1]
    state, token = _room(plan, state, token, T, bptt_capacity)
    token, k0 = _chain(plan, state, token, T)
    assert hist is None or k0 == 0
    beliefs, token = _TSeqFn.apply(x_seq, token, plan, state, k0, hist)
    token._gcm_tw = True
    return beliefs, token, state
