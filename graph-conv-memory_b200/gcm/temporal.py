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
    if plan.gnn.current_key(dev) != ro.key or state.hc_key != ro.key:
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


def sequence_nograd(plan, state, x_seq: torch.Tensor, beliefs: torch.Tensor) -> int:
    """T in-place steps from x_seq [B, T, F] (inner stride 1) into beliefs [B, T, H2] (inner stride 1): the step loop of
    ray_gcm.py:200-202 as ONE C call.  Returns the number of kernels launched.  Caller checked sequence_supported."""
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
    rc = ro.seq_fn(ro.ref, x_seq.data_ptr(), x_seq.stride(0), x_seq.stride(1), beliefs.data_ptr(), beliefs.stride(0),
                   beliefs.stride(1), T, _cabi.stream_ptr(dev))
    if rc:
        _failed(state)
        _cabi.check(rc, "gcm_dense_rollout_fwd")
    _sync_out(ro, state, T)
    return int(c.launches)
