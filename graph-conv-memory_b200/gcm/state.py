"""In-place hidden state of the fused DenseGCM path.

The reference's hidden state is the tuple (nodes[B,N,F], adj[B,N,N], weights, num_nodes[B])
(gcm.py:194-211) and every step returns fresh copies of all of it.  Here the graph lives in HBM as
a node log plus bit-packed adjacency (layout: include/gcm_b200.h) that the step kernel updates in
place; `DenseHidden` is what `DenseGCM.forward` hands back as `m_t`.  It unpacks, indexes and
iterates like the reference 4-tuple and materialises the reference-layout tensors on demand
(`gcm_state_materialize`), so `nodes, adj, weights, num_nodes = m_t` and `list(m_t)` keep working
(the latter is what RayDenseGCM does, ray_gcm.py:203-204).

Aliasing rule: the buffers are shared between successive `m_t`s.  An `m_t` that was
neither materialised nor passed to the next `forward` call before a newer step overwrote the
buffers is STALE; touching it raises instead of silently returning newer data.  Materialised
tensors are private copies and never change.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch

from gcm import _cabi


class DenseState:
    """Device buffers of one batch of graphs + host-side bookkeeping."""

    def __init__(self, B: int, N: int, F: int, device, capacity: Optional[int] = None):
        if N > _cabi.GCM_MAX_N:
            raise _cabi.GcmLibraryError(f"graph_size {N} > {_cabi.GCM_MAX_N}")
        self.B, self.N, self.F = B, N, F
        self.C = int(capacity) if capacity is not None else N
        assert self.C >= N
        self.W = (N + 31) // 32
        self.device = torch.device(device)
        self.nodes = torch.zeros(B, self.C, F, device=device, dtype=torch.float32)
        self.masks = torch.zeros(B, self.C, 2, self.W, device=device, dtype=torch.int32)
        self.count = torch.zeros(B, device=device, dtype=torch.int32)
        self.status = torch.zeros(2, device=device, dtype=torch.int32)   # [flags, max count at ingest]
        self.version = 0            # bumped by every in-place step
        self.steps = 0              # steps applied to these buffers (host mirror of count - count0)
        self.pure_key = ()          # selector-chain key if built from empty by one chain, else None
        self.host_count = 0         # host mirror of count[b] when all graphs share it, else None
        self.weights0: Optional[torch.Tensor] = None  # caller-supplied edge weights (passed through)
        self.d_nodes: Optional[torch.Tensor] = None   # dL/dnodes accumulation buffer (training)
        self.grad_floor = 0         # first step index whose window is fully inside the log
        # DenseEdge-only path (gcm.ones): the adjacency is implicit (all ones over the valid block)
        self.dense_ok = True        # adjacency known to be the all-ones valid block (an empty state is)
        self.masks_stale = False    # steps were taken on the ones path: the bit masks must be rebuilt before use
        self.max_count = 0          # host-side upper bound of count[b]
        self.xsum: Optional[torch.Tensor] = None      # [B, F] sum of the window's rows
        self.rcache: Optional[torch.Tensor] = None    # [B, C, H1] per-node cache: exp(2 W_root1 x_i) (tanh) or W_root1 x_i
        self.rc_key = None          # (W_root1 identity/version, act1, element type) the cache was filled under
        self.rc_bf16 = False        # cache element type: float32 or bfloat16
        self.DZ: Optional[torch.Tensor] = None        # [B, C, H1] dL/d(pre-activation) per node of a BPTT window
        # DenseGCM with a row-wise preprocessor (gcm.py:290-291): `nodes` holds the PREPROCESSED rows the kernels read,
        # `raw` [B, C, F_raw] the observations as the caller sees them in m_t; pre_key = preprocessor weights the rows
        # were computed under
        self.raw: Optional[torch.Tensor] = None
        self.pre_key = None
        self.win = None             # gcm.ones._Window: per-step buffers of the BPTT window being recorded
        self.ones_tmp = None        # per-step scratch of non-recording steps
        # single distance selector (gcm.fused.zc_step): per-node pre-activation cache
        self.zc_ok = True           # every step so far went through the zc kernel (an empty state qualifies)
        self.zcache: Optional[torch.Tensor] = None    # [B, C, H1]
        self.zc_key = None
        self.hcache: Optional[torch.Tensor] = None    # layer-1 row cache [B, ring, H1] (gcm.fused._launch_fwd)
        self.hc_key = None          # weights key the cached rows were computed under
        self.hc_fresh = 0           # newest nodes whose cached row is valid under hc_key
        self.fast_ok = False        # last step was a steady-state row-cache step (gcm.fused.fast_temporal_step may run)
        self.twin = None            # gcm.temporal._TWindow: the BPTT window being recorded on a temporal chain
        self.rollout = None         # gcm.temporal.Rollout: the gcm_rollout descriptor of this state (temporal chains)
        self._c = _cabi.DenseStateC(self.nodes.data_ptr(), self.masks.data_ptr(), self.count.data_ptr(),
                                    B, N, self.C, F, self.W)

    # -- C view ---------------------------------------------------------------------------------
    def c_ref(self):
        return C.byref(self._c)

    def sub(self, nb: int) -> "_cabi.DenseStateC":
        """C view of the first `nb` graphs (used by the one-time plan validation)."""
        return _cabi.DenseStateC(self.nodes.data_ptr(), self.masks.data_ptr(), self.count.data_ptr(),
                                 min(nb, self.B), self.N, self.C, self.F, self.W)

    # -- construction ---------------------------------------------------------------------------
    @classmethod
    def ingest(cls, nodes: torch.Tensor, adj: torch.Tensor, weights: torch.Tensor,
               num_nodes: torch.Tensor, capacity: Optional[int] = None) -> Tuple["DenseState", int]:
        """Caller-supplied reference-layout state -> packed state.  Returns (state, flags); flags has
        FLAG_UNCLEAN / FLAG_BADCOUNT set when the adjacency cannot be represented (host sync)."""
        _cabi.require_cuda(nodes, "DenseGCM hidden state")
        B, N, F = nodes.shape
        st = cls(B, N, F, nodes.device, capacity)
        st.pure_key = None
        st.host_count = None
        nodes_c = nodes.detach().to(torch.float32).contiguous()
        adj_c = adj.detach().to(device=nodes.device, dtype=torch.float32).contiguous()
        nn_c = num_nodes.detach().to(device=nodes.device, dtype=torch.long).contiguous()
        _cabi.check(
            _cabi.lib().gcm_state_ingest(st.c_ref(), nodes_c.data_ptr(), adj_c.data_ptr(), nn_c.data_ptr(),
                                         st.status.data_ptr(), _cabi.stream_ptr(nodes.device)),
            "gcm_state_ingest")
        flags, max_count = (int(v) for v in st.status.tolist())
        st.status.zero_()
        st.dense_ok = not (flags & (_cabi.FLAG_NOTDENSE | _cabi.FLAG_UNCLEAN | _cabi.FLAG_BADCOUNT))
        st.max_count = max_count
        st.zc_ok = False
        if weights is not None and weights.numel() != 0:
            st.weights0 = weights
        return st, flags

    def clone(self) -> "DenseState":
        """An independent copy of the graphs (device copies of the node log, the adjacency masks when they are current,
        the counters): a checkpoint several rollouts / BPTT windows can restart from without going back through the
        reference-layout tensors (materialise + ingest moves the [B,N,N] float adjacency; this moves the log only).
        Caches that depend on the layer weights (per-node caches, cached layer-1 rows) are rebuilt on first use."""
        new = DenseState(self.B, self.N, self.F, self.device, self.C)
        new.nodes.copy_(self.nodes)
        new.count.copy_(self.count)
        if not self.masks_stale:
            new.masks.copy_(self.masks)
        new.masks_stale = self.masks_stale
        new.status = self.status
        new.pure_key, new.host_count, new.dense_ok = self.pure_key, self.host_count, self.dense_ok
        new.max_count, new.weights0 = self.max_count, self.weights0
        new.zc_ok = False
        if self.raw is not None:
            new.raw, new.pre_key = self.raw.clone(), self.pre_key
        if self.dense_ok:
            # DenseEdge states: the running sum of the window's rows depends on the log only -- formed once on the source
            # (a pass over the whole node log: 0.37 ms at BASELINE cfg3) and carried by every clone
            if self.xsum is None:
                self.xsum = torch.empty(self.B, self.F, device=self.device, dtype=torch.float32)
                _cabi.check(_cabi.lib().gcm_dense_ones_xsum(self.c_ref(), self.xsum.data_ptr(),
                                                            _cabi.stream_ptr(self.device)), "gcm_dense_ones_xsum")
            new.xsum = self.xsum.clone()
        return new

    # -- materialisation ------------------------------------------------------------------------
    def sync_masks(self) -> None:
        """The ones path does not maintain the adjacency bit masks; write them before anything reads them."""
        if self.masks_stale:
            _cabi.check(_cabi.lib().gcm_dense_fill_masks(self.c_ref(), _cabi.stream_ptr(self.device)),
                        "gcm_dense_fill_masks")
            self.masks_stale = False

    def raw_ref(self):
        """C view of the raw-observation log (same masks and counters)."""
        r = self.raw
        c = self.__dict__.get("_raw_c")
        if c is None or c.nodes != r.data_ptr():
            c = self._raw_c = _cabi.DenseStateC(r.data_ptr(), self.masks.data_ptr(), self.count.data_ptr(), self.B,
                                                self.N, self.C, r.shape[2], self.W)
        return C.byref(c)

    def materialize(self, want_adj: bool = True):
        self.sync_masks()
        if self.raw is not None:
            # the caller's view of the nodes is the raw observations; adjacency and counters come from the state
            lib = _cabi.lib()
            nodes = torch.empty(self.B, self.N, self.raw.shape[2], device=self.device, dtype=torch.float32)
            adj = torch.empty(self.B, self.N, self.N, device=self.device, dtype=torch.float32) if want_adj else None
            num_nodes = torch.empty(self.B, device=self.device, dtype=torch.long)
            _cabi.check(lib.gcm_state_materialize(self.raw_ref(), nodes.data_ptr(), None, num_nodes.data_ptr(),
                                                  _cabi.stream_ptr(self.device)), "gcm_state_materialize")
            if want_adj:
                _cabi.check(lib.gcm_state_materialize(self.c_ref(), None, adj.data_ptr(), num_nodes.data_ptr(),
                                                      _cabi.stream_ptr(self.device)), "gcm_state_materialize")
            return nodes, adj, num_nodes
        nodes = torch.empty(self.B, self.N, self.F, device=self.device, dtype=torch.float32)
        adj = torch.empty(self.B, self.N, self.N, device=self.device, dtype=torch.float32) if want_adj else None
        num_nodes = torch.empty(self.B, device=self.device, dtype=torch.long)
        _cabi.check(
            _cabi.lib().gcm_state_materialize(self.c_ref(), nodes.data_ptr(), _cabi.ptr(adj),
                                              num_nodes.data_ptr(), _cabi.stream_ptr(self.device)),
            "gcm_state_materialize")
        return nodes, adj, num_nodes

    def materialize_weights(self) -> torch.Tensor:
        """Edge weights as the reference would return them: untouched by the deterministic selectors
        (gcm.py:286,321) but shifted by wrap_overflow (gcm.py:346-352) once per overflowing step."""
        if self.weights0 is None:
            return torch.zeros(0, device=self.device)
        w0 = self.weights0
        N = self.N
        shift = (self.count.long() - N).clamp(min=0)                      # [B] overflow steps so far
        if not bool((shift > 0).any()):
            return w0.clone()
        idx = torch.arange(N, device=self.device).view(1, N) + shift.view(-1, 1)   # source row/col
        ok = idx < N
        idx = idx.clamp(max=N - 1)
        out = w0.gather(1, idx.view(-1, N, 1).expand(-1, N, N)).gather(2, idx.view(-1, 1, N).expand(-1, N, N))
        keep = ok.view(-1, N, 1) & ok.view(-1, 1, N)
        return out * keep.to(out.dtype)

    def check_flags(self) -> None:
        """Lazy poll of the device status word (one host sync).  Keeps the reference's error text."""
        flags = int(self.status[0].item())
        if flags & _cabi.FLAG_NONFINITE:
            self.status.zero_()
            raise AssertionError("Got NaN in returned memory, try using tanh activation")


class DenseHidden:
    """`m_t`: a versioned handle on a DenseState that behaves like the reference's 4-tuple."""

    __slots__ = ("_state", "_version", "_cache", "token")

    def __init__(self, state: DenseState, token=None):
        self._state = state
        self._version = state.version
        self._cache: Optional[Tuple[torch.Tensor, ...]] = None
        self.token = token  # autograd chain token (training), see gcm.fused

    # -- used by DenseGCM -------------------------------------------------------------------------
    def live(self) -> bool:
        return self._version == self._state.version

    def claim(self) -> DenseState:
        if not self.live():
            raise RuntimeError(
                "stale GCM hidden state: it was superseded by a later in-place step. Pass the most "
                "recent m_t, or pin an old one first with tuple(m_t) / m_t.snapshot()."
            )
        return self._state

    # -- reference-tuple behaviour ------------------------------------------------------------------
    def snapshot(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        if self._cache is None:
            st = self.claim()
            st.check_flags()
            nodes, adj, num_nodes = st.materialize()
            self._cache = (nodes, adj, st.materialize_weights(), num_nodes)
        return self._cache

    def __iter__(self):
        return iter(self.snapshot())

    def __len__(self):
        return 4

    def __getitem__(self, i):
        return self.snapshot()[i]

    def detach(self) -> "DenseHidden":
        """Same graph state, cut from the autograd history (truncated BPTT)."""
        return DenseHidden(self.claim(), None)

    def clone(self) -> "DenseHidden":
        """An independent copy of the memory (detached): later steps on either handle do not affect the other.  Use it to
        restart several episodes / BPTT windows from one prepared state."""
        return DenseHidden(self.claim().clone(), None)

    @property
    def num_nodes(self) -> torch.Tensor:
        st = self.claim()
        return st.count.long().clamp(max=st.N)

    def __repr__(self):
        st = self._state
        return (f"DenseHidden(B={st.B}, N={st.N}, F={st.F}, capacity={st.C}, "
                f"{'live' if self.live() else 'stale'})")
