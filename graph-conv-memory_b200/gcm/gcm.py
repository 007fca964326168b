"""DenseGCM — drop-in for the reference's `gcm.gcm.DenseGCM` (/root/reference/src/gcm/gcm.py:151-355).

Same constructor, same `forward(x[B,F], hidden) -> (belief[B,H], hidden)` contract, same error
behaviour.  When the configuration is one the hot path covers (two DenseGraphConv layers +
activation, TemporalBackedge / DenseEdge / EuclideanEdge / CosineEdge / SpatialEdge selectors,
see gcm.fused.build_plan) a step is ONE kernel launch on an in-place, bit-packed graph state
(`gcm_dense_step_fwd`); the returned `hidden` is a `DenseHidden` that unpacks like the
reference's (nodes, adj, weights, num_nodes) tuple.  Everything else (learned edges,
preprocessors, positional encoders, arbitrary user GNNs) takes `_forward_generic`, which
evaluates the reference's step semantics with torch ops on whatever device the tensors live on.
The fused path is CUDA-only and has no CPU fallback.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple, Union

import torch

import gcm.util
from gcm import _cabi, fused, ones, temporal
from gcm.state import DenseHidden, DenseState


class SparseToDense(torch.nn.Module):
    """edge_index -> dense adjacency (reference gcm.py:10-21; legacy adapter, torch ops)."""

    def forward(self, x, edge_index, batch_idx, B, N):
        feat = x.shape[-1]
        dense_x = x.new_zeros(B * N, feat)
        counts = torch.bincount(batch_idx, minlength=B)
        starts = torch.cumsum(counts, 0) - counts
        pos = torch.arange(batch_idx.numel(), device=x.device) - starts[batch_idx]
        dense_x[batch_idx * N + pos] = x
        adj = x.new_zeros(B, N, N)
        eb = batch_idx[edge_index[0]]
        adj.index_put_((eb, edge_index[0] - starts[eb], edge_index[1] - starts[eb]),
                       x.new_ones(eb.numel()), accumulate=True)
        return dense_x.view(B, N, feat), adj


class DenseToSparse(torch.nn.Module):
    """dense adjacency -> edge_index (reference gcm.py:24-53; legacy adapter, torch ops)."""

    def forward(self, x, adj, mask=None):
        assert x.dim() == adj.dim() == 3
        if mask:
            raise NotImplementedError()
        B, N = x.shape[0], x.shape[1]
        b, row, col = torch.nonzero(adj > 0).t()
        edge_index = torch.stack([row + b * N, col + b * N], dim=0).long()
        batch_idx = torch.arange(B, device=x.device).repeat_interleave(N)
        return x.reshape(B * N, x.shape[-1]), edge_index, batch_idx


def _sincos_table(max_len: int, width: int, device) -> torch.Tensor:
    d_model = math.ceil(width / 2) * 2
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model, device=device)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


class RelativePositionalEncoding(torch.nn.Module):
    """Reference gcm.py:56-89 (only used together with aux_edge_selectors; generic path)."""

    def __init__(self, max_len: int = 5000):
        super().__init__()
        self.max_len = max_len

    def forward(self, nodes: torch.Tensor, num_nodes: torch.Tensor) -> torch.Tensor:
        if not hasattr(self, "pe"):
            self.register_buffer("pe", _sincos_table(self.max_len, nodes.shape[-1], nodes.device))
        for b in range(nodes.shape[0]):
            c = int(num_nodes[b])
            pe = self.pe.roll(c, 0)
            nodes[b, : c + 1] = nodes[b, : c + 1] + pe[: c + 1, : nodes.shape[-1]]
        return nodes


class PositionalEncoding(torch.nn.Module):
    """Reference gcm.py:92-143: sinusoidal encoding added to / concatenated onto rows <= num_nodes."""

    def __init__(self, max_len: int = 5000, mode="add", cat_dim: int = 8):
        super().__init__()
        assert mode in ["add", "cat"]
        self.max_len, self.mode, self.cat_dim = max_len, mode, cat_dim

    def forward(self, x: torch.Tensor, num_nodes: torch.Tensor) -> torch.Tensor:
        if not hasattr(self, "pe"):
            self.register_buffer("pe", _sincos_table(self.max_len, x.shape[-1], x.device))
            if self.mode == "cat":
                self.reproject = torch.nn.Linear(x.shape[-1], x.shape[-1] - self.cat_dim, device=x.device)
        b_idxs, n_idxs = gcm.util.idxs_up_to_including_num_nodes(x, num_nodes)
        if self.mode == "add":
            x[b_idxs, n_idxs] = x[b_idxs, n_idxs] + self.pe[n_idxs, : x.shape[-1]]
            return x
        x_reproj = self.reproject(x[b_idxs, n_idxs]).reshape(len(b_idxs), x.shape[-1] - self.cat_dim)
        x = x.clone()
        x[b_idxs, n_idxs, : self.cat_dim] = self.pe[n_idxs, : self.cat_dim]
        x[b_idxs, n_idxs, self.cat_dim:] = x_reproj
        return x


def overflow(num_nodes: torch.Tensor, N: int):
    return torch.any(num_nodes + 1 > N)


def _apply_pre(pre, t: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The row-wise preprocessor on a block of observations (no autograd).  A plain float32 Linear -- RayDenseGCM's,
    ray_gcm.py:118 -- runs on the library's own product kernel (3xTF32 on the tensor cores when both widths are
    multiples of 16): cuBLAS picked a SIMT sgemm plus a separate bias kernel for these narrow shapes, 16 us per cfg2-pre
    step against 4.  out: a contiguous float32 tensor of the result's shape to write into."""
    if (type(pre) is torch.nn.Linear and t.is_cuda and t.dtype is torch.float32 and t.is_contiguous()
            and pre.weight.dtype is torch.float32 and pre.weight.is_cuda and pre.out_features <= 128
            and pre.in_features <= 128 and t.numel() > 0):
        a = t.detach().reshape(-1, pre.in_features)
        w = pre.weight.detach()
        bias = None if pre.bias is None else pre.bias.detach()
        o2 = None
        if out is not None and out.is_contiguous() and out.dtype is torch.float32 and out.numel() == a.shape[0] * pre.out_features:
            o2 = out.view(-1, pre.out_features)
        if pre.in_features % 16 == 0 and pre.out_features % 16 == 0:
            y = ones._lin_tc32(a, w, bias=bias, out=o2)
        else:
            y = ones._lin2(a, w, bias=bias, out=o2)
        return out if o2 is not None else y.view(*t.shape[:-1], pre.out_features)
    y = pre(t)
    if out is not None:
        out.copy_(y)
        return out
    return y


class _PreLinearFn(torch.autograd.Function):
    """y = x W^T + b for a row-wise Linear preprocessor, with autograd, on the library's 3xTF32 kernels (forward and
    dL/dx: gcm_linear_tc32, dL/dW and dL/db: gcm_outer_reduce_tc32).  The training window of RayDenseGCM spent 5 of its 12
    ms in cuBLAS' SIMT sgemm + bias kernels for these 32-wide products.  A [B, T, F] view of time-major data is
    processed (and returned) in its memory order."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        tm = x.dim() == 3 and not x.is_contiguous() and x.transpose(0, 1).is_contiguous()
        xc = (x.transpose(0, 1) if tm else x.contiguous()).detach()
        a = xc.reshape(-1, weight.shape[1])
        y = ones._lin_tc32(a, weight.detach(), bias=None if bias is None else bias.detach())
        ctx.save_for_backward(a, weight)
        ctx.tm, ctx.shape, ctx.has_bias = tm, xc.shape, bias is not None
        y = y.view(*xc.shape[:-1], weight.shape[0])
        return y.transpose(0, 1) if tm else y

    @staticmethod
    def backward(ctx, dy):
        a, weight = ctx.saved_tensors
        dyc = (dy.transpose(0, 1) if ctx.tm else dy).contiguous().float()
        d2 = dyc.reshape(-1, weight.shape[0])
        dx = dw = db = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw = torch.zeros_like(weight)
            db = torch.zeros(weight.shape[0], device=weight.device, dtype=torch.float32)
            temporal._outer(d2, a, dw, db)
            if not ctx.has_bias:
                db = None
        if ctx.needs_input_grad[0]:
            dx = ones._lin_tc32(d2, weight.detach().t().contiguous()).view(ctx.shape)
            if ctx.tm:
                dx = dx.transpose(0, 1)
        return dx, dw, db


def _pre_own_kernels(pre, t: torch.Tensor) -> bool:
    return (type(pre) is torch.nn.Linear and t.is_cuda and t.dtype is torch.float32 and pre.weight.dtype is torch.float32
            and pre.weight.is_cuda and pre.in_features % 16 == 0 and pre.out_features % 16 == 0
            and pre.in_features <= 128 and pre.out_features <= 128 and t.numel() > 0)


def _apply_pre_grad(pre, t: torch.Tensor) -> torch.Tensor:
    """The preprocessor with autograd (training windows)."""
    if _pre_own_kernels(pre, t) and torch.is_grad_enabled():
        return _PreLinearFn.apply(t, pre.weight, pre.bias)
    return pre(t)


class DenseGCM(torch.nn.Module):
    """Graph Associative Memory"""

    did_warn = False

    def __init__(
        self,
        gnn: torch.nn.Module,
        preprocessor: torch.nn.Module = None,
        edge_selectors: torch.nn.Module = None,
        aux_edge_selectors: torch.nn.Module = None,
        graph_size: int = 128,
        pooled: bool = False,
        positional_encoder: torch.nn.Module = None,
        edge_weights: bool = False,
        bptt_capacity: int = 128,
    ):
        """Arguments and defaults as the reference's (gcm.py:156-182).  bptt_capacity (extension, keyword only in
        practice): the longest BPTT window, in steps, whose backward the fused paths can recompute from the node log."""
        super().__init__()
        self.preprocessor = preprocessor
        self.gnn = gnn
        self.graph_size = graph_size
        self.edge_selectors = edge_selectors
        self.aux_edge_selectors = aux_edge_selectors
        self.pooled = pooled
        self.edge_weights = edge_weights
        self.positional_encoder = positional_encoder
        self._plan = None
        self._plan_built = False
        # extra log rows kept while autograd is recording, so that a BPTT window of up to this many
        # steps can be recomputed in backward without any per-step saved activations
        self.bptt_capacity = int(bptt_capacity)
        # None (float32 everywhere, 1e-5 parity) or torch.bfloat16: DenseEdge-only states keep their per-node cache in
        # bfloat16 (BASELINE cfg3's precision, 2e-2 parity); torch.autocast("cuda", dtype=torch.bfloat16) does the same
        self.compute_dtype = None
        # Batch sharding (one process per GPU, every rank owns a slice of the graphs): the hot path needs no exchange
        # EXCEPT for EuclideanEdge, whose distance averages over the current observation of every graph of the batch
        # (reference edge_selectors/distance.py:48-49).  Set batch_group to the torch.distributed group the batch is
        # sharded over (True = the default group) and every step all-gathers the ranks' observations first, so that
        # sharded results equal the unsharded reference's.  None: this process holds the whole batch.
        self.batch_group = None
        self._shard_sizes = None

    # ------------------------------------------------------------------ reference API
    def get_initial_hidden_state(self, x):
        """Zeros hidden state in the reference layout (gcm.py:194-211)."""
        assert x.dim() == 2
        B, feats = x.shape
        edges = torch.zeros(B, self.graph_size, self.graph_size, device=x.device)
        nodes = torch.zeros(B, self.graph_size, feats, device=x.device)
        if self.edge_weights:
            weights = torch.zeros(B, self.graph_size, self.graph_size, device=x.device)
        else:
            weights = torch.zeros(0, device=x.device)
        num_nodes = torch.zeros(B, dtype=torch.long, device=x.device)
        return nodes, edges, weights, num_nodes

    def fused_plan(self):
        if not self._plan_built:
            self._plan = fused.build_plan(self)
            self._plan_built = True
        return self._plan

    def forward(self, x, hidden) -> Tuple[torch.Tensor, Union[DenseHidden, Tuple[torch.Tensor, ...]]]:
        """Add observation x [B,F] to the graph and query the memory for it.

        hidden: None, the previous call's returned hidden, or a reference-style tuple
        (nodes[B,N,F], adj[B,N,N], weights [B,N,N] or empty, num_nodes[B] int64).
        Returns (belief [B,H], hidden)."""
        if hidden.__class__ is DenseHidden and not torch.is_grad_enabled():
            # steady-state rollout: skip everything that cannot have changed since the previous step
            st = hidden._state
            plan = self._plan
            if (st.fast_ok and hidden._version == st.version and hidden.token is None and plan is not None
                    and plan.validated and not plan.pre):
                belief = temporal.fast_step(plan, st, x)
                if belief is not None:
                    if not DenseGCM.did_warn and st.host_count is not None and st.host_count > st.N:
                        print("Overflow detected, wrapping around. Will not warn again")
                        DenseGCM.did_warn = True
                    return belief, DenseHidden(st, None)
        plan = self.fused_plan()
        if plan is None:
            return self._forward_generic(x, hidden)
        if plan.pre:
            return self._forward_pre(plan, x, hidden)
        return self._forward_fused(plan, x, hidden)

    def _forward_pre(self, plan, x, hidden):
        """Fused rollout step of a DenseGCM with a row-wise preprocessor (what RayDenseGCM installs, ray_gcm.py:118,
        133-136).  The reference maps ALL N rows through the preprocessor every step (gcm.py:290-291); a per-row map only
        has to be applied to the NEW observation, whose image goes into the node log the kernels read, while a second log
        keeps the raw observations for the caller's view of m_t.  When the preprocessor's weights change the
        preprocessed log is rebuilt from the raw one.  Anything that needs autograd takes the generic path (exact
        reference semantics: gradients reach the preprocessor through every stored row)."""
        pre = self.preprocessor
        if torch.is_grad_enabled() and (
                x.requires_grad or any(p.requires_grad for p in self.parameters())
                or (isinstance(hidden, DenseHidden) and hidden.token is not None)
                or (isinstance(hidden, (tuple, list)) and hidden[0].requires_grad)):
            res = self._sequence_pre_grad(plan, x.unsqueeze(1), hidden) if x.dim() == 2 else None
            if res is not None:
                return res[0][:, 0], res[1]
            return self._forward_generic(x, hidden)
        assert x.dtype == torch.float32
        _cabi.require_cuda(x, "DenseGCM.forward(x)")
        with torch.no_grad():
            x_raw = x.contiguous()
            plist = self.__dict__.get("_pre_params")
            if plist is None:
                plist = self.__dict__["_pre_params"] = list(pre.parameters())
            pkey = tuple([v for p in plist for v in (p.data_ptr(), p._version)])
            nodes_raw = None
            if isinstance(hidden, DenseHidden):
                state = hidden.claim()
                if state.raw is None:
                    return self._forward_generic(x, hidden)
                if state.pre_key != pkey:
                    # new preprocessor weights: every stored row gets its new image, every cache built on them goes
                    state.nodes.copy_(pre(state.raw))
                    state.pre_key = pkey
                    state.xsum, state.rc_key, state.hc_key, state.hc_fresh = None, None, None, 0
                    state.fast_ok = False
                elif state.fast_ok and hidden.token is None and plan.validated and self._plan is plan:
                    # steady-state rollout: the preprocessor on the new observation (our own Linear kernel when it is
                    # a plain Linear <= 128 wide), the trimmed step (gcm.fused.fast_temporal_step), the raw-log write
                    if (type(pre) is torch.nn.Linear and pre.out_features <= 128 and pre.weight.dtype == torch.float32
                            and pre.weight.is_cuda):
                        y = ones._lin2(x_raw, pre.weight, bias=pre.bias)
                    else:
                        y = pre(x_raw)
                    belief = temporal.fast_step(plan, state, y)
                    if belief is not None:
                        _cabi.check(_cabi.lib().gcm_state_log_write(state.raw_ref(), x_raw.data_ptr(), -1,
                                                                    _cabi.stream_ptr(x.device)), "gcm_state_log_write")
                        return belief, DenseHidden(state, None)
                inner = hidden
            elif hidden is None:
                inner = None
            else:
                nodes_raw, adj, weights, num_nodes = hidden
                assert nodes_raw.dtype == torch.float
                inner = (pre(nodes_raw), adj, weights, num_nodes)
            belief, out = self._forward_fused(plan, pre(x_raw), inner, orig=(x, hidden))
            if not isinstance(out, DenseHidden) or self._plan is not plan:
                return belief, out
            state = out.claim()
            if state.raw is None:
                state.raw = torch.zeros(state.B, state.C, x_raw.shape[1], device=x.device, dtype=torch.float32)
                if nodes_raw is not None:
                    state.raw[:, : state.N] = nodes_raw
                state.pre_key = pkey
            _cabi.check(_cabi.lib().gcm_state_log_write(state.raw_ref(), x_raw.data_ptr(), -1,
                                                        _cabi.stream_ptr(x.device)), "gcm_state_log_write")
        return belief, out

    def _forward_fused(self, plan, x, hidden, orig=None):
        """The fused step on GNN-input rows x.  `orig` = the caller's (x, hidden) when x is a preprocessed image: what
        the generic path must be given if this configuration turns out not to be fusable."""
        def generic():
            return self._forward_generic(*(orig if orig is not None else (x, hidden)))

        assert x.dtype == torch.float32
        _cabi.require_cuda(x, "DenseGCM.forward(x)")
        assert x.dim() == 2 and x.shape[1] == plan.gnn.F, "x must be [B, obs_size] matching the GNN input"
        B = x.shape[0]
        ingest_grad = False
        recording = torch.is_grad_enabled() and (
            x.requires_grad or any(p.requires_grad for p in plan.gnn.params())
            or (isinstance(hidden, DenseHidden) and hidden.token is not None))

        if hidden is None:
            state = DenseState(B, self.graph_size, x.shape[1], x.device,
                               self.graph_size + (self.bptt_capacity if recording else 0))
            token = None
        elif isinstance(hidden, DenseHidden):
            state = hidden.claim()
            token = hidden.token
            assert state.B == B and state.F == x.shape[1]
        else:
            nodes, adj, weights, num_nodes = hidden
            assert nodes.dtype == torch.float
            assert weights.dtype == torch.float
            assert num_nodes.dtype == torch.long
            assert num_nodes.dim() == 1
            N = nodes.shape[1]
            assert N == adj.shape[1] == adj.shape[2], "N must be equal for adj mat and node mat"
            if adj.requires_grad or (weights.numel() != 0 and weights.requires_grad):
                return generic()   # learned / weighted adjacency
            state, flags = DenseState.ingest(nodes, adj, weights, num_nodes,
                                             N + (self.bptt_capacity if recording else 0))
            if flags & (_cabi.FLAG_UNCLEAN | _cabi.FLAG_BADCOUNT):
                return generic()   # not a {0,1} graph over the valid block
            fused.recognise_pure_temporal(plan, state, adj, num_nodes)     # a memory this chain built keeps its fast path
            token = None
            if recording and nodes.requires_grad:
                ingest_grad = True
                token = fused.ingest_token(state, nodes)
            recording = recording or token is not None

        if not DenseGCM.did_warn and self._would_overflow(state):
            print("Overflow detected, wrapping around. Will not warn again")
            DenseGCM.did_warn = True

        xc = x.contiguous()
        if plan.needs_euclid and not (plan.ones and state.dense_ok):
            state.__dict__["_euclid_cur_all"] = self._all_current_obs(xc)
        if plan.ones and state.dense_ok and not ingest_grad and (token is None or getattr(token, "_gcm_ones", False)):
            # DenseEdge-only state: implicit all-ones adjacency, per-node cache (gcm.ones)
            if recording:
                if state.C - state.N < 1:
                    state = fused.grow_state(state, state.N + max(int(self.bptt_capacity), 1))
                    token = None
                belief, token = ones.step_grad(plan, state, xc, token, ones.want_bf16(self, plan))
                token._gcm_ones = True
            else:
                belief = ones.step_nograd(plan, state, xc.detach(), ones.want_bf16(self, plan))
                token = None
        elif recording:
            if (not ingest_grad and (token is None or getattr(token, "_gcm_tw", False))
                    and temporal.grad_supported(plan, state)):
                # forward-only temporal chain on a state it built itself: window-level backward (gcm.temporal)
                belief, token, state = temporal.step_grad(plan, state, xc, token, self.bptt_capacity)
            else:
                belief, token, state = fused.fused_step_grad(plan, state, xc, token, self.bptt_capacity)
        else:
            belief = None
            if plan.zc and state.zc_ok:
                # one distance selector, rollout: per-node pre-activation cache (csrc/gcm_dense_zc.cu)
                belief = fused.zc_step(plan, state, xc.detach())
            if belief is None:
                belief = fused.fused_step_nograd(plan, state, xc.detach())
            token = None
        if not plan.validated:
            plan.validated = True
            if not fused.validate_plan(plan, self, state, belief):
                self._plan = None
                nodes, adj, num_nodes = state.materialize()
                feats = self.gnn(nodes, adj, state.materialize_weights(), B, state.N)
                belief = feats[torch.arange(B, device=x.device), num_nodes - 1]
        return belief, DenseHidden(state, token)

    def forward_sequence(self, x_seq, hidden, time_major: bool = False):
        """T steps at once: x_seq [B, T, F] -> (beliefs [B, T, H], hidden).  Same results as
        `for t in range(T): belief_t, hidden = self(x_seq[:, t], hidden)` (the loop of RayDenseGCM.forward,
        reference ray_gcm.py:200-202), which is also what runs for configurations without a fused sequence kernel.
        Forward-only TemporalBackedge chains take the T steps in ONE C call (gcm.temporal: the step kernels are enqueued
        from C, consecutive cached-row steps in a single multi-step launch); DenseEdge-only states in five launches
        (gcm.ones: the per-node cache is read once).
        time_major=True: x_seq is [T, B, F] and the beliefs come back as [T, B, H] (every step's rows contiguous, like
        torch.nn.RNN's batch_first=False)."""
        if time_major:
            assert x_seq.dim() == 3, "x_seq must be [T, B, obs_size]"
            beliefs, hidden = self._forward_sequence(x_seq.transpose(0, 1), hidden, True)
            return beliefs.transpose(0, 1), hidden
        return self._forward_sequence(x_seq, hidden, False)

    def _forward_sequence(self, x_seq, hidden, tm_out):
        assert x_seq.dim() == 3, "x_seq must be [B, T, obs_size]"
        T = x_seq.shape[1]

        def loop(hidden):
            outs = []
            for t in range(T):
                out, hidden = self(x_seq[:, t], hidden)
                outs.append(out)
            return torch.stack(outs, dim=1), hidden

        plan = self.fused_plan()
        if plan is None or T < 2 or not x_seq.is_cuda or x_seq.dtype != torch.float32:
            return loop(hidden)
        if plan.temporal_key is not None and plan.hc_ring:
            return self._sequence_temporal(plan, x_seq, hidden, loop, tm_out)
        if not plan.ones or plan.pre:
            # (a DenseEdge state behind a preprocessor keeps two logs; its sequence kernels only know one)
            return loop(hidden)
        recording = torch.is_grad_enabled() and (
            x_seq.requires_grad or any(p.requires_grad for p in plan.gnn.params())
            or (isinstance(hidden, DenseHidden) and hidden.token is not None))
        if recording and T > int(self.bptt_capacity):
            return loop(hidden)
        if (hidden.__class__ is DenseHidden and hidden.live() and plan.validated and self._plan is plan
                and x_seq.is_contiguous()):
            # a live handle of a validated plan: nothing to ingest or validate, all T steps go through the sequence
            # kernels -- no separate first step, no concatenation of its belief with the others, and the [B, T, H] result
            # is a VIEW of the time-major buffer the kernels write
            st = hidden._state
            tok = hidden.token
            bf16 = ones.want_bf16(self, plan)
            cap = st.C - st.N + 1
            whole = (st.dense_ok and st.B == x_seq.shape[0] and st.F == x_seq.shape[2]
                     and (tok is None or getattr(tok, "_gcm_ones", False)) and ones.sequence_supported(plan, st, T, bf16))
            if whole and recording:
                win = getattr(st, "win", None)
                used = 0 if (tok is None or win is None) else st.steps - win.chain_start
                whole = st.C - st.N >= 1 and used + T <= cap
            elif whole:
                whole = tok is None
            if whole:
                st = hidden.claim()
                if not DenseGCM.did_warn and st.host_count is not None and st.host_count + T > st.N:
                    print("Overflow detected, wrapping around. Will not warn again")
                    DenseGCM.did_warn = True
                if recording:
                    beliefs, tok = ones.sequence_grad(plan, st, x_seq, tok, bf16)
                    tok._gcm_ones = True
                else:
                    beliefs, tok = ones.sequence_nograd(plan, st, x_seq.detach(), bf16), None
                return beliefs.transpose(0, 1), DenseHidden(st, tok)
        # enter the state exactly as forward() does, by taking the first step through it
        out0, hidden = self(x_seq[:, 0], hidden)
        state = hidden.claim() if isinstance(hidden, DenseHidden) else None
        token = hidden.token if state is not None else None
        bf16 = ones.want_bf16(self, plan)
        ok = (state is not None and state.dense_ok and state.rcache is not None and self._plan is plan
              and (token is None) == (not recording) and (token is None or getattr(token, "_gcm_ones", False))
              and ones.sequence_supported(plan, state, T - 1, bf16))
        if ok and recording:
            ok = state.steps - state.win.chain_start + (T - 1) <= state.C - state.N + 1
        if not ok:
            outs = [out0]
            for t in range(1, T):
                out, hidden = self(x_seq[:, t], hidden)
                outs.append(out)
            return torch.stack(outs, dim=1), hidden
        rest = x_seq[:, 1:].contiguous()
        if not DenseGCM.did_warn and state.host_count is not None and state.host_count + T - 1 > state.N:
            print("Overflow detected, wrapping around. Will not warn again")
            DenseGCM.did_warn = True
        if recording:
            beliefs, token = ones.sequence_grad(plan, state, rest, token, bf16)
            token._gcm_ones = True
        else:
            beliefs = ones.sequence_nograd(plan, state, rest.detach(), bf16)
            token = None
        return torch.cat([out0.unsqueeze(1), beliefs.transpose(0, 1)], dim=1), DenseHidden(state, token)

    def _sequence_temporal(self, plan, x_seq, hidden, loop, tm_out=False):
        """forward_sequence for forward-only TemporalBackedge chains (with or without a row-wise preprocessor): the T
        steps are enqueued by ONE C call (gcm.temporal.sequence_nograd -> gcm_dense_rollout_fwd), which reads x_seq
        [B, T, F] and writes the [B, T, H] result in place.  Anything that records autograd, and any state that is not a
        pure temporal state of this chain, takes the step loop."""
        T = x_seq.shape[1]
        pre = self.preprocessor if plan.pre else None
        if torch.is_grad_enabled() and (
                x_seq.requires_grad or any(p.requires_grad for p in self.parameters())
                or (isinstance(hidden, DenseHidden) and hidden.token is not None)
                or (isinstance(hidden, (tuple, list)) and hidden[0].requires_grad)):
            if pre is not None:
                res = self._sequence_pre_grad(plan, x_seq, hidden)
                return res if res is not None else loop(hidden)
            if T > int(self.bptt_capacity):
                return loop(hidden)
            # recording: one autograd node for the T steps, window-level backward (gcm.temporal)
            out0, start = None, 0
            if not (hidden.__class__ is DenseHidden and hidden.live()):
                out0, hidden = self(x_seq[:, 0], hidden)     # enter the state exactly as forward() does
                start = 1
            state = hidden._state if (hidden.__class__ is DenseHidden and hidden.live()) else None
            token = hidden.token if state is not None else None
            ok = (state is not None and self._plan is plan and temporal.grad_supported(plan, state)
                  and (token is None or getattr(token, "_gcm_tw", False)) and x_seq.dtype == torch.float32)
            if ok and token is not None:
                ok = state.steps - state.twin.chain_start + (T - start) <= state.C - state.N + 1
            if not ok:
                outs = [] if out0 is None else [out0]
                for t in range(start, T):
                    out, hidden = self(x_seq[:, t], hidden)
                    outs.append(out)
                return torch.stack(outs, dim=1), hidden
            if not DenseGCM.did_warn and state.host_count + (T - start) > state.N:
                print("Overflow detected, wrapping around. Will not warn again")
                DenseGCM.did_warn = True
            beliefs, token, state = temporal.sequence_grad(plan, state, x_seq[:, start:], token, self.bptt_capacity)
            if out0 is not None:
                beliefs = torch.cat([out0.unsqueeze(1), beliefs], dim=1)
            return beliefs, DenseHidden(state, token)
        with torch.no_grad():
            start = 0
            out0 = None
            state = hidden._state if (hidden.__class__ is DenseHidden and hidden.live() and hidden.token is None) else None
            if state is None or not plan.validated or self._plan is not plan or state.rollout is None:
                # enter the state exactly as forward() does, by taking the first step through it
                out0, hidden = self(x_seq[:, 0], hidden)
                start = 1
                state = hidden._state if (hidden.__class__ is DenseHidden and hidden.live()) else None
            if pre is not None and state is not None:
                plist = self.__dict__.get("_pre_params")
                if plist is None:
                    plist = self.__dict__["_pre_params"] = list(pre.parameters())
                pkey = tuple([v for p in plist for v in (p.data_ptr(), p._version)])
                if state.raw is None or state.pre_key != pkey:
                    state = None                 # forward() rebuilds the preprocessed log under the new weights
            rest_raw = x_seq[:, start:]
            rest = rest_raw
            if pre is not None:
                rest = None
                if state is not None:
                    # map the observations in the layout they came in: a time-major caller ([T, B, F] behind this
                    # transposed view) gets time-major images, i.e. contiguous rows per step for the kernels
                    tm = rest_raw.transpose(0, 1)
                    rest = _apply_pre(pre, tm).transpose(0, 1) if tm.is_contiguous() else _apply_pre(pre, rest_raw)
            if (state is None or self._plan is not plan or rest.dtype != torch.float32
                    or not temporal.sequence_supported(plan, state, rest)):
                outs = [] if out0 is None else [out0]
                for t in range(start, T):
                    out, hidden = self(x_seq[:, t], hidden)
                    outs.append(out)
                return torch.stack(outs, dim=1), hidden
            if tm_out:      # the caller wants [T, B, H]: every step's belief rows contiguous
                beliefs = torch.empty(T, state.B, plan.gnn.H2, device=x_seq.device, dtype=torch.float32).transpose(0, 1)
            else:
                beliefs = torch.empty(state.B, T, plan.gnn.H2, device=x_seq.device, dtype=torch.float32)
            if out0 is not None:
                beliefs[:, 0] = out0
            if not DenseGCM.did_warn and state.host_count is not None and state.host_count + (T - start) > state.N:
                print("Overflow detected, wrapping around. Will not warn again")
                DenseGCM.did_warn = True
            temporal.sequence_nograd(plan, state, rest, beliefs[:, start:])
            if pre is not None:
                _cabi.check(_cabi.lib().gcm_state_log_write_seq(
                    state.raw_ref(), rest_raw.data_ptr(), rest_raw.stride(0), rest_raw.stride(1), T - start,
                    _cabi.stream_ptr(x_seq.device)), "gcm_state_log_write_seq")
            return beliefs, DenseHidden(state, None)

    def _all_current_obs(self, x):
        """EuclideanEdge under batch sharding: the current observations of EVERY rank's graphs, in batch order (None when
        this process holds the whole batch)."""
        if self.batch_group is None:
            return None
        import torch.distributed as tdist
        from gcm import dist as gdist
        if not tdist.is_initialized():
            return None
        group = None if self.batch_group is True else self.batch_group
        if self._shard_sizes is None or self._shard_sizes[0] != x.shape[0]:
            self._shard_sizes = (x.shape[0], gdist.shard_sizes(x.shape[0], x.device, group))
        return gdist.gather_current_obs(x.detach(), group, self._shard_sizes[1])

    def _sequence_pre_grad(self, plan, x_seq, hidden):
        """Recorded steps of a forward-only TemporalBackedge chain behind a row-wise preprocessor (RayDenseGCM's
        configuration in training) on the window-level backward.  The reference maps ALL stored rows through the
        preprocessor at every step (gcm.py:290-291); a per-row map only has to be applied to the new observations
        (y = pre(x), with autograd: dL/dy comes back from the window's backward) and -- for its parameters' gradient
        through the rows written BEFORE the window -- to the 2 max_hop newest stored raw rows, which enter the autograd
        graph as the `hist` input of gcm.temporal._TSeqFn.  Returns None when the situation is not covered (the caller
        then takes the generic path)."""
        pre = self.preprocessor
        if (not plan.hc_ring or plan.temporal_key is None or x_seq.dtype != torch.float32 or not x_seq.is_cuda
                or x_seq.dim() != 3 or self._plan is not plan):
            return None
        B, T, F_raw = x_seq.shape
        dev = x_seq.device
        mh, N = plan.max_hop, self.graph_size
        cap = max(int(self.bptt_capacity), T)
        plist = self.__dict__.get("_pre_params")
        if plist is None:
            plist = self.__dict__["_pre_params"] = list(pre.parameters())
        pkey = tuple([v for p in plist for v in (p.data_ptr(), p._version)])
        token = None
        with torch.no_grad():
            if hidden is None:
                state = DenseState(B, N, plan.gnn.F, dev, N + cap)
                state.raw = torch.zeros(B, state.C, F_raw, device=dev, dtype=torch.float32)
                state.pre_key = pkey
            elif isinstance(hidden, DenseHidden):
                state, token = hidden.claim(), hidden.token
                if state.raw is None or state.B != B or state.raw.shape[2] != F_raw:
                    return None
                if state.pre_key != pkey:
                    if token is not None:
                        raise RuntimeError("preprocessor parameters were modified in place inside a recorded window")
                    with torch.no_grad():
                        hc = state.host_count
                        if hc is not None and plan.temporal_key is not None and state.pure_key == plan.temporal_key:
                            # new weights: the images of the stored rows are stale.  A forward-only temporal chain only
                            # ever reads the last 2 max_hop nodes again (the step kernels max_hop back, the recomputing
                            # kernel and the window backward 2 max_hop): those get their new image, not the whole log
                            # (12.6 M rows at cfg2: 1.1 ms per training window)
                            pos = torch.arange(max(0, hc - 2 * mh), hc, device=dev)
                            if pos.numel():
                                slots = pos % state.C
                                state.nodes[:, slots] = _apply_pre(pre, state.raw[:, slots].contiguous())
                        else:
                            _apply_pre(pre, state.raw, out=state.nodes)   # every stored row gets its new image
                    state.pre_key = pkey
                    state.xsum, state.rc_key, state.hc_key, state.hc_fresh, state.fast_ok = None, None, None, 0, False
                if token is None and (state.C - state.N + 1 < T or state.C - state.N < 1):
                    state = self._rehome_pre(plan, state, state.N + cap, pkey)
            else:
                nodes_raw, adj, weights, num_nodes = hidden
                if nodes_raw.requires_grad or adj.requires_grad or nodes_raw.dtype != torch.float32 or not nodes_raw.is_cuda:
                    return None
                N = nodes_raw.shape[1]
                state, flags = DenseState.ingest(pre(nodes_raw), adj, weights, num_nodes, N + cap)
                if flags & (_cabi.FLAG_UNCLEAN | _cabi.FLAG_BADCOUNT):
                    return None
                fused.recognise_pure_temporal(plan, state, adj, num_nodes)
                state.raw = torch.zeros(B, state.C, F_raw, device=dev, dtype=torch.float32)
                state.raw[:, :N] = nodes_raw
                state.pre_key = pkey
            if (state is None or not temporal.grad_supported(plan, state)
                    or (token is not None and not getattr(token, "_gcm_tw", False))):
                return None
            if token is not None and state.steps - state.twin.chain_start + T > state.C - state.N + 1:
                return None
        hist = None
        if token is None and any(p.requires_grad for p in plist):
            # GNN-input rows of the 2 max_hop nodes written before this window, as a function of the preprocessor
            P0 = state.host_count
            pos = torch.arange(P0 - 2 * mh, P0, device=dev)
            raw_hist = state.raw[:, pos.clamp(min=0) % state.C]
            hist = _apply_pre_grad(pre, raw_hist) * (pos >= 0).view(1, -1, 1).to(raw_hist.dtype)
        if not DenseGCM.did_warn and state.host_count + T > state.N:
            print("Overflow detected, wrapping around. Will not warn again")
            DenseGCM.did_warn = True
        y_seq = _apply_pre_grad(pre, x_seq)
        beliefs, token, state = temporal.sequence_grad(plan, state, y_seq, token, cap, hist=hist)
        if not plan.validated:
            plan.validated = True
            if not fused._validate_structure(plan, self, dev):
                raise RuntimeError("gcm: the GNN looked like a 2-layer DenseGraphConv stack but does not compute one")
        with torch.no_grad():
            xr = x_seq.detach()
            _cabi.check(_cabi.lib().gcm_state_log_write_seq(state.raw_ref(), xr.data_ptr(), xr.stride(0), xr.stride(1), T,
                                                            _cabi.stream_ptr(dev)), "gcm_state_log_write_seq")
        return beliefs, DenseHidden(state, token)

    def _rehome_pre(self, plan, state, capacity, pkey):
        """A preprocessed state in a log with more spare rows (before the first recorded step of a window)."""
        nodes_raw, adj, num_nodes = state.materialize()
        w = torch.zeros(0, device=state.device) if state.weights0 is None else state.materialize_weights()
        new, _ = DenseState.ingest(self.preprocessor(nodes_raw), adj, w, num_nodes, capacity)
        new.pure_key = state.pure_key
        new.host_count = None if state.host_count is None else min(state.host_count, state.N)
        new.status = state.status
        new.raw = torch.zeros(state.B, new.C, nodes_raw.shape[2], device=state.device, dtype=torch.float32)
        new.raw[:, : state.N] = nodes_raw
        new.pre_key = pkey
        return new

    def _would_overflow(self, state: DenseState) -> bool:
        # host-side mirror only (no device sync): graphs started empty overflow after N steps
        return state.host_count is not None and state.host_count >= state.N

    # ------------------------------------------------------------------ generic (unfused) path
    def _forward_generic(self, x, hidden):
        """Reference step semantics (gcm.py:213-321) with torch ops, for configurations outside the
        fused hot path.  Device-agnostic; selectors from gcm.edge_selectors still run their CUDA
        kernels on the dense adjacency."""
        if hidden is None:
            hidden = self.get_initial_hidden_state(x)
        nodes, adj, weights, num_nodes = hidden
        assert x.dtype == torch.float32
        assert nodes.dtype == torch.float
        assert weights.dtype == torch.float
        assert num_nodes.dtype == torch.long
        assert num_nodes.dim() == 1
        N = nodes.shape[1]
        B = x.shape[0]
        assert N == adj.shape[1] == adj.shape[2], "N must be equal for adj mat and node mat"
        b_idx = torch.arange(B, device=x.device)

        full = num_nodes + 1 > N
        if bool(full.any()):
            if not DenseGCM.did_warn:
                print("Overflow detected, wrapping around. Will not warn again")
                DenseGCM.did_warn = True
            nodes, adj, weights, num_nodes = self.wrap_overflow(nodes.clone(), adj.clone(), weights.clone(),
                                                                num_nodes.clone())
        nodes = nodes.clone()
        nodes[b_idx, num_nodes] = x
        dirty_nodes = nodes.clone()
        if self.edge_selectors:
            adj, weights = self.edge_selectors(dirty_nodes, adj.clone(), weights.clone(), num_nodes, B)
        if self.preprocessor:
            dirty_nodes = self.preprocessor(dirty_nodes)
        if self.aux_edge_selectors:
            sel_in = (self.positional_encoder(dirty_nodes, num_nodes) if self.positional_encoder
                      else dirty_nodes)
            adj, weights = self.aux_edge_selectors(sel_in, adj.clone(), weights.clone(), num_nodes, B)
        node_feats = self.gnn(dirty_nodes, adj, weights, B, N)
        mx = node_feats if self.pooled else node_feats[b_idx, num_nodes]
        assert torch.all(torch.isfinite(mx)), "Got NaN in returned memory, try using tanh activation"
        return mx, (nodes, adj, weights, num_nodes + 1)

    def wrap_overflow(self, nodes, adj, weights, num_nodes):
        """Drop node 0 of every full graph and shift the rest down one slot (reference gcm.py:323-355).
        Mutates and returns its arguments, like the reference."""
        N = nodes.shape[1]
        full = num_nodes + 1 > N
        idx = full.nonzero().flatten()
        if idx.numel():
            def shift2(m):
                out = torch.zeros_like(m)
                out[:, : N - 1, : N - 1] = m[:, 1:, 1:]
                return out

            sub = torch.zeros_like(nodes[idx])
            sub[:, : N - 1] = nodes[idx][:, 1:]
            nodes[idx] = sub
            adj[idx] = shift2(adj[idx])
            if weights.numel() != 0:
                weights[idx] = shift2(weights[idx])
            num_nodes[idx] = num_nodes[idx] - 1
        return nodes, adj, weights, num_nodes
