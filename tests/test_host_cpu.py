"""CPU tier: host logic of the product package (no compute calls): C-ABI exports, plan matching,
selector descriptors, index helpers, loud failure without CUDA."""
import ctypes
import os
import re

import pytest
import torch

import gcm_oracle as oracle
from helpers import make_dense_gnn, make_selector, make_sparse_gnn, make_sparse_selector

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from gcm import _cabi

    header = open(os.path.join(ROOT, "include", "gcm_b200.h")).read()
    declared = set(re.findall(r"\b(gcm_[a-z_0-9]+)\s*\(", header))
    assert declared, "no prototypes parsed"
    lib = ctypes.CDLL(_cabi.lib_path())
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/gcm_b200.h but not exported"
    assert declared == set(_cabi.EXPORTED_SYMBOLS)
    assert _cabi.lib().gcm_version() == _cabi.GCM_ABI_VERSION


def test_abi_struct_layout_matches_header(tmp_path):
    """sizeof/offsetof as gcc sees include/gcm_b200.h == the ctypes mirrors in gcm/_cabi.py."""
    import subprocess

    from gcm import _cabi

    src = tmp_path / "abi.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "gcm_b200.h"\n'
        "int main(void){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(gcm_dense_state), sizeof(gcm_selector),"
        " sizeof(gcm_gnn), sizeof(gcm_gnn_grads), offsetof(gcm_selector, max_distance), offsetof(gcm_selector, dist_param),"
        " offsetof(gcm_gnn, F), offsetof(gcm_dense_state, B), sizeof(gcm_rollout), offsetof(gcm_rollout, sels),"
        " offsetof(gcm_rollout, hcache), offsetof(gcm_rollout, status), offsetof(gcm_rollout, launches), offsetof(gcm_rollout, xrec_from));return 0;}\n")
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [ctypes.sizeof(_cabi.DenseStateC), ctypes.sizeof(_cabi.SelectorC), ctypes.sizeof(_cabi.GnnC),
            ctypes.sizeof(_cabi.GnnGradsC), _cabi.SelectorC.max_distance.offset, _cabi.SelectorC.dist_param.offset,
            _cabi.GnnC.F.offset, _cabi.DenseStateC.B.offset, ctypes.sizeof(_cabi.RolloutC), _cabi.RolloutC.sels.offset,
            _cabi.RolloutC.hcache.offset, _cabi.RolloutC.status.offset, _cabi.RolloutC.launches.offset,
            _cabi.RolloutC.xrec_from.offset]
    assert got == want


def test_plan_matching_accepts_the_reference_gnn_shapes():
    from gcm import fused
    from gcm.gcm import DenseGCM
    from gcm.nn import DenseGraphConv, Sequential

    p = oracle.make_params(8, 32)
    for style in ("readme", "sequential"):
        gnn, _ = make_dense_gnn(8, 32, p, ("tanh", "tanh"), style)
        plan = fused.build_plan(DenseGCM(gnn, edge_selectors=make_selector([("temporal", (1, 2, 4), "forward")])))
        assert plan is not None and (plan.gnn.F, plan.gnn.H1, plan.gnn.H2) == (8, 32, 32)
        assert (plan.gnn.act1, plan.gnn.act2) == ("tanh", "tanh")
        assert plan.temporal_key is not None

    class Readme(torch.nn.Module):                      # README.md:52-62: ONE activation module, used twice
        def __init__(self):
            super().__init__()
            self.gc0 = DenseGraphConv(8, 32)
            self.gc1 = DenseGraphConv(32, 32)
            self.act = torch.nn.Tanh()

        def forward(self, x, adj, weights, B, N):
            return self.act(self.gc1(self.act(self.gc0(x, adj)), adj))

    plan = fused.build_plan(DenseGCM(Readme(), edge_selectors=make_selector([("dense",)])))
    assert plan is not None and (plan.gnn.act1, plan.gnn.act2) == ("tanh", "tanh") and plan.temporal_key is None
    # PyG 1.x bias placement is accepted too
    g = Sequential("x, adj, weights, B, N", [(DenseGraphConv(4, 4, bias_on="root"), "x, adj -> x"), torch.nn.ReLU()])
    assert fused.match_gnn(g) is None                   # a single layer is not the 2-layer hot path
    g = Sequential("x, adj, weights, B, N", [(DenseGraphConv(4, 6, bias_on="root"), "x, adj -> x"), torch.nn.ReLU(),
                                              (DenseGraphConv(6, 5, bias_on="root"), "x, adj -> x")])
    gp = fused.match_gnn(g)
    assert gp is not None and (gp.act1, gp.act2) == ("relu", "none") and gp.H2 == 5


def test_plan_matching_rejects_what_is_outside_the_hot_path():
    from gcm import fused
    from gcm.gcm import DenseGCM, PositionalEncoding
    from gcm.nn import DenseGraphConv, Sequential

    p = oracle.make_params(4, 4)
    gnn, _ = make_dense_gnn(4, 4, p, ("tanh", "tanh"))
    # a per-row preprocessor (what RayDenseGCM installs) is fusable with selectors that do not look at node contents ...
    pre_plan = fused.build_plan(DenseGCM(gnn, preprocessor=torch.nn.Linear(6, 4), edge_selectors=make_selector([("dense",)])))
    assert pre_plan is not None and pre_plan.pre
    assert fused.build_plan(DenseGCM(gnn, preprocessor=torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.Tanh()))).pre
    # ... but not with a distance selector (it reads the RAW rows, gcm.py:284-287) or a map that is not per-row
    assert fused.build_plan(DenseGCM(gnn, preprocessor=torch.nn.Linear(6, 4), edge_selectors=make_selector([("cosine", 0.5)]))) is None
    assert fused.build_plan(DenseGCM(gnn, preprocessor=torch.nn.LayerNorm(4))) is None
    assert fused.build_plan(DenseGCM(gnn, pooled=True)) is None
    assert fused.build_plan(DenseGCM(gnn, aux_edge_selectors=make_selector([("dense",)]))) is None
    assert fused.build_plan(DenseGCM(gnn, positional_encoder=PositionalEncoding())) is None
    assert fused.build_plan(DenseGCM(torch.nn.Sequential(torch.nn.Linear(4, 4)))) is None
    lam = Sequential("x, adj, weights, B, N", [(DenseGraphConv(4, 4), "x, adj -> x"), (lambda x: x, "x -> x"),
                                                (DenseGraphConv(4, 4), "x, adj -> x")])
    assert fused.build_plan(DenseGCM(lam)) is None

    class Custom(torch.nn.Module):
        def forward(self, nodes, adj, w, n, B):
            return adj, w

    assert fused.build_plan(DenseGCM(gnn, edge_selectors=Custom())) is None


def test_selector_descriptors():
    from gcm import _cabi
    from gcm.edge_selectors.distance import CosineEdge, EuclideanEdge, SpatialEdge
    from gcm.edge_selectors.temporal import TemporalBackedge

    s = TemporalBackedge([1, 2, 4], direction="both").fused_spec().to_c(32)
    assert (s.kind, s.direction, s.n_hops, list(s.hops)[:3]) == (_cabi.SEL_TEMPORAL, 2, 3, [1, 2, 4])
    s = SpatialEdge(0.5, slice(0, 2)).fused_spec().to_c(8)
    assert (s.kind, s.a_start, s.b_start, s.slice_len, s.a_step) == (_cabi.SEL_SPATIAL, 0, 0, 2, 1)
    s = SpatialEdge(0.5, slice(1, 7, 2), slice(0, 3)).fused_spec().to_c(8)
    assert (s.a_start, s.a_step, s.b_start, s.slice_len) == (1, 2, 0, 3)
    with pytest.raises(RuntimeError):
        SpatialEdge(0.5, slice(0, 2), slice(0, 3)).fused_spec().to_c(8)
    e = EuclideanEdge(1.0, learned=True)
    assert e.max_distance == 1.0 and e.fused_spec().dist_param is e.dist_param
    assert CosineEdge(0.3).fused_spec().max_distance == pytest.approx(0.3)
    with pytest.raises(NotImplementedError):
        TemporalBackedge(learned=True)
    chain = make_selector([("temporal", (1,), "forward"), ("dense",)])
    from gcm import fused
    assert [q.kind for q in fused.match_selectors(chain)] == [_cabi.SEL_TEMPORAL, _cabi.SEL_DENSE]


def test_sparse_plan_matching():
    from gcm.sparse_gcm import SparseGCM

    p = oracle.make_params(6, 5)
    for style in ("readme", "sequential"):
        gnn, _ = make_sparse_gnn(6, 5, p, ("tanh", "tanh"), style)
        m = SparseGCM(gnn, edge_selectors=make_sparse_selector([("temporal", (1, 3))]),
                      aux_edge_selectors=make_sparse_selector([("spatial_radius", slice(0, 2), 0.25)]))
        (c1, c2, a1, a2), hops, radius = m.fused_plan()
        assert (a1, a2, hops, radius[1]) == ("tanh", "tanh", (1, 3), 0.25)
        assert SparseGCM(gnn, max_hops=1).fused_plan() is not None      # masked aggregation (forward only)
        assert SparseGCM(gnn, max_hops=2).fused_plan() is not None
        assert SparseGCM(gnn, preprocessor=torch.nn.Linear(6, 6)).fused_plan() is None


def test_index_helpers_match_the_reference_definitions():
    """util.py:176-240, 426-452 of the reference are python loops over the batch; ours are closed-form."""
    from gcm import util

    T = torch.tensor([2, 0, 5, 1])
    taus = torch.tensor([3, 4, 0, 2])
    B = 4
    b, k = util.get_new_node_idxs(T, taus, B)
    assert b.tolist() == [0, 0, 0, 1, 1, 1, 1, 3, 3] and k.tolist() == [2, 3, 4, 0, 1, 2, 3, 1, 2]
    b, k = util.get_nonpadded_idxs(T, taus, B)
    assert k.tolist() == [0, 1, 2, 0, 1, 2, 3, 0, 1]
    b, k = util.get_valid_node_idxs(T, taus, B)
    assert b.tolist() == [0] * 5 + [1] * 4 + [2] * 5 + [3] * 3
    starts, ends = util.get_batch_offsets(T + taus)
    assert starts.tolist() == [0, 5, 9, 14] and ends.tolist() == [5, 9, 14, 17]
    nodes = torch.arange(B * 8 * 2, dtype=torch.float).reshape(B, 8, 2)
    flat, out_idx = util.flatten_nodes(nodes, T, taus, B)
    assert flat.shape == (17, 2) and out_idx.tolist() == [2, 3, 4, 5, 6, 7, 8, 15, 16]
    e = util.get_causal_edges_one_batch(torch.tensor(2), torch.tensor(2))
    assert e.tolist() == [[2, 2, 3, 3, 3], [0, 1, 0, 1, 2]]
    # flatten / unflatten and pack / unpack round trips (reference tests/test_sparse_gcm.py:17-304)
    idx = torch.tensor([[0, 0, 1, 3], [1, 4, 2, 2], [0, 2, 0, 1]])
    adj = torch.sparse_coo_tensor(idx, torch.ones(4), size=(B, 8, 8))
    fe, fw, fb = util.flatten_adj(adj, T, taus, B)
    assert fe.tolist() == [[1, 4, 7, 16], [0, 2, 5, 15]]
    back = util.unflatten_adj(fe, fw, fb, T, taus, B, 8).coalesce()
    assert torch.equal(back.indices(), adj.coalesce().indices())
    packed = util.pack_hidden((nodes, adj, T), B, max_edges=5)
    assert packed[1].shape == (B, 2, 5) and packed[1][0, :, :2].tolist() == [[1, 4], [0, 2]] and packed[1][2, 0, 0] == -1
    n2, adj2, T2 = util.unpack_hidden(packed, B)
    assert torch.equal(adj2.coalesce().indices(), adj.coalesce().indices())


def test_fused_paths_fail_loudly_without_cuda():
    from gcm import _cabi
    from gcm.gcm import DenseGCM
    from gcm.sparse_gcm import SparseGCM

    p = oracle.make_params(4, 8)
    gnn, _ = make_dense_gnn(4, 8, p, ("tanh", "tanh"))
    mod = DenseGCM(gnn, edge_selectors=make_selector([("temporal", (1,), "forward")]), graph_size=6)
    with pytest.raises(_cabi.GcmLibraryError, match="no CPU fallback"):
        mod(torch.randn(2, 4), None)
    with pytest.raises(_cabi.GcmLibraryError, match="no CPU fallback"):
        make_selector([("dense",)])(torch.zeros(2, 6, 4), torch.zeros(2, 6, 6), torch.zeros(0), torch.zeros(2, dtype=torch.long), 2)
    sg, _ = make_sparse_gnn(4, 8, p, ("tanh", "tanh"))
    sm = SparseGCM(sg, edge_selectors=make_sparse_selector([("temporal", (1,))]), graph_size=6)
    with pytest.raises(_cabi.GcmLibraryError, match="no CPU fallback"):
        sm(torch.randn(2, 3, 4), torch.tensor([3, 2]), None)


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    from gcm import _cabi

    monkeypatch.setenv("GCM_B200_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_cabi, "_lib", None)
    with pytest.raises(_cabi.GcmLibraryError, match="no CPU or eager fallback"):
        _cabi.lib()


def test_shard_bounds_partition_the_batch():
    from gcm import dist as gdist

    for batch, world in ((65536, 8), (10, 4), (3, 8), (16384, 3)):
        spans = [gdist.shard_bounds(batch, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == batch
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) needs no GPU: one JSON line with the
    contract's keys, timed on the oracle port."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for wl in ("cfg2", "cfg3"):
        out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", wl,
                              "--steps", "1", "--warmup", "1", "--cpu-batch", "4"], capture_output=True, text=True,
                             timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        line = json.loads(out.stdout.strip().splitlines()[-1])
        assert line["impl"] == "reference" and line["value"] > 0 and line["gpu_launches"] == 0
        for key in ("metric", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
                    "config", "cpu_baseline", "e2e"):
            assert key in line, key
        assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["e2e"]["h2d_bytes_per_step"] == 0


def test_bench_workload_table_and_roofline_accounting_are_consistent():
    """Every bench workload has its SURVEY 8(d) accounting (no GPU needed): the bound, a positive per-step quantity
    and a unit, for the shapes in the table."""
    import os
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench

    for name, (desc, B, N, F, H, spec, mode) in bench.WORKLOADS.items():
        extra = (B * N, 20 * B * N) if mode == "sparse" else None
        bound, qty, unit = bench.algorithmic(name, B, N, F, H, extra)
        assert bound in ("hbm", "tensor") and qty > 0 and unit in ("bytes", "flop"), name
        assert mode in ("rollout", "bptt", "sparse") and isinstance(desc, str)
    # cfg2's figure is SURVEY's 1 440 B per graph-step
    assert bench.algorithmic("cfg2", 65536, 128, 32, 32)[1] == 65536 * 1440


def test_ones_window_buffers_grow_without_losing_recorded_steps():
    """gcm.ones._Window: the per-step buffers of a BPTT window grow geometrically up to the log's capacity and keep what
    was recorded (host logic only; runs on CPU tensors)."""
    from types import SimpleNamespace

    from gcm import ones

    st = SimpleNamespace(B=3, device=torch.device("cpu"))
    g = SimpleNamespace(F=4, H1=8, H2=8)
    win = ones._Window(st, g)
    win.ensure_fwd(0, cap=200)
    assert win.K == 64 and win.S.shape == (64, 3, 4) and win.E.shape == (64, 3, 8)
    win.G[5].fill_(7.0)
    win.ensure_fwd(64, cap=200)
    assert win.K == 128 and float(win.G[5].min()) == 7.0
    win.ensure_fwd(150, cap=200)
    assert win.K == 200
    win.need_dx = True
    win.ensure_bwd()
    assert win.do.shape == (200, 3, 8) and win.dcs.shape == (200, 3, 8)


def test_gnn_key_is_matches_current_key():
    """GnnPlan.key_is (the per-call rollout path's weights check) agrees with current_key: same parameters -> True; an
    in-place update, a replaced parameter, another device, a bias that appears -> False."""
    import torch
    from gcm import fused
    from gcm.gcm import DenseGCM
    from helpers import make_dense_gnn, make_selector
    import gcm_oracle as oracle

    gnn, convs = make_dense_gnn(8, 16, oracle.make_params(8, 16), ("tanh", "tanh"))
    mod = DenseGCM(gnn, edge_selectors=make_selector([("temporal", (1,), "forward")]), graph_size=8)
    plan = mod.fused_plan()
    assert plan is not None
    g = plan.gnn
    dev = torch.device("cpu")
    key = g.current_key(dev)
    assert g.key_is(key, dev) and not g.key_is(key, torch.device("cuda:0")) and not g.key_is(None, dev)
    assert not g.key_is(key[:-3] + key[-1:], dev)
    with torch.no_grad():
        convs[0].lin_rel.weight.add_(1.0)                    # in-place update: version counter
    assert not g.key_is(key, dev) and g.key_is(g.current_key(dev), dev)
    key = g.current_key(dev)
    convs[1].lin_root.weight = torch.nn.Parameter(convs[1].lin_root.weight.detach().clone())     # replaced parameter
    assert not g.key_is(key, dev) and g.key_is(g.current_key(dev), dev)
