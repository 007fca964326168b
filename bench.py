#!/usr/bin/env python
"""bench.py — GCM env-steps/s on B200 (BASELINE.json metric), one process per GPU.

Default workload = BASELINE.json configs[1] ("cfg2"): DenseGCM graph_size 128, hidden 32,
TemporalBackedge([1,2,4]), 65536 graphs per GPU, forward pass in steady state (graphs full, the
oldest node is dropped every step).  A "step" is one DenseGCM step over the whole batch = ONE launch
of the fused step kernel.  Graphs are independent, so the batch shards across ranks with no data-path
collective (weak scaling: every rank owns --batch graphs).

JSON line (rank 0):
  value        env-steps/s, observations already resident in HBM (CUDA events, max over ranks).  For temporal
               chains the K timed steps are ONE `DenseGCM.forward_sequence(x[B,K,F], m_t)` call -- the loop of
               RayDenseGCM.forward (reference ray_gcm.py:200-202) enqueued from C; `per_call` holds the same K steps
               as K `DenseGCM.forward` calls (one Python round trip per step).
  e2e          same metric through the public API from HOST buffers: every step copies its [B,F]
               observation from pinned host memory and reads the [B,H] belief back
  roofline     algorithmic bytes per launch / kernel duration (CUDA events around
               back-to-back launches), against MEASURED_PEAKS.json
  cpu_baseline the reference step on the host cores (bounded sample)
  also         records of the other BASELINE configs in the same run (cfg4-cosine, cfg5, cfg5-train: one GPU only) and the
               fwd+bwd records: cfg2-bptt (BPTT over T=64 steps of the cfg2 chain) and cfg3 (DenseEdge
               N=256 H=128 T=64 with the NCCL gradient all-reduce), each with its own roofline / clocks / e2e

Other workloads (`--workload cfg1|cfg2-pre|cfg2-bptt|cfg2-pre-bptt|cfg3|cfg3-seq|cfg4-cosine|cfg4-euclid|cfg5|cfg5-train`) report the
remaining BASELINE configs with the same line format; `--impl reference` times the reference on CPU (the unmodified
reference package from baseline/_ref when it is installed, else the oracle port).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_PATHS = (os.path.join(ROOT, "graph-conv-memory_b200"), os.path.join(ROOT, "oracle"))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
STANDIN = os.path.join(ROOT, "tests", "golden", "standin")

# name -> (description, B, N, F, H, selector spec, mode)
WORKLOADS = {
    "cfg1": ("cfg1 README quickstart: DenseGCM N=128 F=8 H=32 TemporalBackedge([1]) B=16 rollout fwd",
             16, 128, 8, 32, [("temporal", (1,), "forward")], "rollout"),
    "cfg2": ("cfg2: DenseGCM N=128 F=32 H=32 TemporalBackedge([1,2,4]) rollout fwd",
             65536, 128, 32, 32, [("temporal", (1, 2, 4), "forward")], "rollout"),
    "cfg2-pre": ("cfg2 with RayDenseGCM's Linear preprocessor (SURVEY 8(f) rank 2): DenseGCM(preprocessor=Linear(32,32)) "
                 "N=128 H=32 TemporalBackedge([1,2,4]) rollout fwd", 65536, 128, 32, 32, [("temporal", (1, 2, 4), "forward")],
                 "rollout"),
    "cfg2-bptt": ("cfg2 chain, BPTT T=64 fwd+bwd on full graphs (truncated BPTT on a running rollout: m_t.detach() per "
                  "window), SGD step, gradient all-reduce", 65536, 128, 32, 32, [("temporal", (1, 2, 4), "forward")], "bptt"),
    "cfg2-pre-bptt": ("cfg2-bptt with RayDenseGCM's Linear(32,32) preprocessor, whose parameters train too (SURVEY 8(f) rank 2: "
                      "gradients reach it through every stored row, gcm.py:290-291)", 65536, 128, 32, 32,
                      [("temporal", (1, 2, 4), "forward")], "bptt"),
    "cfg3": ("cfg3: DenseGCM DenseEdge N=256 F=H=128 BPTT T=64 fwd+bwd (DenseEdge-only kernels, bf16 per-node cache, fp32 accumulate)",
             16384, 256, 128, 128, [("dense",)], "bptt"),
    "cfg3-seq": ("cfg3 through DenseGCM.forward_sequence (SURVEY 8(f) rank 1): the T=64 steps of a window in one call, "
                 "bf16 per-node cache, fwd+bwd", 16384, 256, 128, 128, [("dense",)], "bptt"),
    "cfg4-cosine": ("cfg4: DenseGCM CosineEdge(0.5) N=512 F=64 H=64 rollout fwd",
                    4096, 512, 64, 64, [("cosine", 0.5)], "rollout"),
    "cfg4-euclid": ("cfg4: DenseGCM EuclideanEdge(2.0) N=512 F=64 H=64 rollout fwd (cross-batch mean)",
                    4096, 512, 64, 64, [("euclidean", 2.0)], "rollout"),
    "cfg5": ("cfg5: SparseGCM TemporalEdge([1]) + SpatialRadiusEdge(0.25) N=4096 F=H=64 all-at-once",
             1024, 4096, 64, 64, None, "sparse"),
    "cfg5-train": ("cfg5 forward + backward (loss = mean of the outputs; gradients of the GraphConv weights and of x)",
                   1024, 4096, 64, 64, None, "sparse"),
}
BPTT_T = 64


def progress(msg):
    """stage markers on stderr (GCM_BENCH_VERBOSE=1): a killed run still shows how far it got"""
    if os.environ.get("GCM_BENCH_VERBOSE"):
        print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def _product_paths():
    for p in PKG_PATHS:
        if p not in sys.path:
            sys.path.insert(0, p)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the benchmark runs (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        # index: GPU index, a comma-separated list of indices (one sampler process for all ranks of the job: eight
        # nvidia-smi pollers queue on the driver's locks and show up as launch hiccups on every rank), or None = idle
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        if self.index is None:
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def window(self, t0=None, t1=None):
        """clocks summary of the samples taken between two perf_counter stamps (the sampler keeps running)"""
        if self.index is None:
            return None
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        rows = [r for t, r in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1 + 0.1)]
        if not rows:
            rows = [r for _, r in self.rows]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()


def make_selector(spec):
    from gcm.edge_selectors.dense import DenseEdge
    from gcm.edge_selectors.distance import CosineEdge, EuclideanEdge
    from gcm.edge_selectors.temporal import TemporalBackedge

    s = spec[0]
    if s[0] == "temporal":
        return TemporalBackedge(list(s[1]), direction=s[2])
    if s[0] == "dense":
        return DenseEdge()
    if s[0] == "cosine":
        return CosineEdge(s[1])
    return EuclideanEdge(s[1])


def build_dense(dev, N, F, H, spec, pre=False):
    import torch
    from gcm.gcm import DenseGCM
    from gcm.nn import DenseGraphConv

    class GNN(torch.nn.Module):          # the README's user GNN (README.md:52-62 of the reference)
        def __init__(self):
            super().__init__()
            self.gc0 = DenseGraphConv(F, H)
            self.gc1 = DenseGraphConv(H, H)
            self.act = torch.nn.Tanh()

        def forward(self, x, adj, weights, B, N):
            x = self.act(self.gc0(x, adj))
            return self.act(self.gc1(x, adj))

    torch.manual_seed(7)
    gnn = GNN().to(dev)
    pre = torch.nn.Linear(F, F).to(dev) if pre else None     # RayDenseGCM's Linear pre-projection (ray_gcm.py:118)
    return DenseGCM(gnn, preprocessor=pre, edge_selectors=make_selector(spec), graph_size=N)


def build_sparse(dev, N, F, H):
    import torch
    from gcm.nn import GraphConv
    from gcm.sparse_edge_selectors.spatial import SpatialRadiusEdge
    from gcm.sparse_edge_selectors.temporal import TemporalEdge
    from gcm.sparse_gcm import SparseGCM

    class GNN(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gc0 = GraphConv(F, H)
            self.gc1 = GraphConv(H, H)
            self.act = torch.nn.Tanh()

        def forward(self, x, edges, weights):
            x = self.act(self.gc0(x, edges, weights))
            return self.act(self.gc1(x, edges, weights))

    torch.manual_seed(7)
    return SparseGCM(GNN().to(dev), edge_selectors=TemporalEdge([1]),
                     aux_edge_selectors=SpatialRadiusEdge(slice(0, 2), 0.25), graph_size=N)


def synth_obs(gen, n, B, F, spec):
    """N(0,1) observations; clustered (K=16 centres, shared schedule) for distance edges."""
    import torch

    if spec and spec[0][0] in ("cosine", "euclidean"):
        centres = torch.randn(16, F, generator=gen)
        sched = torch.randint(0, 16, (n,), generator=gen)
        return (centres[sched].unsqueeze(1) + 0.05 * torch.randn(n, B, F, generator=gen)).contiguous()
    return torch.randn(n, B, F, generator=gen)


def algorithmic(workload, B, N, F, H, extra=None):
    """(bound, per-step algorithmic quantity, unit) — SURVEY.md §8(d)."""
    if workload in ("cfg1", "cfg2", "cfg2-pre"):
        hops = WORKLOADS[workload][5][0][1]
        r2 = len({0} | set(hops) | {a + b for a in hops for b in hops})
        per = r2 * F * 4 + F * 4 + F * 4 + N // 8 + H * 4 + 16
        return "hbm", per * B, "bytes"
    if workload in ("cfg2-bptt", "cfg2-pre-bptt"):
        # (cfg2-pre-bptt: the preprocessor's own rows are NOT counted, so its fraction is a lower bound)
        # per graph-step of a BPTT window: the forward's 1 440 B (SURVEY 8(d) k-hop figure) + the
        # backward's streams: dL/dbelief and the belief (act2'), h_t and its act1', the node row, dL/dx written
        hops = WORKLOADS[workload][5][0][1]
        r2 = len({0} | set(hops) | {a + b for a in hops for b in hops})
        fwd = r2 * F * 4 + F * 4 + F * 4 + N // 8 + H * 4 + 16
        bwd = 2 * H * 4 + H * 4 + F * 4 + F * 4
        return "hbm", (fwd + bwd) * B, "bytes"
    if workload.startswith("cfg4"):
        if workload == "cfg4-euclid":
            # the cross-batch mean distance: 2 B^2 N F useful flops per step, issued as THREE tf32 MMAs per product
            # (3xTF32: fp32-accurate) by k_euclid_tc -> the tensor-pipe work is 3x the useful figure
            return "tensor", 3 * 2.0 * B * B * N * F, "flop"
        per = N * F * 4 + 2 * F * 4 + N // 8 * 2 + H * 4 + 16
        return "hbm", per * B, "bytes"
    if workload in ("cfg3", "cfg3-seq"):
        # SURVEY.md 8(d), all-ones structure exploited: 66.8 KB per graph-step of the forward (bf16 per-node rows at
        # n = N); k_ones_fwd is that pass.  The backward is one pass per window, not 2x this per step.
        return "hbm", B * (N * F * 2 + 2 * F * 4 + N // 8 + H * 4 + 16), "bytes"
    if workload in ("cfg5", "cfg5-train"):
        n, E = extra
        per_layer = n * F * 4 + E * 8 + (n + 1) * 8 + n * H * 4
        return "hbm", 2 * per_layer * (3 if workload == "cfg5-train" else 1), "bytes"
    raise ValueError(workload)


# ------------------------------------------------------------------------------------------------
# CPU arms
# ------------------------------------------------------------------------------------------------
def reference_installed():
    return os.path.exists(os.path.join(REF_DIR, "gcm", "gcm.py"))


def cpu_reference_rate_real(workload, batch, steps, warm):
    """The UNMODIFIED reference package (pip-installed from /root/reference into baseline/_ref by
    __graft_entry__.build()) on the host cores.  Its torch_geometric dependency is not installable here; the stand-in
    under tests/golden/standin restates DenseGraphConv / Sequential per PyG's published definitions in plain torch.
    Must run in a process that has NOT imported this repository's own `gcm` package (same top-level name)."""
    assert "gcm" not in sys.modules, "the reference arm needs a process of its own"
    sys.path.insert(0, STANDIN)
    sys.path.insert(0, REF_DIR)
    import torch
    import torch_geometric
    from gcm.edge_selectors.dense import DenseEdge
    from gcm.edge_selectors.distance import CosineEdge, EuclideanEdge
    from gcm.edge_selectors.temporal import TemporalBackedge
    from gcm.gcm import DenseGCM

    desc, _, N, F, H, spec, mode = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)

    class GNN(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gc0 = torch_geometric.nn.DenseGraphConv(F, H)
            self.gc1 = torch_geometric.nn.DenseGraphConv(H, H)
            self.act = torch.nn.Tanh()

        def forward(self, x, adj, weights, B, N):
            x = self.act(self.gc0(x, adj))
            return self.act(self.gc1(x, adj))

    s = spec[0]
    sel = (TemporalBackedge(list(s[1]), direction=s[2]) if s[0] == "temporal" else DenseEdge() if s[0] == "dense"
           else CosineEdge(s[1]) if s[0] == "cosine" else EuclideanEdge(s[1]))
    torch.manual_seed(7)
    mod = DenseGCM(GNN(), edge_selectors=sel, graph_size=N)
    gen = torch.Generator().manual_seed(1002)
    DenseGCM.did_warn = True
    hidden = (torch.randn(batch, N, F, generator=gen), torch.zeros(batch, N, N), torch.zeros(0),
              torch.full((batch,), N, dtype=torch.long))
    obs = synth_obs(gen, 4, batch, F, spec)
    if mode == "bptt":
        opt = torch.optim.SGD(mod.parameters(), lr=1e-3)
        T = min(BPTT_T, 8)

        def window(hidden):
            opt.zero_grad(set_to_none=True)
            hidden = tuple(h.detach() for h in hidden)
            tot = 0
            for t in range(T):
                belief, hidden = mod(obs[t % 4], hidden)
                tot = tot + belief.mean()
            (tot / T).backward()
            opt.step()
            return hidden

        for _ in range(min(warm, 1)):
            hidden = window(hidden)
        t0 = time.perf_counter()
        for _ in range(steps):
            hidden = window(hidden)
        dt = time.perf_counter() - t0
        return batch * T * steps / dt, dt / steps, cores, f"B={batch} full graphs, windows of T={T} fwd+bwd+SGD"
    with torch.no_grad():
        for i in range(warm):
            _, hidden = mod(obs[i % 4], hidden)
        t0 = time.perf_counter()
        for i in range(steps):
            _, hidden = mod(obs[i % 4], hidden)
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, cores, f"B={batch} full graphs (wrap every step)"


def cpu_reference_rate(workload, batch, steps, warm):
    """The oracle port of the reference (oracle/gcm_oracle.py) on the host cores."""
    _product_paths()
    import torch

    import gcm_oracle as oracle

    desc, _, N, F, H, spec, mode = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    p = oracle.make_params(F, H)
    gen = torch.Generator().manual_seed(1002)
    if mode == "sparse":
        n_obs = min(N, 512)
        x = torch.randn(batch, n_obs, F, generator=gen)
        x[..., 0:2] = torch.cumsum(0.1 * torch.randn(batch, n_obs, 2, generator=gen), dim=1)
        taus = torch.full((batch,), n_obs)
        with torch.no_grad():
            t0 = time.perf_counter()
            for _ in range(max(1, steps)):
                oracle.sparse_gcm_forward(x, taus, None, [("temporal", (1,))], p, graph_size=N,
                                          aux_selectors=[("spatial_radius", slice(0, 2), 0.25)])
            dt = time.perf_counter() - t0
        return batch * n_obs * max(1, steps) / dt, dt / max(1, steps), cores, f"B={batch}, {n_obs} obs per graph"
    # dense: start full so every timed step includes the overflow shift, like the GPU arm's steady state
    hidden = (torch.randn(batch, N, F, generator=gen), torch.zeros(batch, N, N), torch.zeros(0),
              torch.full((batch,), N, dtype=torch.long))
    obs = synth_obs(gen, 4, batch, F, spec)
    ctx = torch.no_grad() if mode == "rollout" else torch.enable_grad()
    with ctx:
        for i in range(warm):
            _, hidden = oracle.dense_gcm_step(obs[i % 4], hidden, spec, p, graph_size=N)
        t0 = time.perf_counter()
        for i in range(steps):
            _, hidden = oracle.dense_gcm_step(obs[i % 4], hidden, spec, p, graph_size=N)
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, cores, f"B={batch} full graphs (wrap every step)"


def cpu_arm(workload, batch, steps, warm):
    """(rate, seconds per step, cores, sample text, kind): the real reference when it is installed and covers the
    workload's mode, else the oracle port."""
    mode = WORKLOADS[workload][6]
    if reference_installed() and mode != "sparse" and "gcm" not in sys.modules:
        rate, per, cores, what = cpu_reference_rate_real(workload, batch, steps, warm)
        return rate, per, cores, ("unmodified reference package (baseline/_ref) + torch_geometric stand-in "
                                  "(tests/golden/standin), " + what), "reference"
    rate, per, cores, what = cpu_reference_rate(workload, batch, steps, warm)
    return rate, per, cores, "oracle port of the reference, " + what, "port"


def run_reference(args, rank):
    if rank != 0:
        return
    desc = WORKLOADS[args.workload][0]
    mode = WORKLOADS[args.workload][6]
    rate, per, cores, what, kind = cpu_arm(args.workload, args.cpu_batch, args.steps, min(args.warmup, 4))
    sample = f"{what}; GPU arm runs {args.batch or WORKLOADS[args.workload][1]} graphs per GPU"
    print(json.dumps({
        "impl": "reference", "metric": "GCM env-steps/sec (fwd+bwd)" if mode == "bptt" else "GCM env-steps/sec (fwd)",
        "value": rate, "unit": "env-steps/s" if mode != "sparse" else "node-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "batch_per_step": args.cpu_batch, "timing": "host wall clock, CPU only"},
        "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def cpu_baseline_subprocess(workload, batch, steps, warm):
    """The CPU baseline of the default run, in a process of its own (the reference package shares its top-level name
    with this repository's)."""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload,
                              "--steps", str(steps), "--warmup", str(warm), "--cpu-batch", str(batch)],
                             capture_output=True, text=True, timeout=600)
        line = json.loads(out.stdout.strip().splitlines()[-1])
        cb = line["cpu_baseline"]
        cb["sample"] += f", {line['ms_per_step']:.1f} ms/step"
        cb["unit"] = line["unit"]
        return cb
    except Exception as e:  # noqa: BLE001 - the baseline is context, never fatal
        return {"value": None, "unit": "env-steps/s", "cores": os.cpu_count(), "kind": "port",
                "sample": f"unavailable: {type(e).__name__}: {e}"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local_rank):
    """Pin this rank to the CPUs of its GPU's NUMA node (before any pinned host buffer is allocated, so that first touch
    puts the pages there): with all ranks on node 0 the per-step H2D / D2H copies of 8 ranks share one memory controller
    and one PCIe root (round 1: e2e 1.76x at 8 GPUs).  Best effort: silently does nothing when sysfs / affinity say no."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(local_rank)
        pci = f"{bus.pci_domain_id:04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{pci}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return {"numa_node": node, "bound": False}
        os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "bound": True, "cpus": len(allowed)}
    except Exception:  # noqa: BLE001
        return None


class Ctx:
    def __init__(self, dev, rank, world, dist, sampler):
        self.dev, self.rank, self.world, self.dist, self.sampler = dev, rank, world, dist, sampler

    def barrier(self):
        import torch

        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        import torch

        t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def run_rollout(ctx, workload, B, K, W, args):
    import torch
    from gcm import _cabi

    desc, _, N, F, H, spec, mode = WORKLOADS[workload]
    dev, world = ctx.dev, ctx.world
    lib = _cabi.lib()
    gen = torch.Generator().manual_seed(1002 + ctx.rank)
    mod = build_dense(dev, N, F, H, spec, pre=workload == "cfg2-pre")
    n_obs = 16 if workload != "cfg4-euclid" else 4
    obs_host = synth_obs(gen, n_obs, B, F, spec).pin_memory()
    obs_dev = obs_host.to(dev)
    obs_steps = [obs_dev[i] for i in range(n_obs)]       # per-step views made once (2 us of host time per step)
    seq = spec[0][0] == "temporal"                       # temporal chains have the C rollout entry behind forward_sequence
    fill = max(0, N + 8 - W)                             # untimed: the graphs are FULL before the warm-up starts
    out = {}
    hidden = None
    with torch.no_grad():
        if seq:
            def xs(n, off=0):                            # [n, B, F]: step k of the call reads observation (off + k) % n_obs
                return torch.stack([obs_steps[(off + k) % n_obs] for k in range(n)], dim=0).contiguous()

            if fill:
                _, hidden = mod.forward_sequence(xs(fill), hidden, time_major=True)
            progress(f"{workload}: filled")
            x_warm, x_seq = xs(max(W, 2), fill), xs(K, fill + W)
            _, hidden = mod.forward_sequence(x_warm, hidden, time_major=True)
            _, hidden = mod.forward_sequence(x_seq, hidden, time_major=True)   # same call shape as the timed one
            ctx.barrier()
            t_lo = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # The K steps are ONE call, so its host-side latency (Python + the first launch, ~50 us; more with N ranks
            # sharing the host) would sit in front of ~450 us of device work and be charged to every step.  A short
            # spinning kernel queued BEFORE the start event keeps the stream busy while the host makes the call -- the
            # state any step but the first of a rollout is in -- so the events bracket exactly the K steps on the device.
            torch.cuda._sleep(int(2.0e6))                      # ~1 ms at 1.9 GHz, outside the timed region (200 us was not
                                                               # always enough with several ranks sharing the host: 22.9
                                                               # instead of 20.6 us per step at N = 2)
            e0.record()
            beliefs, hidden = mod.forward_sequence(x_seq, hidden, time_major=True)
            e1.record()
            ctx.barrier()
            total_ms = ctx.max_over_ranks(e0.elapsed_time(e1))
            # kernel-only duration: the same call queued behind a spinning blocker kernel
            n_l0 = lib.gcm_launch_count()
            torch.cuda._sleep(int(1.0e7))
            e0.record()
            beliefs, hidden = mod.forward_sequence(x_seq, hidden, time_major=True)
            e1.record()
            torch.cuda.synchronize()
            kern_ms = e0.elapsed_time(e1) / K
            launches = int(lib.gcm_launch_count() - n_l0)
            kernel_name = lib.gcm_last_kernel().decode()
            if workload == "cfg2-pre":
                kernel_name = "k_step_temporal_hc"       # (the call ends with the raw-log write)
            # the same K steps as K DenseGCM.forward calls (one Python round trip per step)
            for i in range(max(W, 3)):
                belief, hidden = mod(obs_steps[i % n_obs], hidden)
            ctx.barrier()
            h0 = time.perf_counter()
            e0.record()
            for i in range(K):
                belief, hidden = mod(obs_steps[i % n_obs], hidden)
            e1.record()
            host_us = (time.perf_counter() - h0) / K * 1e6
            ctx.barrier()
            pc_ms = ctx.max_over_ranks(e0.elapsed_time(e1))
            out["per_call"] = {"value": B * world * K / (pc_ms * 1e-3), "ms_per_step": pc_ms / K,
                               "host_us_per_call": host_us,
                               "what": "the same K steps as K DenseGCM.forward(x[B,F], m_t) calls (Python per step)"}
        else:
            for i in range(fill + W):
                belief, hidden = mod(obs_steps[i % n_obs], hidden)
            ctx.barrier()
            t_lo = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(K):
                belief, hidden = mod(obs_steps[i % n_obs], hidden)
            e1.record()
            ctx.barrier()
            total_ms = ctx.max_over_ranks(e0.elapsed_time(e1))
            n_l0 = lib.gcm_launch_count()
            torch.cuda._sleep(int(2.0e7))                      # ~10 ms at 1.9 GHz: covers K host-side launches
            h0 = time.perf_counter()
            e0.record()
            for i in range(K):
                belief, hidden = mod(obs_steps[i % n_obs], hidden)
            e1.record()
            host_us = (time.perf_counter() - h0) / K * 1e6
            torch.cuda.synchronize()
            kern_ms = e0.elapsed_time(e1) / K
            launches = int(lib.gcm_launch_count() - n_l0)
            kernel_name = lib.gcm_last_kernel().decode()
            if workload == "cfg4-euclid":
                kernel_name = "k_euclid_tc (cross-batch mean distance, 3xTF32 tcgen05) + " + kernel_name
            out["host_us_per_call"] = host_us
        # end to end through the public API from HOST buffers: every step copies its observation from pinned
        # host memory and reads its belief back; the copies run on their own streams (PCIe is full duplex)
        # and overlap the neighbouring steps' kernels, ordered by events
        R = 4
        main = torch.cuda.current_stream()
        h2d, d2h = torch.cuda.Stream(), torch.cuda.Stream()
        ring = [torch.empty(B, F, device=dev) for _ in range(R)]
        belief_host = [torch.empty(B, H).pin_memory() for _ in range(R)]
        ev_in = [torch.cuda.Event() for _ in range(R)]
        ev_done = [torch.cuda.Event() for _ in range(R)]

        def e2e_steps(n, hidden):
            for i in range(n):
                j = i % R
                with torch.cuda.stream(h2d):
                    if i >= R:
                        h2d.wait_event(ev_done[j])
                    ring[j].copy_(obs_host[i % n_obs], non_blocking=True)
                    ev_in[j].record(h2d)
                main.wait_event(ev_in[j])
                belief, hidden = mod(ring[j], hidden)
                ev_done[j].record(main)
                with torch.cuda.stream(d2h):
                    d2h.wait_event(ev_done[j])
                    belief_host[j].copy_(belief, non_blocking=True)
                    belief.record_stream(d2h)
            main.wait_stream(d2h)
            main.wait_stream(h2d)
            return hidden

        progress(f"{workload}: device-resident timing done, e2e next")
        hidden = e2e_steps(2 * R, hidden)
        e2e_runs = []
        for _ in range(3):                      # PCIe / host jitter: median of three timed passes of K steps
            ctx.barrier()
            e0.record()
            hidden = e2e_steps(K, hidden)
            e1.record()
            ctx.barrier()
            e2e_runs.append(ctx.max_over_ranks(e0.elapsed_time(e1)))
        e2e_ms = sorted(e2e_runs)[1]
        t_hi = time.perf_counter()
    hidden.claim().check_flags()
    out.update({
        "total_ms": total_ms, "kern_ms": kern_ms, "launches": launches, "kernel_name": kernel_name, "unit_per_step": B,
        "t_lo": t_lo, "t_hi": t_hi, "extra": None,
        "state": ("in-place node log + bit-packed adjacency; steady state: graphs full before the warm-up "
                  f"({fill} untimed fill steps + {W} warm-up steps), the oldest node is dropped every step"),
        "timed_call": (f"ONE DenseGCM.forward_sequence(x[{K},B,F], m_t, time_major=True) call: the C rollout entry walks "
                       f"the {K} steps in {launches} launch(es) of the step kernel; CUDA events around the call, the "
                       "stream kept busy by a spin kernel queued before the start event while the host makes the call"
                       if seq else f"{K} DenseGCM.forward calls"),
        "e2e": {"value": B * world * K / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": B * F * 4,
                "d2h_bytes_per_step": B * H * 4, "ms_per_step": e2e_ms / K,
                "ms_per_step_runs": [r / K for r in e2e_runs],
                "how": "per-step DenseGCM.forward; H2D / step kernel / D2H on three streams, 4-deep buffer ring, "
                       "pinned host memory"}})
    return out


def run_bptt(ctx, workload, B, K, W, args):
    import torch
    from gcm import _cabi
    from gcm import dist as gdist
    from gcm.state import DenseHidden, DenseState

    desc, _, N, F, H, spec, mode = WORKLOADS[workload]
    dev, world = ctx.dev, ctx.world
    lib = _cabi.lib()
    gen = torch.Generator().manual_seed(1002 + ctx.rank)
    T = BPTT_T
    temporal = spec[0][0] == "temporal"
    mod = build_dense(dev, N, F, H, spec, pre=workload == "cfg2-pre-bptt")
    mod.bptt_capacity = T
    if not temporal:
        mod.compute_dtype = torch.bfloat16 if args.cache == "bf16" else None
    opt = torch.optim.SGD(mod.parameters(), lr=1e-3)
    obs_host = ((1.0 if temporal else 0.5) * torch.randn(T, B, F, generator=gen)).pin_memory()
    obs_dev = obs_host.to(dev)
    obs_stage = torch.empty_like(obs_dev)
    seq = workload == "cfg3-seq"
    obs_bt = obs_dev.transpose(0, 1).contiguous() if seq else None         # [B, T, F] for the sequence entry
    loss_host = torch.empty(1).pin_memory()
    if temporal:
        # truncated BPTT on a running rollout: fill the graphs without autograd, then one window per optimiser step
        with torch.no_grad():
            _, carry = mod.forward_sequence(torch.randn(N + 8, B, F, device=dev), None, time_major=True)
        carry = [carry]

        def window(obs, obs_seq):
            opt.zero_grad(set_to_none=True)
            hidden = carry[0].detach()
            beliefs, hidden = mod.forward_sequence(obs, hidden, time_major=True)
            loss = beliefs.mean()
            loss.backward()
            gdist.allreduce_grads(mod.parameters(), average=True)
            opt.step()
            carry[0] = hidden
            return loss
    else:
        nn0 = torch.full((B,), N - T, dtype=torch.long, device=dev)            # pre-filled with N - T nodes
        nodes0 = 0.5 * torch.randn(B, N, F, device=dev)
        nodes0[:, N - T:] = 0
        adj0 = torch.zeros(B, N, N, device=dev)
        adj0[:, : N - T, : N - T] = 1                                            # DenseEdge history: all ones

        # the pre-filled memory every window starts from, ingested ONCE from the reference-layout
        # tensors; a window restarts from a copy of that handle (DenseHidden.clone(): the node log, not the [B,N,N]
        # float adjacency)
        with torch.no_grad():
            st0, flags0 = DenseState.ingest(nodes0, adj0, torch.zeros(0, device=dev), nn0, N + T)
        assert flags0 == 0
        h0 = DenseHidden(st0, None)
        start_from = {"handle": True}

        def window(obs, obs_seq):
            hidden = h0.clone() if start_from["handle"] else (nodes0, adj0, torch.zeros(0, device=dev), nn0)
            opt.zero_grad(set_to_none=True)
            if seq:
                beliefs, hidden = mod.forward_sequence(obs_seq, hidden)
                tot = beliefs.mean() * T
            else:
                tot = 0
                for t in range(T):
                    belief, hidden = mod(obs[t], hidden)
                    tot = tot + belief.mean()
            loss = tot / T
            loss.backward()
            gdist.allreduce_grads(mod.parameters(), average=True)               # the one NCCL collective
            opt.step()
            return loss

    for _ in range(min(W, 2)):
        window(obs_dev, obs_bt)
        progress(f"{workload}: warm-up window done")
    ctx.barrier()
    t_lo = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        window(obs_dev, obs_bt)
    e1.record()
    ctx.barrier()
    total_ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    n_l0 = lib.gcm_launch_count()
    window(obs_dev, obs_bt)
    torch.cuda.synchronize()
    progress(f"{workload}: timed windows done")
    launches = int(lib.gcm_launch_count() - n_l0) * K
    # end to end from HOST buffers: the window's observations come from pinned host memory, the loss is read back
    def e2e_window():
        obs_stage.copy_(obs_host, non_blocking=True)
        ob = obs_stage.transpose(0, 1).contiguous() if seq else None
        loss = window(obs_stage, ob)
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(loss_host[0])

    e2e_window()
    ctx.barrier()
    Ke = max(1, min(K, 3))
    e0.record()
    for _ in range(Ke):
        e2e_window()
    e1.record()
    ctx.barrier()
    e2e_ms = ctx.max_over_ranks(e0.elapsed_time(e1)) / Ke
    t_hi = time.perf_counter()
    window_ms = total_ms / K
    extra = {"window_ms": window_ms, "window_T": T}
    if temporal:
        # whole-window roofline: (forward + backward algorithmic bytes of the T steps) / window time; dominant kernel =
        # the forward step kernel, timed alone on a no-grad sequence call
        with torch.no_grad():
            hid = carry[0].detach()
            _, hid = mod.forward_sequence(obs_dev, hid, time_major=True)
            torch.cuda._sleep(int(1.0e7))
            e0.record()
            _, hid = mod.forward_sequence(obs_dev, hid, time_major=True)
            e1.record()
            torch.cuda.synchronize()
            carry[0] = hid
        step_kernel_ms = e0.elapsed_time(e1) / T
        kern_ms = window_ms / T
        kernel_name = ("whole window: T x k_step_temporal_hc forward + the window-level backward "
                       "(gcm.temporal_bwd), per graph-step")
        extra.update({"fwd_step_kernel_ms": step_kernel_ms, "fwd_kernel_share_of_window": step_kernel_ms * T / window_ms})
    else:
        from gcm import ones as _ones
        with torch.no_grad():
            _, hid = mod(obs_dev[0], (nodes0, adj0, torch.zeros(0, device=dev), nn0))
            for t in range(1, T):
                _, hid = mod(obs_dev[t], hid)
        kern_ms = _ones.time_fwd_kernel(mod.fused_plan(), hid.claim())
        # the same window started from the reference-layout TUPLE (ingest of the [B,N,N] float adjacency inside the window)
        start_from["handle"] = False
        window(obs_dev, obs_bt)
        torch.cuda.synchronize()
        e0.record()
        window(obs_dev, obs_bt)
        e1.record()
        torch.cuda.synchronize()
        extra["window_ms_from_tuple"] = e0.elapsed_time(e1)
        start_from["handle"] = True
        _, algo, _ = algorithmic(workload, B, N, F, H)
        pk, _ = peaks()
        extra.update({"fwd_kernel_share_of_window": kern_ms * T / window_ms,
                      "window_fwd_bytes_frac": algo * T / (window_ms * 1e-3) / 1e9 / pk["hbm_gbs"]})
        kernel_name = ("k_ones_fwd (1 launch per step; the backward is ONE k_ones_window_bwd per window, "
                       "see DESIGN.md section 3)")
        if seq:
            extra["note"] = ("the sequence entry replaces the 64 k_ones_fwd launches by ONE k_ones_window_fwd (cache rows "
                             "read once, MUFU-bound); kernel_ms / frac above are those of the per-step kernel for reference")
    return {"total_ms": total_ms, "kern_ms": kern_ms, "launches": launches, "kernel_name": kernel_name,
            "unit_per_step": B * T, "t_lo": t_lo, "t_hi": t_hi, "extra": extra, "state": mode,
            "timed_call": f"{K} windows of T={T}: forward, backward, gradient all-reduce, SGD step",
            "e2e": {"value": B * T * world / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": T * B * F * 4,
                    "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms,
                    "how": "per window: the [T,B,F] observations copied from pinned host memory, the loss read back"}}


def run_sparse(ctx, workload, B, K, W, args):
    import torch

    desc, _, N, F, H, spec, mode = WORKLOADS[workload]
    dev = ctx.dev
    gen = torch.Generator().manual_seed(1002 + ctx.rank)
    mod = build_sparse(dev, N, F, H)
    x = torch.randn(B, N, F, generator=gen)
    x[..., 0:2] = torch.cumsum(0.1 * torch.randn(B, N, 2, generator=gen), dim=1)
    x_host = x.pin_memory()
    x_dev = x_host.to(dev)
    taus = torch.full((B,), N, dtype=torch.long, device=dev)
    train = workload == "cfg5-train"
    if train:
        x_dev.requires_grad_(True)

    def call():
        if train:
            for p_ in mod.parameters():
                p_.grad = None
            x_dev.grad = None
            out, hid = mod(x_dev, taus, None)
            out.mean().backward()
            return out, hid
        with torch.no_grad():
            return mod(x_dev, taus, None)

    for _ in range(W):
        out, hid = call()
    ctx.barrier()
    t_lo = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        out, hid = call()
    e1.record()
    ctx.barrier()
    total_ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    t_hi = time.perf_counter()
    E = int(hid[1]._nnz())
    kernel_name = "whole call: edge search + expansion, node write, 2 x (k_csr_gather + k_graphconv_fwd_tc)" + (
        " + backward (k_csr_transpose_smem, k_act_bwd_v4, k_linear_tc32, k_outer_tc32, k_graphconv_bwd_gather,"
        " k_sparse_write_flatten_bwd)" if train else "")
    lib_launches = 8 + (15 if train else 0)       # this library's kernels per call (torch's scans / fills not counted)
    return {"total_ms": total_ms, "kern_ms": total_ms / K, "launches": K * lib_launches, "kernel_name": kernel_name,
            "unit_per_step": B * N, "t_lo": t_lo, "t_hi": t_hi, "extra": (B * N, E), "state": mode,
            "timed_call": f"{K} SparseGCM.forward calls (all {N} observations of every graph at once)", "e2e": None}


def run_workload(ctx, workload, K, W, args, batch=None):
    """One workload -> the JSON line's dictionary (without cpu_baseline)."""
    import torch

    desc, B0, N, F, H, spec, mode = WORKLOADS[workload]
    B = batch or B0
    progress(f"{workload}: start (B={B}, K={K}, W={W})")
    r = {"rollout": run_rollout, "bptt": run_bptt, "sparse": run_sparse}[mode](ctx, workload, B, K, W, args)
    progress(f"{workload}: timed region done, {r['total_ms'] / K:.4f} ms per step")
    total_ms, kern_ms, extra = r["total_ms"], r["kern_ms"], r["extra"]
    value = r["unit_per_step"] * ctx.world * K / (total_ms * 1e-3)
    clocks = ctx.sampler.window(r["t_lo"], r["t_hi"])
    torch.cuda.empty_cache()
    if ctx.rank != 0:
        return None
    pk, pk_kind = peaks()
    bound, algo, algo_unit = algorithmic(workload, B, N, F, H, extra if isinstance(extra, tuple) else None)
    if algo_unit == "bytes":
        achieved, peak, unit = algo / (kern_ms * 1e-3) / 1e9, pk["hbm_gbs"], "GB/s"
    else:
        # tf32 dense peak = half of the measured bf16 peak (MEASURED_PEAKS.json has no tf32 entry of its own)
        achieved, peak, unit = algo / (kern_ms * 1e-3) / 1e12, pk["bf16_tflops"] / 2.0, "TFLOP/s"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(workload)
    fwdbwd = mode == "bptt" or workload == "cfg5-train"
    line = {
        "metric": "GCM env-steps/sec (fwd+bwd)" if fwdbwd else "GCM env-steps/sec (fwd)",
        "value": value, "unit": "env-steps/s" if mode != "sparse" else "node-steps/s", "n_gpus": ctx.world,
        "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": "bf16 cache / f32 accumulate" if (mode == "bptt" and spec[0][0] == "dense" and args.cache == "bf16") else "f32",
        "data": "synthetic",
        "config": {"workload": desc, "batch_per_gpu": B, "graph_size": N, "obs_size": F, "hidden": H,
                   "state": r["state"], "timed_call": r["timed_call"],
                   "l2": f"per-GPU state {B * N * F * 4 / 1e6:.0f} MB vs 126 MB L2; fresh observations every step",
                   "parallelism": f"batch-sharded x{ctx.world}, no data-path collective"
                   + (" (one NCCL all-reduce of the weight gradients per window)" if mode == "bptt" else "")},
        "clocks": clocks, "gpu_launches": r["launches"],
        "roofline": {"bound": "hbm" if bound == "hbm" else "tensor", "achieved": achieved, "peak": peak, "unit": unit,
                     "frac": achieved / peak, "traffic": traffic, "peak_source": pk_kind if unit == "GB/s"
                     else pk_kind + " bf16 dense / 2 = tf32 dense; achieved counts the 3 tf32 MMAs of every 3xTF32 product",
                     "kernel": r["kernel_name"], "kernel_ms": kern_ms, "algorithmic_per_launch": algo,
                     "note": "achieved = algorithmic bytes of ONE step / kernel_ms, kernel_ms = duration of the timed launch(es) / "
                             "steps (a multi-step launch walks all K steps: per-launch figures are these times K); traffic "
                             "is per step as well"},
    }
    for k in ("per_call", "host_us_per_call"):
        if k in r:
            (line if k == "per_call" else line["roofline"])[k] = r[k]
    if r["e2e"] is not None:
        line["e2e"] = r["e2e"]
    else:
        line["e2e"] = {"value": value, "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                       "note": "device-resident only for this auxiliary workload"}
    if isinstance(extra, dict):
        line["roofline"].update(extra)
    elif extra is not None:
        line["config"]["flat_nodes"], line["config"]["edges"] = extra
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None, help="graphs per GPU (default: the workload's)")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-batch", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="cfg2 only: skip the fwd+bwd records (cfg2-bptt, cfg3)")
    ap.add_argument("--cache", default="bf16", choices=["bf16", "f32"],
                    help="cfg3: element type of the per-node cache (bf16 = the config's stated precision)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    desc, B0, N, F, H, spec, mode = WORKLOADS[args.workload]
    if args.warmup is None:
        args.warmup = 8 if mode == "rollout" else 3
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        if args.workload == "cfg5":
            args.cpu_batch = min(args.cpu_batch, 16)
        elif args.workload != "cfg2":
            args.cpu_batch = min(args.cpu_batch, 64)
        run_reference(args, rank)
        return

    _product_paths()
    import torch

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    numa = bind_to_gpu_numa(local) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    # rank 0 samples the clocks of every GPU of the job (local ranks 0 .. world-1 of this node)
    sampler = ClockSampler(",".join(str(i) for i in range(world)) if rank == 0 else None)
    sampler.start()
    ctx = Ctx(dev, rank, world, dist, sampler)
    line = run_workload(ctx, args.workload, args.steps, args.warmup, args, args.batch)
    also = []
    if args.workload == "cfg2" and not args.no_also and args.batch is None:
        # the fwd+bwd half of BASELINE.json's metric, measured in the same run
        wls = [("cfg2-bptt", 4), ("cfg3", 3), ("cfg3-seq", 3)]
        if world == 1:
            # the distance-selector and sparse configs of BASELINE.json in the same driver-run record (one GPU: they shard
            # like cfg2, no collective on their forward path)
            wls += [("cfg4-cosine", 20), ("cfg5", 3), ("cfg5-train", 3)]
        for wl, k in wls:
            try:
                rec = run_workload(ctx, wl, k, 3, args)
            except Exception as exc:      # an auxiliary record must not take the headline line down with it
                if wl in ("cfg2-bptt", "cfg3", "cfg3-seq"):
                    raise
                progress(f"{wl}: skipped ({type(exc).__name__}: {exc})")
                rec = None
            if rec is not None:
                rec = {key: rec[key] for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                                                 "dtype", "config", "clocks", "gpu_launches", "roofline", "e2e")}
                rec["workload"] = wl
                also.append(rec)
    sampler.stop()
    if rank == 0:
        if also:
            line["also"] = also
        if numa is not None:
            line["config"]["host_binding"] = numa
        if world == 1 and not args.no_cpu_baseline:
            cb = {"cfg2": args.cpu_batch, "cfg5": 8}.get(args.workload, 64)
            progress("cpu baseline: start")
            line["cpu_baseline"] = cpu_baseline_subprocess(args.workload, cb, 6 if mode != "sparse" else 1, 2)
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
