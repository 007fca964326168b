#!/usr/bin/env python
"""bench.py — GCM env-steps/s on B200 (BASELINE.json metric), one process per GPU.

Default workload = BASELINE.json configs[1] ("cfg2"): DenseGCM graph_size 128, hidden 32,
TemporalBackedge([1,2,4]), 65536 graphs per GPU, forward rollout in steady state (graphs full, the
oldest node is dropped every step).  A "step" is one DenseGCM.forward over the whole batch = ONE launch
of the fused step kernel.  Graphs are independent, so the batch shards across ranks with no data-path
collective (weak scaling: every rank owns --batch graphs).

JSON line (rank 0):
  value        env-steps/s, observations already resident in HBM (CUDA events, max over ranks)
  e2e          same metric through the public API from HOST buffers: every step copies its [B,F]
               observation from pinned host memory and reads the [B,H] belief back
  roofline     algorithmic bytes per launch / kernel duration (CUDA events around
               back-to-back launches), against MEASURED_PEAKS.json
  cpu_baseline the oracle port of the reference step on the host cores (bounded sample)

Other workloads (`--workload cfg1|cfg2-pre|cfg3|cfg3-seq|cfg4-cosine|cfg4-euclid|cfg5|cfg5-train`) report the remaining BASELINE
configs with the same line format; `--impl reference` times the oracle port of the reference on CPU.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "graph-conv-memory_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

# name -> (description, B, N, F, H, selector spec, mode)
WORKLOADS = {
    "cfg1": ("cfg1 README quickstart: DenseGCM N=128 F=8 H=32 TemporalBackedge([1]) B=16 rollout fwd",
             16, 128, 8, 32, [("temporal", (1,), "forward")], "rollout"),
    "cfg2": ("cfg2: DenseGCM N=128 F=32 H=32 TemporalBackedge([1,2,4]) rollout fwd",
             65536, 128, 32, 32, [("temporal", (1, 2, 4), "forward")], "rollout"),
    "cfg2-pre": ("cfg2 with RayDenseGCM's Linear preprocessor (SURVEY 8(f) rank 2): DenseGCM(preprocessor=Linear(32,32)) "
                 "N=128 H=32 TemporalBackedge([1,2,4]) rollout fwd", 65536, 128, 32, 32, [("temporal", (1, 2, 4), "forward")],
                 "rollout"),
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

    def stop(self, t0=None, t1=None):
        if self.index is None:
            return None
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for t, r in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1 + 0.1)]
        if not rows:
            rows = [r for _, r in self.rows]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


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


def cpu_reference_rate(workload, batch, steps, warm):
    """The oracle port of the reference (oracle/gcm_oracle.py) on the host cores."""
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


def run_reference(args, rank):
    if rank != 0:
        return
    desc = WORKLOADS[args.workload][0]
    rate, per, cores, what = cpu_reference_rate(args.workload, args.cpu_batch, args.steps, min(args.warmup, 4))
    sample = f"oracle port of the reference ({what}; GPU arm runs {args.batch or WORKLOADS[args.workload][1]} graphs per GPU)"
    print(json.dumps({
        "impl": "reference", "metric": "GCM env-steps/sec (fwd)", "value": rate, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "batch_per_step": args.cpu_batch, "timing": "host wall clock, CPU only"},
        "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


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
    ap.add_argument("--cache", default="bf16", choices=["bf16", "f32"],
                    help="cfg3: element type of the per-node cache (bf16 = the config's stated precision)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    desc, B0, N, F, H, spec, mode = WORKLOADS[args.workload]
    if args.warmup is None:
        args.warmup = N + 8 if mode == "rollout" else 3   # fill the graphs, then steady state
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        if args.workload == "cfg5":
            args.cpu_batch = min(args.cpu_batch, 16)
        elif args.workload != "cfg2":
            args.cpu_batch = min(args.cpu_batch, 64)
        run_reference(args, rank)
        return

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    B = args.batch or B0
    K, W = args.steps, args.warmup
    gen = torch.Generator().manual_seed(1002 + rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # rank 0 samples the clocks of every GPU of the job (local ranks 0 .. world-1 of this node)
    sampler = ClockSampler(",".join(str(i) for i in range(world)) if rank == 0 else None)
    sampler.start()
    extra, launches, e2e, kern_ms, kernel_name, host_us = None, K, None, None, None, None
    unit_per_step = B

    if mode == "rollout":
        mod = build_dense(dev, N, F, H, spec, pre=args.workload == "cfg2-pre")
        n_obs = 16 if args.workload != "cfg4-euclid" else 4
        obs_host = synth_obs(gen, n_obs, B, F, spec).pin_memory()
        obs_dev = obs_host.to(dev)
        obs_steps = [obs_dev[i] for i in range(n_obs)]       # per-step views made once (2 us of host time per step)
        hidden = None
        with torch.no_grad():
            for i in range(W):
                belief, hidden = mod(obs_steps[i % n_obs], hidden)
            barrier()
            t_lo = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(K):
                belief, hidden = mod(obs_steps[i % n_obs], hidden)
            e1.record()
            barrier()
            total_ms = max_over_ranks(e0.elapsed_time(e1))
            # kernel-only duration: the SAME public-API calls, queued behind a spinning blocker kernel so the
            # host runs ahead and the K step kernels execute back to back; CUDA events on the launching stream
            from gcm import _cabi
            lib = _cabi.lib()
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
            if args.workload == "cfg4-euclid":
                kernel_name = "k_euclid_tc (cross-batch mean distance, 3xTF32 tcgen05) + " + kernel_name
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

            hidden = e2e_steps(2 * R, hidden)
            e2e_runs = []
            for _ in range(3):                      # PCIe / host jitter: median of three timed passes of K steps
                barrier()
                e0.record()
                hidden = e2e_steps(K, hidden)
                e1.record()
                barrier()
                e2e_runs.append(max_over_ranks(e0.elapsed_time(e1)))
            e2e_ms = sorted(e2e_runs)[1]
            t_hi = time.perf_counter()
        e2e = {"value": B * world * K / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": B * F * 4,
               "d2h_bytes_per_step": B * H * 4, "ms_per_step": e2e_ms / K,
               "ms_per_step_runs": [r / K for r in e2e_runs],
               "how": "H2D / step kernel / D2H on three streams, 4-deep buffer ring, pinned host memory"}
        hidden.claim().check_flags()
    elif mode == "bptt":
        from gcm import dist as gdist

        T = 64
        mod = build_dense(dev, N, F, H, spec)
        mod.bptt_capacity = T
        mod.compute_dtype = torch.bfloat16 if args.cache == "bf16" else None
        opt = torch.optim.SGD(mod.parameters(), lr=1e-3)
        obs_host = (0.5 * torch.randn(T, B, F, generator=gen)).pin_memory()
        obs_dev = obs_host.to(dev)
        nn0 = torch.full((B,), N - T, dtype=torch.long, device=dev)            # pre-filled with N - T nodes
        nodes0 = 0.5 * torch.randn(B, N, F, device=dev)
        nodes0[:, N - T:] = 0
        adj0 = torch.zeros(B, N, N, device=dev)
        adj0[:, : N - T, : N - T] = 1                                            # DenseEdge history: all ones

        seq = args.workload == "cfg3-seq"
        obs_bt = obs_dev.transpose(0, 1).contiguous() if seq else None         # [B, T, F] for the sequence entry

        def window(obs):
            hidden = (nodes0, adj0, torch.zeros(0, device=dev), nn0)
            opt.zero_grad(set_to_none=True)
            if seq:
                beliefs, hidden = mod.forward_sequence(obs_bt, hidden)
                tot = beliefs.mean() * T
            else:
                tot = 0
                for t in range(T):
                    belief, hidden = mod(obs[t], hidden)
                    tot = tot + belief.mean()
            (tot / T).backward()
            gdist.allreduce_grads(mod.parameters(), average=True)               # the one NCCL collective
            opt.step()
            return tot

        for _ in range(min(W, 2)):
            window(obs_dev)
        barrier()
        t_lo = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            window(obs_dev)
        e1.record()
        barrier()
        total_ms = max_over_ranks(e0.elapsed_time(e1))
        t_hi = time.perf_counter()
        unit_per_step = B * T
        from gcm import _cabi
        n_l0 = _cabi.lib().gcm_launch_count()
        window(obs_dev)
        torch.cuda.synchronize()
        launches = int(_cabi.lib().gcm_launch_count() - n_l0) * K
        # dominant kernel: the forward pass over the per-node cache (one launch per step), timed alone on the state
        # the last window left behind (t = N - 1: the longest stream of the window)
        from gcm import ones as _ones
        with torch.no_grad():
            _, hid = mod(obs_dev[0], (nodes0, adj0, torch.zeros(0, device=dev), nn0))
            for t in range(1, T):
                _, hid = mod(obs_dev[t], hid)
        kern_ms = _ones.time_fwd_kernel(mod.fused_plan(), hid.claim())
        extra = {"window_ms": total_ms / K, "fwd_kernel_share_of_window": kern_ms * T / (total_ms / K)}
        kernel_name = ("k_ones_fwd (1 launch per step; the backward is ONE k_ones_window_bwd per window, "
                       "see DESIGN.md section 3)")
        if seq:
            extra["note"] = ("the sequence entry replaces the 64 k_ones_fwd launches by ONE k_ones_window_fwd (cache rows "
                             "read once, MUFU-bound); kernel_ms / frac above are those of the per-step kernel for reference")
    else:  # sparse, all-at-once
        mod = build_sparse(dev, N, F, H)
        x = torch.randn(B, N, F, generator=gen)
        x[..., 0:2] = torch.cumsum(0.1 * torch.randn(B, N, 2, generator=gen), dim=1)
        x_host = x.pin_memory()
        x_dev = x_host.to(dev)
        taus = torch.full((B,), N, dtype=torch.long, device=dev)
        train = args.workload == "cfg5-train"
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
        barrier()
        t_lo = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            out, hid = call()
        e1.record()
        barrier()
        total_ms = max_over_ranks(e0.elapsed_time(e1))
        t_hi = time.perf_counter()
        E = int(hid[1]._nnz())
        extra = (B * N, E)
        unit_per_step = B * N
        launches = K * 5
        kernel_name = "k_graphconv_fwd x2 (+ edge build)" + (" + backward (k_linear2, k_outer_reduce, k_graphconv_bwd_gather)" if train else "")
        kern_ms = total_ms / K

    value = unit_per_step * world * K / (total_ms * 1e-3)
    clocks = sampler.stop(t_lo, t_hi)

    if rank == 0:
        pk, pk_kind = peaks()
        bound, algo, algo_unit = algorithmic(args.workload, B, N, F, H, extra)
        if algo_unit == "bytes":
            achieved, peak, unit = algo / (kern_ms * 1e-3) / 1e9, pk["hbm_gbs"], "GB/s"
        else:
            # tf32 dense peak = half of the measured bf16 peak (MEASURED_PEAKS.json has no tf32 entry of its own)
            achieved, peak, unit = algo / (kern_ms * 1e-3) / 1e12, pk["bf16_tflops"] / 2.0, "TFLOP/s"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(args.workload)
        line = {
            "metric": "GCM env-steps/sec (fwd+bwd)" if (mode == "bptt" or args.workload == "cfg5-train") else "GCM env-steps/sec (fwd)",
            "value": value, "unit": "env-steps/s" if mode != "sparse" else "node-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 cache / f32 accumulate" if (mode == "bptt" and args.cache == "bf16") else "f32",
            "data": "synthetic",
            "config": {"workload": desc, "batch_per_gpu": B, "graph_size": N, "obs_size": F, "hidden": H,
                       "state": "in-place node log + bit-packed adjacency; steady state (graphs full)"
                       if mode == "rollout" else mode,
                       "l2": f"per-GPU state {B * N * F * 4 / 1e6:.0f} MB vs 126 MB L2; fresh observations every step",
                       "parallelism": f"batch-sharded x{world}, no data-path collective"},
            "clocks": clocks, "gpu_launches": launches,
            "roofline": {"bound": "hbm" if bound == "hbm" else "tensor", "achieved": achieved, "peak": peak, "unit": unit,
                         "frac": achieved / peak, "traffic": traffic, "peak_source": pk_kind if unit == "GB/s"
                         else pk_kind + " bf16 dense / 2 = tf32 dense; achieved counts the 3 tf32 MMAs of every 3xTF32 product", "kernel": kernel_name, "kernel_ms": kern_ms,
                         "algorithmic_per_launch": algo, "host_us_per_call": host_us},
        }
        if e2e is not None:
            line["e2e"] = e2e
        else:
            line["e2e"] = {"value": value, "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                           "note": "device-resident only for this auxiliary workload"}
        if isinstance(extra, dict):
            line["roofline"].update(extra)
        elif extra is not None:
            line["config"]["flat_nodes"], line["config"]["edges"] = extra
        if world == 1 and not args.no_cpu_baseline:
            cb = {"cfg2": args.cpu_batch, "cfg5": 8}.get(args.workload, 64)
            rate, per, cores, what = cpu_reference_rate(args.workload, cb, 6 if mode != "sparse" else 1, 2)
            line["cpu_baseline"] = {"value": rate, "unit": line["unit"], "cores": cores, "kind": "port",
                                    "sample": f"oracle port of the reference, {what}, {per * 1e3:.1f} ms/step"}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
