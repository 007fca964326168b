#!/usr/bin/env python
"""bench.py — GCM env-steps/s on B200 (BASELINE.json metric), one process per GPU.

A "step" is one DenseGCM.forward over the whole batch of independent graphs (BASELINE.json
configs[1]: graph_size 128, hidden 32, TemporalBackedge([1,2,4]), batch 65536 rollout).  The batch
shards across ranks with no data-path collective (weak scaling: every rank holds --batch graphs).

  value   env-steps/s with the step's observations already resident in HBM (CUDA events, max over ranks)
  e2e     same metric through the public API with HOST observations: every step copies its [B,F]
          observation from pinned host memory and reads the [B,H] belief back
  roofline  algorithmic bytes/step (1440 B per graph-step for cfg 2) / kernel time
  cpu_baseline  the oracle port of the reference step on the host cores, bounded sample

`--impl reference` times the oracle port (the reference's algorithm on CPU) for the same metric.
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

WORKLOAD = "cfg2: DenseGCM graph_size=128 F=32 H=32 TemporalBackedge([1,2,4]) rollout fwd"
N, F, H, HOPS = 128, 32, 32, (1, 2, 4)
ALGO_BYTES_PER_GRAPH_STEP = 1440  # SURVEY.md §8(d) primary (k-hop) figure for cfg 2


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_module(dev):
    from gcm.edge_selectors.temporal import TemporalBackedge
    from gcm.gcm import DenseGCM
    from gcm.nn import DenseGraphConv

    class GNN(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gc0 = DenseGraphConv(F, H)
            self.gc1 = DenseGraphConv(H, H)
            self.act = torch.nn.Tanh()

        def forward(self, x, adj, weights, B, N):
            x = self.act(self.gc0(x, adj))
            return self.act(self.gc1(x, adj))

    torch.manual_seed(7)
    return DenseGCM(GNN().to(dev), edge_selectors=TemporalBackedge(list(HOPS)), graph_size=N)


def oracle_params(mod):
    g = mod.gnn
    return {"w_rel1": g.gc0.lin_rel.weight, "b1": g.gc0.lin_rel.bias, "w_root1": g.gc0.lin_root.weight,
            "w_rel2": g.gc1.lin_rel.weight, "b2": g.gc1.lin_rel.bias, "w_root2": g.gc1.lin_root.weight}


def cpu_reference_rate(batch, steps, warm):
    """The oracle port of the reference step (oracle/gcm_oracle.py) on the host cores, steady state."""
    import gcm_oracle as oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    p = oracle.make_params(F, H)
    spec = [("temporal", HOPS, "forward")]
    gen = torch.Generator().manual_seed(1002)
    # start full so every timed step includes the overflow shift, like the GPU arm's steady state
    hidden = (torch.randn(batch, N, F, generator=gen), torch.zeros(batch, N, N), torch.zeros(0),
              torch.full((batch,), N, dtype=torch.long))
    obs = torch.randn(batch, F, generator=gen)
    with torch.no_grad():
        for _ in range(warm):
            _, hidden = oracle.dense_gcm_step(obs, hidden, spec, p, graph_size=N)
        t0 = time.perf_counter()
        for _ in range(steps):
            _, hidden = oracle.dense_gcm_step(obs, hidden, spec, p, graph_size=N)
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, cores


def run_reference(args, rank, world):
    if rank != 0:
        return
    batch = args.cpu_batch
    rate, per, cores = cpu_reference_rate(batch, args.steps, args.warmup)
    sample = f"oracle port of the reference step, B={batch} graphs (of {args.batch}), full graphs (wrap every step)"
    line = {
        "impl": "reference", "metric": "GCM env-steps/sec (fwd)", "value": rate, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": batch, "timing": "host wall clock, CPU only"},
        "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=136)   # N + 8: fill the graphs, then steady state (wrapping)
    ap.add_argument("--batch", type=int, default=65536, help="graphs per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-batch", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    mod = build_module(dev)
    gen = torch.Generator().manual_seed(1002 + rank)
    K, W = args.steps, args.warmup
    n_obs = 16  # distinct observation batches cycled through (fresh data every step)
    obs_host = torch.randn(n_obs, B, F, generator=gen).pin_memory()
    obs_dev = obs_host.to(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (value) ----------------
    hidden = None
    with torch.no_grad():
        for i in range(W):
            belief, hidden = mod(obs_dev[i % n_obs], hidden)
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        ev[0].record()
        for i in range(K):
            belief, hidden = mod(obs_dev[i % n_obs], hidden)
            ev[i + 1].record()
        barrier()
        total_ms = ev[0].elapsed_time(ev[K])
        clocks = sampler.stop()
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = B * world * K / (total_ms * 1e-3)

    # kernel-only duration: consecutive launches without host work in between (CUDA graph replay)
    g = torch.cuda.CUDAGraph()
    with torch.no_grad():
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            belief, hidden = mod(obs_dev[0], hidden)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            for i in range(8):
                belief, hidden = mod(obs_dev[i % n_obs], hidden)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(4, K // 8)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        kern_ms = e0.elapsed_time(e1) / (reps * 8)

    # ---------------- end to end through the public API with host buffers ----------------
    belief_host = torch.empty(B, H).pin_memory()
    with torch.no_grad():
        for i in range(3):
            belief, hidden = mod(obs_host[i % n_obs].to(dev, non_blocking=True), hidden)
            belief_host.copy_(belief, non_blocking=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            belief, hidden = mod(obs_host[i % n_obs].to(dev, non_blocking=True), hidden)
            belief_host.copy_(belief, non_blocking=True)
        e1.record()
        barrier()
        e2e_ms = e0.elapsed_time(e1)
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = B * world * K / (float(t.item()) * 1e-3)

    if rank == 0:
        pk, pk_kind = peaks()
        algo = ALGO_BYTES_PER_GRAPH_STEP * B
        achieved = algo / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": "GCM env-steps/sec (fwd)", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "graph_size": N, "obs_size": F, "hidden": H,
                       "state": "in-place ring + bit-packed adjacency, graphs full (overflow every step)",
                       "l2": f"state {B * N * F * 4 / 1e6:.0f} MB per GPU > 126 MB L2; fresh obs each step",
                       "parallelism": f"batch-sharded x{world}, no collective"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": B * F * 4,
                    "d2h_bytes_per_step": B * H * 4},
            "gpu_launches": K,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / pk["hbm_gbs"], "traffic": None, "peak_source": pk_kind,
                         "kernel": "k_step_temporal<32>", "kernel_ms": kern_ms,
                         "algo_bytes_per_launch": algo},
        }
        if world == 1 and not args.no_cpu_baseline:
            rate, per, cores = cpu_reference_rate(args.cpu_batch, 8, 2)
            line["cpu_baseline"] = {
                "value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port",
                "sample": f"oracle port, B={args.cpu_batch} full graphs, 8 steps ({per * 1e3:.1f} ms/step)"}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
