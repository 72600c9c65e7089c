#!/usr/bin/env python
"""bench.py -- train images/sec of the EAD-GAN CelebA 64x64 step (G + D/Q, three phases, three
Adams; celebA/EAD-GAN_celebA.py:296-401) on N B200s, plus the kernel roofline and the CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B_per_gpu] [--impl ours|reference]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
              --master-port P bench.py --gpus N --steps K --warmup W

One JSON line on rank 0.  `value`: inputs resident in HBM, CUDA-event timed, barrier + synchronize on
both sides, max over ranks.  `e2e`: the same step called with HOST (pinned) buffers -- H2D of the
step's inputs and D2H of its losses inside the timed region.  Weak scaling: per-GPU batch fixed.
Data is synthetic, weights random-init (no datasets/checkpoints offline).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train images/sec (G+D+E step, 64x64 CelebA)"
GFLOP_PER_IMG = 18.865  # useful algorithmic GFLOP / image / step


def ncu_traffic(entry_name):
    """DRAM bytes per launch of this entry point's kernel from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by tools/ncu_summarize.py), or None when it was not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f).get(entry_name)
        return None if t is None else {"bytes": t["dram_bytes_per_launch"], "capture": t["capture"],
                                       "tensor_pipe_pct": t.get("tensor_pipe_pct")}
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sust": p["bf16_tflops_sustained"],
                "src": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sust": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference(batch, steps, warmup):
    """the oracle restatement of the reference step (stock torch.nn, fp32) on the host cores."""
    from oracle import torch_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    st = O.build_celeba(seed=0, device="cpu")
    imgs = O.synth_celeba_images(batch, 0)
    rs = np.random.RandomState(0)
    times = []
    for i in range(warmup + steps):
        d = O.sample_celeba(rs, batch)
        t0 = time.perf_counter()
        O.step_celeba(st, imgs, d, record=False)
        t1 = time.perf_counter()
        if i >= warmup:
            times.append(t1 - t0)
    med = statistics.median(times)
    return batch / med, med


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.cpu_batch
    ips, med = cpu_reference(B, max(1, args.steps), max(1, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "CelebA EAD-GAN_celebA 64x64 RGB G/D step (configs[3]/[4])",
                       "batch_per_step": B, "device": "host CPU, torch oneDNN"},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{args.steps} steps of batch {B} (oracle/torch_oracle.py, pinned to the "
                                       "reference scripts; CPU images/s is flat in batch size)"},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _OUT.emit(json.dumps(line))


class _QuietStdout:
    """Everything written to fd 1 before the result (NCCL prints its version banner there, libraries may print
    warnings) goes to stderr, so that stdout carries exactly ONE line: the JSON result."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(line, flush=True)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


_OUT = None


def main():
    global _OUT
    with _QuietStdout() as _OUT:
        _main()


def _main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1024, help="per-GPU batch (weak scaling)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-batch", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default=None, help="write the per-entry-point time table here (json)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying "
                    "the captured whole-step CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    from eadgan_b200 import _lib, parallel
    from eadgan_b200.steps.celeba import CelebAStep
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sm_100a path has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dp = parallel.init_from_env()
    precision = os.environ.get("EADGAN_PRECISION", "bf16")
    B = args.batch
    Bg = B * world

    step = CelebAStep(seed=0, device=dev)
    parallel.attach(*step.optimizers())

    # global batch drawn once from the seeded host RNG, sharded contiguously (rank r: rows r*B..)
    from eadgan_b200.synthetic import celeba_images as synth_celeba_images   # the oracle is only used by cpu_baseline
    R = 2  # ring of host batches
    host = []
    for i in range(R):
        rs = np.random.RandomState(100 + i)
        imgs = synth_celeba_images(Bg, i)[rank * B:(rank + 1) * B]
        z = torch.tensor(rs.normal(0, 1, (Bg, 200)), dtype=torch.float32)[rank * B:(rank + 1) * B]
        code = torch.tensor(rs.uniform(-1, 1, (Bg, 8)), dtype=torch.float32)[rank * B:(rank + 1) * B]
        labels = torch.tensor(rs.randint(0, 10, Bg), dtype=torch.long)[rank * B:(rank + 1) * B]
        host.append(tuple(t.contiguous().pin_memory() for t in (imgs, z, code, labels)))
    resident = [tuple(t.to(dev) for t in hb) for hb in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    eager_step = step
    use_graph = not args.no_graph   # under DP the NCCL all-reduces (gradient buckets, SyncBN) are captured too
    if use_graph:
        from eadgan_b200.graph import GraphedStep
        step = GraphedStep(eager_step, resident[0], warmup=args.warmup)   # warm-up steps run inside, eagerly
    for i in range(args.warmup):
        step(*resident[i % R])
    # ---- device-resident timing ---------------------------------------------------------
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    k0 = _lib.lib().eadgan_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    window = os.environ.get("EADGAN_PROFILE_WINDOW") == "1"   # ncu --profile-from-start off: only the timed steps
    if window:
        torch.cuda.profiler.start()
    e0.record()
    for i in range(args.steps):
        step(*resident[i % R])
    e1.record()
    torch.cuda.synchronize()
    if window:
        torch.cuda.profiler.stop()
    launches = _lib.lib().eadgan_kernel_launches() - k0
    if use_graph:
        launches = step.kernels_per_replay * args.steps   # replays do not pass through the C-ABI launch counter
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    # ---- end-to-end: host buffers in, losses out, every step ------------------------------
    def from_host(hb):
        # graph: the replay wrapper copies the pinned host tensors straight into its static device inputs
        return hb if use_graph else [t.to(dev, non_blocking=True) for t in hb]

    # graph: while replay i runs, the pinned host batch of step i + 1 is copied on a copy stream (every step's H2D
    # copy and the D2H read of its losses are inside the timed region; only the very first batch is staged before)
    def e2e_step(i):
        if use_graph:
            return step(*host[i % R], prefetch=host[(i + 1) % R])
        return step(*from_host(host[i % R]))

    for i in range(2):
        e2e_step(i)
    barrier()
    d2h_bytes = 0
    e0.record()
    for i in range(args.steps):
        out = e2e_step(2 + i)
        vals = torch.stack([out["g_loss"], out["d_loss"], out["info_loss"]]).cpu()
        d2h_bytes = vals.numel() * vals.element_size()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    losses = [float(v) for v in vals]

    # ---- per-entry-point profile of one step -> dominant kernel roofline --------------------
    _lib.profile_start()
    eager_step(*resident[0])          # per-call CUDA events need the eager launch path
    prof = _lib.profile_stop()
    pk = peaks()
    conv = {k: v for k, v in prof.items() if v["flops"] > 0}
    total_ms = sum(v["ms"] for v in prof.values())
    roof = None
    if conv:
        name, r = max(conv.items(), key=lambda kv: kv[1]["ms"])
        ach = r["flops"] / (r["ms"] * 1e-3) / 1e12
        tr = ncu_traffic(name)
        roof = {"bound": "tensor", "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                "frac": ach / pk["tf_sust"], "traffic": None if tr is None else tr["bytes"],
                "traffic_source": None if tr is None else f"profiles/{tr['capture']} (ncu --set full, dram read + write per launch)",
                "ncu_tensor_pipe_pct": None if tr is None else tr["tensor_pipe_pct"],
                "algorithmic_flops_per_launch": r["flops"] / r["calls"],
                "kernel": name, "launches": r["calls"],
                "avg_launch_ms": r["ms"] / r["calls"], "share_of_step": r["ms"] / total_ms if total_ms else None,
                "peak_source": pk["src"] + ", sustained bf16 (kernel timed inside a long step)"}
    # supplementary HBM-bound roofline: the fused multi-tensor Adam, 28 B per parameter update (read p, g, m, v;
    # write p, m, v), all three optimiser steps of the iteration
    roof_hbm = None
    adam = prof.get("eadgan_adam_step")
    if adam and adam["ms"] > 0:
        n_updates = sum(p.numel() for o in eager_step.optimizers() for g in o.param_groups for p in g["params"]
                        if p.grad is not None)
        gbs = 28.0 * n_updates / (adam["ms"] * 1e-3) / 1e9
        roof_hbm = {"bound": "hbm", "kernel": "eadgan_adam_step", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": gbs / pk["hbm"], "algorithmic_bytes_per_step": 28.0 * n_updates, "launches": adam["calls"],
                    "peak_source": pk["src"]}
    if args.profile_out and rank == 0:
        with open(args.profile_out, "w") as f:
            json.dump({"batch_per_gpu": B, "precision": precision, "step_ms_sum": total_ms,
                       "entry_points": dict(sorted(prof.items(), key=lambda kv: -kv[1]["ms"]))}, f, indent=1)

    if rank != 0:
        return
    ips = Bg * args.steps / (ms * 1e-3)
    ips_e2e = Bg * args.steps / (ms_e2e * 1e-3)
    line = {
        "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "CelebA EAD-GAN_celebA 64x64 RGB G/D step, 3 phases + 3 Adam (BASELINE configs[3])",
                   "batch_per_gpu": B, "global_batch": Bg, "parallelism": f"dp{world}", "precision": precision,
                   "l2": "inputs+activations per step >> 126 MB L2 (no flush needed)",
                   "launch": "whole-step CUDA graph replay" if use_graph else "eager (one C-ABI call per kernel)",
                   "weights": "random-init seed 0"},
        "clocks": clocks,
        "e2e": {"value": ips_e2e, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": ms_e2e / args.steps,
                "input_pipeline": ("step i+1's pinned host batch is copied on a copy stream while step i's graph replays; "
                                   "every step's H2D copy and loss read-back are inside the timed region") if use_graph
                else "synchronous H2D copy in front of every step"},
        "gpu_launches": int(launches),
        "roofline": roof,
        "roofline_hbm": roof_hbm,
        "step_tensor_frac": {"achieved_tflops": GFLOP_PER_IMG * 1e9 * ips / 1e12,
                             "peak_tflops": pk["tf_sust"] * world, "frac": GFLOP_PER_IMG * 1e9 * ips / 1e12 / (pk["tf_sust"] * world)},
        "losses_last_step": losses,
    }
    if world == 1 and not args.no_cpu_baseline:
        cb, med = cpu_reference(args.cpu_batch, 4, 2)
        line["cpu_baseline"] = {"value": cb, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"4 steps of batch {args.cpu_batch} after 2 warm-up on the host CPU "
                                          "(oracle/torch_oracle.py; CPU images/s is flat in batch size)"}
    _OUT.emit(json.dumps(line))


if __name__ == "__main__":
    main()
