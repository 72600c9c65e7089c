"""Device timeline of ONE graph-replayed step (torch.profiler / CUPTI kernel records; no ncu, kernels run back to back
as in the benchmark): per stream busy time, the idle gaps of the device and what sits around them, and the time per
kernel name as it is INSIDE the step (warm caches, overlapped side streams) -- the complement of the serialised ncu
launch list.  usage: timeline.py [--config celeba] [--batch 1024] [--out gpurun_out/timeline.txt]
Under torchrun every rank profiles itself; rank 0 writes the file."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="celeba")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.txt"))
    args = ap.parse_args()
    from eadgan_b200 import parallel
    from eadgan_b200.graph import GraphedStep
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    parallel.init_from_env()
    _, sampled, host = bench.build_step(args.config, dev, rank, args.batch)
    resident = [h.to(dev) for h in host]
    step = GraphedStep(sampled, [resident[0]], warmup=3)
    for i in range(5):
        step(resident[i % 2])
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(resident[0])
        step(resident[1])
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    evs.sort(key=lambda e: e.time_range.start)
    # second replay only: split at the largest start-to-start gap near the middle
    n = len(evs) // 2
    first_end = max(e.time_range.end for e in evs[:n])
    between = evs[n].time_range.start - first_end      # device idle between two back-to-back replays
    evs = evs[n:]
    t0 = evs[0].time_range.start
    t1 = max(e.time_range.end for e in evs)
    lines = [f"# {args.config} B={args.batch} world={world}: one graph replay, {len(evs)} device activities, span {(t1 - t0) / 1e3:.3f} ms"
             f" (idle between two back-to-back replays: {between:.1f} us; a two-replay burst runs at boost clocks, the 20-step"
             " benchmark under the power cap: compare shares, not the span)"]
    streams = {}
    for e in evs:
        streams.setdefault(getattr(e, "device_resource_id", 0), []).append(e)
    for sid, L in sorted(streams.items(), key=lambda kv: -len(kv[1])):
        busy = sum(e.time_range.end - e.time_range.start for e in L)
        lines.append(f"stream {sid}: {len(L)} activities, busy {busy / 1e3:.3f} ms")
    # device idle: union of all activity intervals
    cur_end = t0
    idle = []
    prev = None
    for e in evs:
        s, en = e.time_range.start, e.time_range.end
        if s > cur_end:
            idle.append((s - cur_end, prev.name if prev else "-", e.name, (cur_end - t0)))
        if en > cur_end:
            cur_end = en
            prev = e
    tot_idle = sum(g[0] for g in idle)
    lines.append(f"device idle inside the step: {tot_idle / 1e3:.3f} ms in {len(idle)} gaps "
                 f"(>= 5 us: {sum(g[0] for g in idle if g[0] >= 5) / 1e3:.3f} ms in {sum(1 for g in idle if g[0] >= 5)})")
    lines.append("largest gaps: us, at ms, after -> before")
    for g in sorted(idle, reverse=True)[:25]:
        lines.append(f"  {g[0]:8.1f} us at {g[3] / 1e3:7.3f} ms  {g[1][:60]} -> {g[2][:60]}")
    # collectives: how much of every NCCL kernel runs with NO compute kernel beside it (exposed), by kernel
    nccl = [e for e in evs if "nccl" in e.name.lower()]
    if nccl:
        comp = sorted((e.time_range.start, e.time_range.end) for e in evs if "nccl" not in e.name.lower())
        merged = []
        for s_, e_ in comp:
            if merged and s_ <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], e_)
            else:
                merged.append([s_, e_])
        def covered(a, b):
            tot = 0.0
            for s_, e_ in merged:
                if e_ <= a:
                    continue
                if s_ >= b:
                    break
                tot += min(b, e_) - max(a, s_)
            return tot
        lines.append("collectives (NCCL kernels): count, total us, exposed us (no compute kernel running beside them)")
        agg = {}
        for e in nccl:
            a, b = e.time_range.start, e.time_range.end
            d = agg.setdefault(e.name[:70], [0, 0.0, 0.0])
            d[0] += 1
            d[1] += b - a
            d[2] += (b - a) - covered(a, b)
        for name, (cnt, us, ex) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            lines.append(f"  {name:70s} {cnt:4d} {us:9.1f} {ex:9.1f}")
        big = sorted(nccl, key=lambda e: -(e.time_range.end - e.time_range.start))[:12]
        lines.append("longest collectives: us, exposed us, at ms")
        for e in big:
            a, b = e.time_range.start, e.time_range.end
            lines.append(f"  {b - a:8.1f} {(b - a) - covered(a, b):8.1f} at {(a - t0) / 1e3:7.3f} ms  {e.name[:60]}")
    by = {}
    for e in evs:
        d = by.setdefault(e.name, [0, 0.0])
        d[0] += 1
        d[1] += e.time_range.end - e.time_range.start
    lines.append("kernel time inside the step (us; sums over all streams):")
    tot = sum(v[1] for v in by.values())
    for name, (cnt, us) in sorted(by.items(), key=lambda kv: -kv[1][1])[:45]:
        lines.append(f"  {name[:90]:90s} {cnt:4d} {us:9.1f} {us / tot * 100:5.1f}%")
    lines.append(f"  total {tot / 1e3:.3f} ms over {sum(v[0] for v in by.values())} activities")
    if rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        open(args.out, "w").write("\n".join(lines) + "\n")
        print("\n".join(lines[:40]))
    if world > 1:
        os._exit(0)     # a live graph holding NCCL kernels blocks process-group teardown


if __name__ == "__main__":
    main()
