"""gpurun_out/<tag>_step_kernels.csv and <tag>_tensor.csv (tools/profile_hbm.sh) ->
profiles/<tag>_step_kernels.txt : per kernel NAME of one eager CelebA step at B = 1024: launches, total time, DRAM bytes
    (read + write), achieved DRAM GB/s = bytes / time and that as a fraction of the measured HBM peak;
profiles/<tag>_tensor_pipe.txt  : per big-layer launch of tools/bench_layers.py: time, tensor-pipe active %, DRAM bytes.
usage: python tools/ncu_hbm_summarize.py <tag>"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3,
         "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "%": 1.0, "": 1.0}


def launches(path):
    """-> list of {name, metric: value (SI)} per launch, from ncu's long-format CSV log"""
    rows = list(csv.reader(open(path, errors="replace")))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
    h = rows[start]
    iid, ik, im, iu, iv = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
    out = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= iv:
            continue
        d = out.setdefault(r[iid], {"name": re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("<unnamed>::", "")[:64]})
        try:
            d[r[im]] = float(r[iv].replace(",", "")) * SCALE.get(r[iu], 1.0)
        except ValueError:
            pass
    return list(out.values())


try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    src = "measured (MEASURED_PEAKS.json)"
except Exception:
    peak, src = 6650.0, "fallback (B200_PROFILING.md)"

p1 = os.path.join(ROOT, "gpurun_out", f"{tag}_step_kernels.csv")
if os.path.exists(p1):
    agg = collections.OrderedDict()
    for d in launches(p1):
        a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    lines = [f"# {tag}: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... --clock-control none of",
             "# ONE eager CelebA step at B = 1024 (tools/profile_hbm.sh; bench.py --no-graph --steps 1, profiler window = the timed step).",
             "# achieved = (dram read + write bytes) / kernel time, summed over the launches of a kernel name; of_peak vs "
             f"{peak} GB/s, {src}.",
             f"# {sum(a[0] for a in agg.values())} launches, {tot * 1e3:.2f} ms of serialised kernel time (cold caches: compare shares, not absolutes)",
             f"{'kernel':66s} {'launches':>8s} {'time_us':>9s} {'share':>6s} {'read_MB':>9s} {'write_MB':>9s} {'GB/s':>7s} {'of_peak':>7s}"]
    for k, (n, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = (rd + wr) / t / 1e9 if t > 0 else 0.0
        lines.append(f"{k:66s} {n:8d} {t * 1e6:9.1f} {100 * t / tot:5.1f}% {rd / 1e6:9.1f} {wr / 1e6:9.1f} {gbs:7.0f} {gbs / peak:7.2f}")
    open(os.path.join(ROOT, "profiles", f"{tag}_step_kernels.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:45]))

p2 = os.path.join(ROOT, "gpurun_out", f"{tag}_tensor.csv")
if os.path.exists(p2):
    lines = [f"# {tag}: tensor-pipe activity of every tcgen05 launch of tools/bench_layers.py 1024 (one launch per case, in the order",
             "# of gpurun_out/<tag>_layers_plain.log: c = 32, 128, 256, 512; fprop x5 variants, dgrad x5, wgrad), ncu --clock-control none",
             f"{'#':>3s} {'kernel':64s} {'grid':>5s} {'time_us':>9s} {'tensor_pipe_%':>13s} {'dram_MB':>9s}"]
    for i, d in enumerate(launches(p2)):
        lines.append(f"{i:3d} {d['name']:64s} {int(d.get('launch__grid_size', 0)):5d} {d.get('gpu__time_duration.sum', 0) * 1e6:9.1f} "
                     f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):13.1f} "
                     f"{(d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)) / 1e6:9.1f}")
    open(os.path.join(ROOT, "profiles", f"{tag}_tensor_pipe.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
