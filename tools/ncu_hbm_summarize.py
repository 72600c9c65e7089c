"""gpurun_out/<tag>_hbm.ncu-rep (tools/profile_hbm.sh) -> profiles/<tag>_hbm_kernels.txt: per kernel NAME the number of
launches in one step, total time, DRAM bytes (read + write), achieved DRAM GB/s = bytes / time, and that as a
fraction of the measured HBM peak (MEASURED_PEAKS.json).  usage: python tools/ncu_hbm_summarize.py <tag>"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rep = os.path.join(ROOT, "gpurun_out", f"{tag}_hbm.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
         "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}


def col(r, name):
    j = hdr.index(name)
    return float(r[j].replace(",", "")) * scale.get(units[j], 1.0)


try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    src = "measured (MEASURED_PEAKS.json)"
except Exception:
    peak, src = 6650.0, "fallback (B200_PROFILING.md)"
agg = collections.OrderedDict()
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "").replace("<unnamed>::", "")[:60]
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += col(r, "gpu__time_duration.sum")
    a[2] += col(r, "dram__bytes_read.sum")
    a[3] += col(r, "dram__bytes_write.sum")
    a[4] = max(a[4], col(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"))
lines = [f"# {tag}: ncu --set full --clock-control none of the non-GEMM kernels of ONE eager CelebA step, B = 1024 (tools/profile_hbm.sh)",
         f"# achieved = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration.sum, summed over the launches of a kernel;",
         f"# peak = {peak} GB/s, {src}.  Times under ncu are cold-cache and serialised.",
         f"{'kernel':62s} {'launches':>8s} {'time_us':>9s} {'read_MB':>9s} {'write_MB':>9s} {'GB/s':>8s} {'of_peak':>8s} {'max_dram%':>9s}"]
for k, (n, t, rd, wr, pct) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = (rd + wr) / t / 1e9 if t > 0 else 0.0
    lines.append(f"{k:62s} {n:8d} {t * 1e6:9.1f} {rd / 1e6:9.1f} {wr / 1e6:9.1f} {gbs:8.0f} {gbs / peak:8.2f} {pct:9.1f}")
out = os.path.join(ROOT, "profiles", f"{tag}_hbm_kernels.txt")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
