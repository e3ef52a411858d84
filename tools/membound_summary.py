"""gpurun_out/membound_<tag>.csv (+ membound_pre_<tag>.csv) from tools/ncu_membound.sh -> profiles/<tag>_membound_kernels.txt
    python tools/membound_summary.py <tag>"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
peak = 6462.1
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass
out = [f"# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; one warm "
       f"step (B=32 @518) + the two pre-processing kernels (32 x 1024x1024 -> 518); per-launch averages; peak = {peak} GB/s"]
for name in (f"membound_{tag}.csv", f"membound_pre_{tag}.csv"):
    path = os.path.join(ROOT, "gpurun_out", name)
    if not os.path.exists(path):
        continue
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("aaclip::", "")
        agg.setdefault((k, r["ID"]), {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    by = collections.OrderedDict()
    for (k, _), m in agg.items():
        t, u = m["gpu__time_duration.sum"]
        t *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
        mb = lambda x: m[x][0] * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(m[x][1], 1)  # noqa: E731
        b = by.setdefault(k, [0, 0.0, 0.0, 0.0])
        b[0] += 1; b[1] += t; b[2] += mb("dram__bytes_read.sum"); b[3] += mb("dram__bytes_write.sum")
    for k, (n, t, rd, wr) in by.items():
        gbs = (rd + wr) / t * 1e3
        out.append(f"{k:44s} n={n:2d} avg {t / n:7.1f} us  dram rd {rd / n:7.1f} MB wr {wr / n:7.1f} MB  -> {gbs:6.0f} GB/s "
                   f"({gbs / peak * 100:4.1f} % of peak)")
txt = "\n".join(out) + "\n"
open(os.path.join(ROOT, "profiles", f"{tag}_membound_kernels.txt"), "w").write(txt)
print(txt)
