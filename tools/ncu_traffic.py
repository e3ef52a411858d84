"""Turns the per-launch ncu CSV of one warm step (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum;
tools/r02_profile.sh) into
  profiles/<tag>_step_traffic.txt : per kernel: launches, time, DRAM bytes per launch, achieved GB/s (cold-cache,
                                    serialised launches: the bytes are what counts, not the absolute times)
  profiles/gemm_traffic.json      : measured DRAM bytes per GEMM launch (launch-weighted mean over the step's 108
                                    gemm_pair_kernel launches), the number bench.py reports as roofline.traffic
    python tools/ncu_traffic.py <tag> [S] [B] [dtype]
"""
import collections
import csv
import json
import os
import re
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 518
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
dtype = sys.argv[4] if len(sys.argv) > 4 else "fp16"
path = os.path.join(REPO, "gpurun_out", f"traffic_{tag}.csv")
lines = [l for l in open(path) if not l.startswith("==")]
per_launch = collections.OrderedDict()
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    if row["Metric Name"] == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)           # -> us
    else:
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)  # -> bytes
    per_launch.setdefault(row["ID"], {"name": row["Kernel Name"]})[row["Metric Name"]] = v
agg = collections.OrderedDict()
for rec in per_launch.values():
    name = rec["name"]
    m = re.search(r"gemm_(?:tcgen05|pair)_kernel<\(?(?:int\))?(\d+)", name)
    key = f"gemm_pair_kernel<EPI={m.group(1)}>" if m else re.sub(r"\(.*", "", name).replace("void ", "")
    a = agg.setdefault(key, dict(n=0, us=0.0, rd=0.0, wr=0.0))
    a["n"] += 1
    a["us"] += rec.get("gpu__time_duration.sum", 0.0)
    a["rd"] += rec.get("dram__bytes_read.sum", 0.0)
    a["wr"] += rec.get("dram__bytes_write.sum", 0.0)
tot_us = sum(a["us"] for a in agg.values())
tot_b = sum(a["rd"] + a["wr"] for a in agg.values())
with open(os.path.join(REPO, "profiles", f"{tag}_step_traffic.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one warm "
            f"step (batch {B} @{S}, {dtype} operands); launches are serialised with cold caches: compare SHARES and BYTES\n")
    f.write(f"# total {tot_us / 1e3:.2f} ms, {tot_b / 1e9:.2f} GB of DRAM traffic over {sum(a['n'] for a in agg.values())} launches\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        by = (a["rd"] + a["wr"]) / a["n"]
        f.write(f"{k:44s} n={a['n']:3d} share={100 * a['us'] / tot_us:5.1f}% avg={a['us'] / a['n']:7.1f} us  "
                f"dram read {a['rd'] / a['n'] / 1e6:7.1f} MB + write {a['wr'] / a['n'] / 1e6:7.1f} MB per launch "
                f"= {by / (a['us'] / a['n']) / 1e3:6.0f} GB/s\n")
g = [a for k, a in agg.items() if k.startswith("gemm_pair_kernel")]
n = sum(a["n"] for a in g)
out_path = os.path.join(REPO, "profiles", "gemm_traffic.json")
table = json.load(open(out_path)) if os.path.exists(out_path) else {}
table[f"S{S}_B{B}_{dtype}"] = {
    "dram_bytes_per_gemm_launch": sum(a["rd"] + a["wr"] for a in g) / n, "gemm_launches": n,
    "source": f"profiles/{tag}_step_traffic.txt (ncu dram__bytes_read.sum + dram__bytes_write.sum, one warm step)",
    "per_epilogue": {k: (a["rd"] + a["wr"]) / a["n"] for k, a in agg.items() if k.startswith("gemm_pair_kernel")}}
json.dump(table, open(out_path, "w"), indent=1)
print(open(os.path.join(REPO, "profiles", f"{tag}_step_traffic.txt")).read())
print(json.dumps(table[f"S{S}_B{B}_{dtype}"], indent=1))
