"""Turns gpurun_out/{launches.csv,*.ncu-rep} into the compact text summaries committed under profiles/.
    python tools/ncu_summary.py <tag>      (run in the build container; needs `ncu` for --page raw)
"""
import collections
import csv
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "sm__cycles_elapsed.max", "gpc__cycles_elapsed.avg.per_second", "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def launches():
    path = os.path.join(OUT, f"launches_{tag}.csv")
    if not os.path.exists(path):
        path = os.path.join(OUT, "launches.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name", "gpu__time_duration.sum") != "gpu__time_duration.sum":
            continue                      # the round-2 capture carries DRAM byte counters in the same CSV
        name = row["Kernel Name"]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        m = re.search(r"gemm_(?:tcgen05|pair)_kernel<\(?(?:int\))?(\d+)", name)
        key = f"gemm_pair_kernel<EPI={m.group(1)}>" if m else re.sub(r"\(.*", "", name).replace("void ", "")
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(os.path.join(REPO, "profiles", f"{tag}_launches.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, one warm step (batch 32 @518), "
                f"cold-cache serialised launches: compare SHARES\n# total {tot / 1e3:.2f} ms over "
                f"{sum(a[0] for a in agg.values())} launches\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:52s} n={n:4d} total={t / 1e3:8.3f} ms share={100 * t / tot:5.1f}% avg={t / n:8.1f} us\n")


def report(rep):
    path = os.path.join(OUT, rep + ".ncu-rep")
    if not os.path.exists(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(REPO, "profiles", f"{tag}_{rep}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ; {rep}.ncu-rep ; selected raw metrics per launch\n")
        for row in rows[2:]:
            f.write("--- " + row[hdr.index("Kernel Name")][:100] + "\n")
            for k in KEEP:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"  {k} [{units[i]}] = {row[i]}\n")


launches()
for r in sys.argv[2:] or ["prof_gemm", "prof_attn"]:
    report(r)
print(sorted(os.listdir(os.path.join(REPO, "profiles"))))
