"""Per-instruction stall summary of one kernel from an .ncu-rep (needs -lineinfo / --import-source on).
    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, r = rows[0], rows[1], rows[2]
for k in ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpc__cycles_elapsed.avg.per_second",
          "dram__bytes_read.sum", "dram__bytes_write.sum"]:
    for i, x in enumerate(h):
        if x == k:
            print(f"{k} [{u[i]}] = {r[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]
data = rows[2:]
ix = {k: i for i, k in enumerate(h)}
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(x[ix["# Samples"]]) for x in data)
print("total samples", tot)
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:topn]
for i in sorted(top):
    x = data[i]
    s = {k[6:]: int(x[ix[k]]) for k in stalls if int(x[ix[k]]) > 0}
    print(i, x[ix["Source"]].strip()[:64].ljust(64), x[ix["# Samples"]].rjust(6), x[ix["Instructions Executed"]].rjust(9),
          dict(sorted(s.items(), key=lambda kv: -kv[1])[:3]))
agg = {k[6:]: sum(int(x[ix[k]]) for x in data) for k in stalls}
print(sorted(agg.items(), key=lambda kv: -kv[1]))
