"""Runs bench.py (ours arm) and prints a one-line digest: images/s, ms/step and per-kernel-class ms.
    python tools/bench_brief.py [bench.py args]"""
import json
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--no-cpu-baseline", *sys.argv[1:]],
                     capture_output=True, text=True)
if out.returncode != 0:
    print("bench failed:", out.stderr[-2000:])
    sys.exit(1)
d = json.loads(out.stdout.strip().splitlines()[-1])
kc = d["kernel_classes"]
print(f"{os.environ.get('AACLIP_NVCC_EXTRA', ''):44s} {d['value']:7.1f} img/s  e2e {d['e2e']['value']:7.1f}  {d['ms_per_step']:6.2f} ms/step  "
      f"gemm {kc['gemm']['ms_per_step']:5.2f} ({kc['gemm']['tflops']:.0f} TF/s)  attn {kc['attention']['ms_per_step']:5.2f}  "
      f"ln {kc['layernorm']['ms_per_step']:4.2f}  other {kc['other']['ms_per_step']:4.2f}  clk {d['clocks']['sm_mhz']}\n    " +
      "  ".join(f"{k[5:]} {kc[k]['ms_per_step']:5.2f} ({kc[k]['tflops']:.0f})" for k in ("gemm_qkv", "gemm_out", "gemm_fc", "gemm_proj", "gemm_misc")),
      flush=True)
