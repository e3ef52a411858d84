"""Steady-state (power-capped) cost of the attention kernel for pre-built library variants: ms, W, MHz and mJ per launch.
The variants are built in the dev container (tools/build_variants.sh) into tools/variants/*.so and travel with the repo.
    python tools/attn_power.py tools/variants/*.so"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, threading, time
sys.path.insert(0, %r)
import pynvml, torch
import aaclip_b200
from aaclip_b200 import engine
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(int(os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0]))
qkv = (torch.randn(32, 1370, 3072, device="cuda") * 1.5).half()
fn = lambda: engine.attention(qkv)
for _ in range(3): fn()
torch.cuda.synchronize()
t_end = time.time() + 1.5
while time.time() < t_end:
    for _ in range(20): fn()
    torch.cuda.synchronize()
P, C, stop = [], [], threading.Event()
def samp():
    while not stop.is_set():
        P.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3); C.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)); stop.wait(0.05)
th = threading.Thread(target=samp, daemon=True); th.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 0; e0.record(); t_end = time.time() + 3.0
while time.time() < t_end:
    for _ in range(20): fn()
    n += 20; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize(); stop.set(); th.join()
ms = e0.elapsed_time(e1) / n
p = sorted(P)[len(P) // 2]; c = sorted(C)[len(C) // 2]
print(f"{os.path.basename(os.environ['AACLIP_LIB_PATH']):34s} {ms:7.4f} ms/launch  {p:6.0f} W  {c:5d} MHz  {p * ms:7.1f} mJ/launch", flush=True)
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, AACLIP_LIB_PATH=os.path.abspath(lib))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=120)
    print(r.stdout.strip() if r.returncode == 0 else f"{lib}: failed\n{r.stderr[-800:]}", flush=True)
