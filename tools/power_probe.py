"""Steady-state board power and SM clock of each hot kernel run alone in a loop (NVML, 50 ms sampling, ~2.5 s per
kernel after 1 s of settling): which kernels sit at the power cap, and what a launch costs in joules.
    python tools/power_probe.py"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml
import torch

import aaclip_b200  # noqa: F401
from aaclip_b200 import _lib, engine

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(int(os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0]))
B, dev = 32, "cuda"
M = 1370 * B


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.p, self.c, self.stop_evt = [], [], threading.Event()

    def run(self):
        while not self.stop_evt.is_set():
            self.p.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
            self.c.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.stop_evt.wait(0.05)


def probe(name, fn, flops):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t_end = time.time() + 1.0
    while time.time() < t_end:          # settle
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
    s = Sampler()
    s.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    e0.record()
    t_end = time.time() + 2.5
    while time.time() < t_end:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    s.stop_evt.set()
    s.join()
    ms = e0.elapsed_time(e1) / n
    p = sorted(s.p)[len(s.p) // 2]
    c = sorted(s.c)[len(s.c) // 2]
    print(f"{name:28s} {ms:7.3f} ms/launch  {flops / ms / 1e9:7.0f} TFLOP/s  {p:6.0f} W  {c:5d} MHz  {p * ms:7.1f} mJ/launch  "
          f"{p * ms * 1e-3 / flops * 1e12 if flops else 0:5.2f} pJ/flop", flush=True)
    time.sleep(1.0)


x = torch.randn(M, 1024, device=dev)
stats = torch.stack([x.view(M, 8, 128).sum(-1), (x * x).view(M, 8, 128).sum(-1)], -1).contiguous()
x16 = x.half()
w16, cs, bf = engine.ln_fold(torch.randn(3072, 1024, device=dev) * 0.03, torch.zeros(3072, device=dev),
                             torch.ones(1024, device=dev), torch.zeros(1024, device=dev))
probe("qkv  LN-fold", lambda: engine.gemm16_ln(x16, w16, _lib.EPI_LN_BIAS_BF16, bf, cs, stats, 1024), 2.0 * M * 3072 * 1024)
w16f, csf, bff = engine.ln_fold(torch.randn(4096, 1024, device=dev) * 0.03, torch.zeros(4096, device=dev),
                                torch.ones(1024, device=dev), torch.zeros(1024, device=dev))
probe("fc   LN-fold + GELU", lambda: engine.gemm16_ln(x16, w16f, _lib.EPI_LN_BIAS_GELU_BF16, bff, csf, stats, 1024), 2.0 * M * 4096 * 1024)
for (K, nm) in [(1024, "out  resid+x16+stats"), (4096, "proj resid+x16+stats")]:
    a = (torch.randn(M, K, device=dev) * 0.5).half()
    w = (torch.randn(1024, K, device=dev) * K ** -0.5).half()
    bias = torch.zeros(1024, device=dev)
    probe(nm, lambda: engine.gemm16_resid(a, w, bias, x), 2.0 * M * 1024 * K)
    del a, w
qkv = torch.randn(B, 1370, 3072, device=dev).half()
probe("attention", lambda: engine.attention(qkv), 4.0 * B * 16 * 1370 * 1370 * 64)
a8 = torch.randn(8192, 8192, device=dev).half()
b8 = torch.randn(8192, 8192, device=dev).half()
probe("cuBLAS fp16 8192^3 (ref.)", lambda: torch.matmul(a8, b8), 2.0 * 8192 ** 3)
# cuBLAS on the step's own shapes (plain GEMM, no fused epilogue): the energy floor a library kernel reaches there
for (N, K, nm) in [(3072, 1024, "cuBLAS qkv shape"), (4096, 1024, "cuBLAS fc shape"), (1024, 4096, "cuBLAS proj shape"),
                   (1024, 1024, "cuBLAS out shape")]:
    a = (torch.randn(M, K, device=dev) * 0.5).half()
    w = (torch.randn(N, K, device=dev) * K ** -0.5).half()
    o = torch.empty(M, N, device=dev, dtype=torch.float16)
    probe(nm, lambda: torch.matmul(a, w.t(), out=o), 2.0 * M * N * K)
    del a, w, o
