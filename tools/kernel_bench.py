"""Isolated timings of the hot kernels at the batch-32 @518 shapes (CUDA events, 3 warm + 10 timed launches each).
    python tools/kernel_bench.py [B]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aaclip_b200  # noqa: F401
from aaclip_b200 import _lib, engine

_lib.build()   # honours AACLIP_LIB_PATH / AACLIP_NVCC_EXTRA (tools/sweep.sh)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda"
M = 1370 * B


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (N, K, epi, nm) in [(3072, 1024, _lib.EPI_BIAS_BF16, "qkv  bias->f16"), (1024, 1024, _lib.EPI_BIAS_RESID_F32, "out  bias+resid"),
                        (4096, 1024, _lib.EPI_BIAS_GELU_BF16, "fc   bias+gelu->f16"), (1024, 4096, _lib.EPI_BIAS_RESID_F32, "proj bias+resid"),
                        (1024, 1024, _lib.EPI_LEAKY_F32, "adapter leaky->f32"), (768, 1024, _lib.EPI_F32, "seg proj ->f32")]:
    a = (torch.randn(M, K, device=dev) * 0.5).half()
    w = (torch.randn(N, K, device=dev) * K ** -0.5).half()
    bias = torch.randn(N, device=dev) * 0.1
    o16 = torch.empty(M, N, device=dev, dtype=torch.float16)
    o32 = torch.zeros(M, N, device=dev)
    ms = timeit(lambda: engine.gemm16(a, w, epi, bias=bias, out_f32=o32, out_16=o16 if epi < 3 else None))
    print(f"gemm {nm:22s} M={M} N={N} K={K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s", flush=True)

# the folded-LayerNorm forms aaclip_forward actually launches
x = torch.randn(M, 1024, device=dev)
stats = torch.stack([x.view(M, 8, 128).sum(-1), (x * x).view(M, 8, 128).sum(-1)], -1).contiguous()
x16 = x.half()
for (N, epi, nm) in [(3072, _lib.EPI_LN_BIAS_BF16, "qkv  LN-fold bias->f16"), (4096, _lib.EPI_LN_BIAS_GELU_BF16, "fc   LN-fold gelu->f16")]:
    w16, cs, bf = engine.ln_fold(torch.randn(N, 1024, device=dev) * 0.03, torch.zeros(N, device=dev),
                                 torch.ones(1024, device=dev), torch.zeros(1024, device=dev))
    ms = timeit(lambda: engine.gemm16_ln(x16, w16, epi, bf, cs, stats, 1024))
    print(f"gemm {nm:22s} M={M} N={N} K=1024: {ms:.3f} ms  {2.0 * M * N * 1024 / ms / 1e9:7.1f} TFLOP/s", flush=True)
hi, lo = engine.split_hilo(x)
for (K, nm) in [(1024, "out  resid (hi,lo)+stats"), (4096, "proj resid (hi,lo)+stats")]:
    a = (torch.randn(M, K, device=dev) * 0.5).half()
    w = (torch.randn(1024, K, device=dev) * K ** -0.5).half()
    bias = torch.randn(1024, device=dev) * 0.1
    ms = timeit(lambda: engine.gemm16_resid_hl(a, w, bias, hi, lo))
    print(f"gemm {nm:22s} M={M} N=1024 K={K}: {ms:.3f} ms  {2.0 * M * 1024 * K / ms / 1e9:7.1f} TFLOP/s", flush=True)

qkv = torch.randn(B, 1370, 3072, device=dev).half()
ms = timeit(lambda: engine.attention(qkv))
print(f"attention B={B} N=1370 H=16 d=64: {ms:.3f} ms  {4.0 * B * 16 * 1370 * 1370 * 64 / ms / 1e9:7.1f} TFLOP/s (algorithmic)")
x = torch.randn(M, 1024, device=dev)
g, b = torch.ones(1024, device=dev), torch.zeros(1024, device=dev)
ms = timeit(lambda: engine.layernorm(x, g, b))
print(f"layernorm rows={M}: {ms:.3f} ms  {M * 1024 * 6 / ms / 1e6:7.1f} GB/s")
