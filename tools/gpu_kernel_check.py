"""Developer diagnostic (run on the GPU box): each building-block kernel against a plain torch fp32 computation,
with enough error-pattern output to debug descriptor / layout mistakes in one round trip.

    python tools/gpu_kernel_check.py [gemm] [ln] [attn] [maps]  > gpurun_out/kernel_check.log
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aaclip_b200  # noqa: F401
from aaclip_b200 import _lib, engine

torch.manual_seed(0)
dev = torch.device("cuda")
what = set(sys.argv[1:]) or {"gemm", "ln", "attn", "maps"}
failures = []


def report(name, got, ref, tol):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    bad = ~torch.isfinite(got)
    mx = err[~bad].max().item() if (~bad).any() else float("nan")
    ok = (not bad.any().item()) and mx <= tol
    print(f"[{'OK ' if ok else 'BAD'}] {name}: max_abs_err={mx:.3e} tol={tol:.1e} ref_absmax={ref.abs().max().item():.3e} "
          f"nonfinite={int(bad.sum())}", flush=True)
    if not ok:
        failures.append(name)
    return ok, err


def gemm_case(M, N, K, epi, name):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * (K ** -0.5)).bfloat16()
    bias = torch.randn(N, device=dev) * 0.1
    ref = a.float() @ w.float().t()
    if epi == _lib.EPI_BIAS_BF16:
        out = engine.gemm_bf16(a, w, epi, bias=bias)
        ref = ref + bias
        tol = 3e-2
    elif epi == _lib.EPI_BIAS_GELU_BF16:
        out = engine.gemm_bf16(a, w, epi, bias=bias)
        ref = torch.nn.functional.gelu(ref + bias)
        tol = 3e-2
    elif epi == _lib.EPI_BIAS_QGELU_BF16:
        out = engine.gemm_bf16(a, w, epi, bias=bias)
        y = ref + bias
        ref = y * torch.sigmoid(1.702 * y)
        tol = 3e-2
    elif epi == _lib.EPI_BIAS_RESID_F32:
        res = torch.randn(M, N, device=dev)
        out = res.clone()
        cp = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        engine.gemm_bf16(a, w, epi, bias=bias, out_f32=out, out_bf16=cp)
        ref = res + ref + bias
        tol = 2e-3
        report(name + " (bf16 copy)", cp, ref, 3e-2)
    elif epi == _lib.EPI_F32:
        out = engine.gemm_bf16(a, w, epi)
        tol = 2e-3
    elif epi == _lib.EPI_LEAKY_F32:
        out = engine.gemm_bf16(a, w, epi)
        ref = torch.nn.functional.leaky_relu(ref, 0.01)
        tol = 2e-3
    elif epi == _lib.EPI_BIAS_F32:
        out = engine.gemm_bf16(a, w, epi, bias=bias)
        ref = ref + bias
        tol = 2e-3
    torch.cuda.synchronize()
    ok, err = report(name, out, ref, tol)
    if not ok:
        # error pattern by tile-local coordinates
        e = err.clone()
        e[~torch.isfinite(e)] = 1e9
        rows_bad = (e.max(dim=1).values > tol).nonzero().flatten()
        cols_bad = (e.max(dim=0).values > tol).nonzero().flatten()
        print(f"    bad rows: {rows_bad.numel()}/{M} first {rows_bad[:12].tolist()} ; bad cols: {cols_bad.numel()}/{N} "
              f"first {cols_bad[:12].tolist()}")
        print("    got[0,:8] ", out[0, :8].float().tolist())
        print("    ref[0,:8] ", ref[0, :8].tolist())
        print("    got[1,:8] ", out[1, :8].float().tolist())
        print("    ref[1,:8] ", ref[1, :8].tolist())
        r = min(M - 1, 37)
        print(f"    got[{r},256:264] ", out[r, 256:264].float().tolist() if N > 264 else None)
        print(f"    ref[{r},256:264] ", ref[r, 256:264].tolist() if N > 264 else None)
    return ok


if "gemm" in what:
    print("== GEMM ==")
    ok = gemm_case(128, 256, 64, _lib.EPI_F32, "gemm 128x256x64 f32 (one tile, one k-block)")
    ok = gemm_case(128, 256, 256, _lib.EPI_F32, "gemm 128x256x256 f32 (pipeline wrap-free)") and ok
    ok = gemm_case(128, 256, 1024, _lib.EPI_F32, "gemm 128x256x1024 f32 (ring wraps)") and ok
    ok = gemm_case(300, 512, 1024, _lib.EPI_F32, "gemm 300x512x1024 f32 (ragged M, 2 n-tiles)") and ok
    for epi, nm in [(_lib.EPI_BIAS_BF16, "bias_bf16"), (_lib.EPI_BIAS_GELU_BF16, "bias_gelu_bf16"),
                    (_lib.EPI_BIAS_QGELU_BF16, "bias_qgelu_bf16"), (_lib.EPI_BIAS_RESID_F32, "bias_resid_f32"),
                    (_lib.EPI_LEAKY_F32, "leaky_f32"), (_lib.EPI_BIAS_F32, "bias_f32")]:
        gemm_case(1370 * 2, 768 if "leaky" in nm else 1024, 1024, epi, f"gemm 2740xNx1024 {nm}")
    # big persistent case + timing
    M, N, K = 1370 * 32, 3072, 1024
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * (K ** -0.5)).bfloat16()
    bias = torch.randn(N, device=dev) * 0.1
    out = engine.gemm_bf16(a, w, _lib.EPI_BIAS_BF16, bias=bias)
    ref = (a[:4096].float() @ w.float().t()) + bias
    report("gemm 43840x3072x1024 bias_bf16 (first 4096 rows)", out[:4096], ref, 3e-2)
    ref = (a[-2000:].float() @ w.float().t()) + bias
    report("gemm 43840x3072x1024 bias_bf16 (last 2000 rows)", out[-2000:], ref, 3e-2)
    for (M, N, K, nm) in [(43840, 3072, 1024, "qkv"), (43840, 1024, 1024, "out"), (43840, 4096, 1024, "fc"),
                          (43840, 1024, 4096, "proj")]:
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        w = (torch.randn(N, K, device=dev) * (K ** -0.5)).bfloat16()
        bias = torch.randn(N, device=dev) * 0.1
        outb = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        epi = _lib.EPI_BIAS_GELU_BF16 if nm == "fc" else _lib.EPI_BIAS_BF16
        for _ in range(3):
            engine.gemm_bf16(a, w, epi, bias=bias, out_bf16=outb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            engine.gemm_bf16(a, w, epi, bias=bias, out_bf16=outb)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"    timing {nm}: {ms:.3f} ms  -> {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            torch.matmul(a, w.t())
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 10
        print(f"    (library comparator torch.matmul bf16: {ms:.3f} ms -> {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s)")

if "ln" in what:
    print("== LayerNorm ==")
    for D in (1024, 768):
        x = torch.randn(2 * 1370, D, device=dev) * 3 + 0.5
        g = torch.randn(D, device=dev) * 0.1 + 1
        b = torch.randn(D, device=dev) * 0.1
        ob, of = engine.layernorm(x, g, b, want_bf16=True, want_f32=True)
        ref = torch.nn.functional.layer_norm(x, (D,), g, b, 1e-5)
        report(f"layernorm D={D} f32", of, ref, 2e-5)
        report(f"layernorm D={D} bf16", ob, ref, 4e-2)
    x = torch.randn(3 * 1370, 1024, device=dev)
    ob, of = engine.layernorm(x, g := torch.ones(1024, device=dev), b := torch.zeros(1024, device=dev), skip_cls_P=1369,
                              want_bf16=False, want_f32=True)
    ref = torch.nn.functional.layer_norm(x.view(3, 1370, 1024)[:, 1:], (1024,), g, b, 1e-5).reshape(-1, 1024)
    report("layernorm skip-CLS rows", of, ref, 2e-5)


def attn_ref(qkv, causal):
    B, N, D3 = qkv.shape
    D = D3 // 3
    H = D // 64
    q, k, v = qkv.float().split(D, dim=-1)
    q = q.view(B, N, H, 64).transpose(1, 2)
    k = k.view(B, N, H, 64).transpose(1, 2)
    v = v.view(B, N, H, 64).transpose(1, 2)
    s = (q * 0.125) @ k.transpose(-1, -2)
    if causal:
        s = s + torch.full((N, N), float("-inf"), device=qkv.device).triu(1)
    p = torch.softmax(s, dim=-1)
    return (p @ v).transpose(1, 2).reshape(B, N, D)


if "attn" in what:
    print("== attention ==")
    for (B, N, D, causal, nm) in [(1, 128, 64, False, "1 head, 1 kv tile, 1 q tile"),
                                  (1, 256, 64, False, "1 head, 2 kv tiles, 2 q tiles"),
                                  (1, 200, 128, False, "2 heads, ragged"),
                                  (2, 1370, 1024, False, "ViT-L 518px shape"),
                                  (3, 77, 768, True, "text tower shape, causal")]:
        qkv = (torch.randn(B, N, 3 * D, device=dev)).bfloat16()
        out = engine.attention(qkv, causal=causal)
        torch.cuda.synchronize()
        ref = attn_ref(qkv, causal)
        ok, err = report(f"attention B={B} N={N} D={D} causal={causal} ({nm})", out, ref, 2e-2)
        if not ok:
            e = err.clone(); e[~torch.isfinite(e)] = 1e9
            rb = (e[0].max(dim=1).values > 2e-2).nonzero().flatten()
            cb = (e[0].max(dim=0).values > 2e-2).nonzero().flatten()
            print(f"    bad q rows {rb.numel()}/{N} first {rb[:10].tolist()} ; bad cols {cb.numel()}/{D} first {cb[:10].tolist()}")
            print("    got[0,0,:8]", out[0, 0, :8].float().tolist())
            print("    ref[0,0,:8]", ref[0, 0, :8].tolist())
            print("    got[0,-1,:8]", out[0, -1, :8].float().tolist())
            print("    ref[0,-1,:8]", ref[0, -1, :8].tolist())
    B, N, D = 32, 1370, 1024
    qkv = torch.randn(B, N, 3 * D, device=dev).bfloat16()
    for _ in range(3):
        out = engine.attention(qkv)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = engine.attention(qkv)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 4.0 * B * 16 * N * N * 64
    print(f"    timing attention B=32: {ms:.3f} ms -> {fl / ms / 1e9:.1f} TFLOP/s (algorithmic)")
    q, k, v = qkv.view(B, N, 3, 16, 64).permute(2, 0, 3, 1, 4).contiguous()
    for _ in range(2):
        torch.nn.functional.scaled_dot_product_attention(q, k, v)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        torch.nn.functional.scaled_dot_product_attention(q, k, v)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"    (library comparator torch SDPA bf16: {ms:.3f} ms -> {fl / ms / 1e9:.1f} TFLOP/s)")

if "maps" in what:
    print("== score maps ==")
    import torch.nn.functional as F
    B, G, E, S = 2, 37, 768, 518
    seg = [F.normalize(torch.randn(B, G * G, E, device=dev), dim=-1) for _ in range(4)]
    text = F.normalize(torch.randn(E, 2, device=dev), dim=0)
    det = F.normalize(torch.randn(B, G * G, E, device=dev), dim=-1).mean(1)
    maps, score = engine.score_maps(seg, text, S, test=True, ksize=7, sigma=1.0, det=det)

    def blur(x, k, sg):
        xs = torch.linspace(-(k - 1) / 2, (k - 1) / 2, k, device=dev)
        w = torch.exp(-0.5 * (xs / sg) ** 2)
        w = w / w.sum()
        k2 = (w[:, None] @ w[None, :])[None, None]
        return F.conv2d(F.pad(x, (k // 2,) * 4, mode="reflect"), k2)

    ref = 0
    for f in seg:
        s = 100.0 * f @ text
        pp = s.permute(0, 2, 1).view(B, 2, G, G)
        m = (pp[:, 1] + 1 - pp[:, 0]) / 2
        ref = ref + F.interpolate(blur(m.unsqueeze(1), 7, 1.0), size=S, mode="bilinear", align_corners=True)[:, 0]
    report("score_maps test-mode sum of 4 levels", maps, ref, 2e-4)
    report("image score", score, ((det @ text)[:, 1] + 1) / 2, 1e-5)
    m2, _ = engine.score_maps(seg[:1], text, S, test=False)
    s = 100.0 * seg[0] @ text
    pp = s.permute(0, 2, 1).view(B, 2, G, G)
    ref2 = torch.softmax(F.interpolate(pp, size=S, mode="bilinear", align_corners=True), dim=1)
    report("score_maps train-mode softmax", m2, ref2, 2e-5)

print("FAILURES:", failures)
sys.exit(1 if failures else 0)
