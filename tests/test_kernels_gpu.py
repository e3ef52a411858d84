"""GPU parity of the building-block kernels (through the C ABI) against plain torch fp32 of the same op."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DTYPES = [torch.float16, torch.bfloat16]
EPS = {torch.float16: 2 ** -11, torch.bfloat16: 2 ** -8}  # unit roundoff of one 16-bit store


@pytest.fixture(scope="module")
def eng():
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import _lib, engine
    return _lib, engine


def _ab(M, N, K, dtype):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(dtype)
    w = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).to(dtype)
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    return a, w, b


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K", [(1, 256, 64), (128, 256, 64), (129, 512, 192), (300, 768, 1024), (2740, 1024, 4096),
                                   (1369 * 3, 1024, 640)])
def test_gemm_f32_epilogues(eng, M, N, K, dtype):
    _lib, engine = eng
    a, w, b = _ab(M, N, K, dtype)
    ref = a.float() @ w.float().t()
    assert (engine.gemm16(a, w, _lib.EPI_F32) - ref).abs().max().item() < 1e-3
    assert (engine.gemm16(a, w, _lib.EPI_BIAS_F32, bias=b) - (ref + b)).abs().max().item() < 1e-3
    assert (engine.gemm16(a, w, _lib.EPI_LEAKY_F32) - F.leaky_relu(ref, 0.01)).abs().max().item() < 1e-3
    res = torch.randn(M, N, device="cuda")
    out, cp = res.clone(), torch.empty(M, N, device="cuda", dtype=dtype)
    engine.gemm16(a, w, _lib.EPI_BIAS_RESID_F32, bias=b, out_f32=out, out_16=cp)
    assert (out - (res + ref + b)).abs().max().item() < 1e-3
    assert (cp.float() - out).abs().max().item() <= out.abs().max().item() * EPS[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K", [(257, 3072, 1024), (2740, 4096, 1024)])
def test_gemm_16bit_epilogues(eng, M, N, K, dtype):
    _lib, engine = eng
    a, w, b = _ab(M, N, K, dtype)
    y = a.float() @ w.float().t() + b
    tol = lambda r: r.abs().max().item() * EPS[dtype] + 1e-3  # one 16-bit rounding of the result
    assert (engine.gemm16(a, w, _lib.EPI_BIAS_BF16, bias=b).float() - y).abs().max().item() < tol(y)
    g = F.gelu(y)
    assert (engine.gemm16(a, w, _lib.EPI_BIAS_GELU_BF16, bias=b).float() - g).abs().max().item() < tol(g)
    q = y * torch.sigmoid(1.702 * y)
    assert (engine.gemm16(a, w, _lib.EPI_BIAS_QGELU_BF16, bias=b).float() - q).abs().max().item() < tol(q)


def test_fp16_stores_saturate_instead_of_overflowing(eng):
    _lib, engine = eng
    a = torch.full((128, 64), 200.0, device="cuda", dtype=torch.float16)
    w = torch.full((256, 64), 200.0, device="cuda", dtype=torch.float16)
    out = engine.gemm16(a, w, _lib.EPI_BIAS_BF16, bias=torch.zeros(256, device="cuda"))
    assert torch.isfinite(out).all() and out.float().max().item() == 65504.0


def test_gemm_rejects_bad_shapes(eng):
    _lib, engine = eng
    a, w, b = _ab(64, 200, 64, torch.float16)
    with pytest.raises(RuntimeError, match="unsupported shape"):
        engine.gemm16(a, w, _lib.EPI_F32)
    a, w, b = _ab(64, 256, 64, torch.float16)
    with pytest.raises(RuntimeError, match="needs a bias"):
        engine.gemm16(a, w, _lib.EPI_BIAS_F32)


def _attn_ref(qkv, causal):
    B, N, D3 = qkv.shape
    D = D3 // 3
    H = D // 64
    q, k, v = qkv.float().split(D, dim=-1)
    q, k, v = (t.view(B, N, H, 64).transpose(1, 2) for t in (q, k, v))
    s = (q * 0.125) @ k.transpose(-1, -2)
    if causal:
        s = s + torch.full((N, N), float("-inf"), device=qkv.device).triu(1)
    return (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, N, D)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,N,D,causal", [(1, 1, 64, False), (1, 128, 64, False), (2, 129, 128, False),
                                          (1, 577, 1024, False), (2, 1370, 1024, False), (3, 77, 768, True),
                                          (1, 300, 128, True)])
def test_attention(eng, B, N, D, causal, dtype):
    _lib, engine = eng
    g = torch.Generator(device="cuda").manual_seed(N + D)
    qkv = (torch.randn(B, N, 3 * D, device="cuda", generator=g) * 1.5).to(dtype)
    out = engine.attention(qkv, causal=causal).float()
    ref = _attn_ref(qkv, causal)
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() < (2e-2 if dtype == torch.bfloat16 else 4e-3)


def test_attention_extreme_logits(eng):
    """Large score range: the online-softmax rescale path (running max changes across K/V tiles)."""
    _lib, engine = eng
    g = torch.Generator(device="cuda").manual_seed(3)
    qkv = torch.randn(1, 700, 3 * 64, device="cuda", generator=g)
    qkv[:, :, :128] *= 6.0  # |q.k|/8 up to ~100s
    qkv = qkv.half()
    out = engine.attention(qkv).float()
    assert torch.isfinite(out).all()
    assert (out - _attn_ref(qkv, False)).abs().max().item() < 1e-2


@pytest.mark.parametrize("dtype", DTYPES)
def test_attention_growing_logits_redo_path(eng, dtype):
    """Every K/V tile's scores exceed the previous tile's by far more than the lazy rescale threshold (2^8): the running
    maximum moves, and O in TMEM and the row sum are rescaled, on every tile (attention_tcgen05.cuh)."""
    _lib, engine = eng
    g = torch.Generator(device="cuda").manual_seed(11)
    N = 1370
    qkv = torch.randn(2, N, 3 * 128, device="cuda", generator=g)
    ramp = 1.0 + 4.0 * (torch.arange(N, device="cuda") // 128).float()   # key norm grows tile by tile
    qkv[:, :, 128:256] *= ramp[None, :, None]
    qkv[:, :, :128] *= 3.0
    qkv = qkv.to(dtype)
    out = engine.attention(qkv).float()
    assert torch.isfinite(out).all()
    assert (out - _attn_ref(qkv, False)).abs().max().item() < (3e-2 if dtype == torch.bfloat16 else 1e-2)


def test_attention_persistent_schedule(eng):
    """The kernel is persistent: CTAs draw (image, head, query block) items from a counter slot that every launch must
    leave zeroed. (1) more items than SMs and fewer items than SMs give the same rows; (2) launches in flight on two
    streams at once use different slots; (3) more launches than there are slots (256) wrap around cleanly."""
    _lib, engine = eng
    g = torch.Generator(device="cuda").manual_seed(5)
    big = (torch.randn(12, 700, 3 * 1024, device="cuda", generator=g) * 1.5).half()       # 12*16*3 = 576 items
    ref = engine.attention(big).clone()
    for i in (0, 7, 11):                                                                     # 48 items: fewer than SMs
        assert torch.equal(engine.attention(big[i:i + 1])[0], ref[i])
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    other = (torch.randn(12, 700, 3 * 1024, device="cuda", generator=g) * 1.5).half()
    ref_other = engine.attention(other).clone()
    torch.cuda.synchronize()
    outs = []
    for _ in range(4):
        with torch.cuda.stream(s1):
            a = engine.attention(big)
        with torch.cuda.stream(s2):
            b = engine.attention(other)
        outs.append((a, b))
    torch.cuda.synchronize()
    for a, b in outs:
        assert torch.equal(a, ref) and torch.equal(b, ref_other)
    # the two operand dtypes are separate kernel instantiations but draw their slots from ONE counter
    other_bf = other.bfloat16()
    ref_bf = engine.attention(other_bf).clone()
    torch.cuda.synchronize()
    outs = []
    for _ in range(4):
        with torch.cuda.stream(s1):
            a = engine.attention(big)
        with torch.cuda.stream(s2):
            b = engine.attention(other_bf)
        outs.append((a, b))
    torch.cuda.synchronize()
    for a, b in outs:
        assert torch.equal(a, ref) and torch.equal(b, ref_bf)
    small = (torch.randn(1, 130, 3 * 64, device="cuda", generator=g)).half()
    first = engine.attention(small).clone()
    for _ in range(600):
        out = engine.attention(small)
    assert torch.equal(out, first)
    assert torch.equal(engine.attention(big), ref)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("D", [768, 1024])
def test_layernorm(eng, D, dtype):
    _lib, engine = eng
    x = torch.randn(3 * 50, D, device="cuda") * 3 + 0.5
    g = torch.randn(D, device="cuda") * 0.1 + 1
    b = torch.randn(D, device="cuda") * 0.1
    ob, of = engine.layernorm(x, g, b, want_16=True, want_f32=True, dtype=dtype)
    ref = F.layer_norm(x, (D,), g, b, 1e-5)
    assert (of - ref).abs().max().item() < 1e-5
    assert (ob.float() - ref).abs().max().item() < ref.abs().max().item() * EPS[dtype]
    _, of2 = engine.layernorm(x, g, b, skip_cls_P=49, want_16=False, want_f32=True)
    assert torch.equal(of2, of.view(3, 50, D)[:, 1:].reshape(-1, D))


def test_score_maps_against_oracle(eng):
    _lib, engine = eng
    B, G, E, S = 2, 37, 768, 518
    seg = [F.normalize(torch.randn(B, G * G, E, device="cuda"), dim=-1) for _ in range(4)]
    text = F.normalize(torch.randn(E, 2, device="cuda"), dim=0)
    det = F.normalize(torch.randn(B, G * G, E, device="cuda"), dim=-1).mean(1)
    from oracle import aaclip_oracle as orc
    for dom, (k, sg) in {"Industrial": (7, 1.0), "Medical": (9, 1.5)}.items():
        maps, score = engine.score_maps(seg, text, S, test=True, ksize=k, sigma=sg, det=det)
        ref = orc.anomaly_maps([s.cpu() for s in seg], text.cpu(), S, dom)
        assert (maps.cpu() - ref).abs().max().item() < 1e-4
        assert (score.cpu() - orc.image_score(det.cpu(), text.cpu())).abs().max().item() < 1e-5
    tr, _ = engine.score_maps(seg[:1], text, S, test=False)
    ref = orc.calculate_similarity_map(seg[0].cpu(), text.cpu(), S, test=False)
    assert (tr.cpu() - ref).abs().max().item() < 1e-5
    with pytest.raises(RuntimeError, match="exactly one level"):
        engine.score_maps(seg, text, S, test=False)
    # ragged / tiny grids and odd sizes
    seg2 = [F.normalize(torch.randn(1, 16 * 16, E, device="cuda"), dim=-1)]
    m, _ = engine.score_maps(seg2, text, 225, test=True, ksize=9, sigma=1.5)
    ref = orc.anomaly_maps([seg2[0].cpu()], text.cpu(), 225, "Medical")
    assert (m.cpu() - ref).abs().max().item() < 1e-4


def test_score_maps_raw_normalises_in_the_similarity_pass(eng):
    """aaclip_score_maps_raw on un-normalised projections == aaclip_score_maps on F.normalize of them (the oracle's
    order of operations, model/adapter.py:182 then forward_utils.py:199), including an all-zero row (eps 1e-12)."""
    _lib, engine = eng
    B, G, E, S = 2, 24, 768, 336
    raw = [torch.randn(B, G * G, E, device="cuda") * (0.3 + l) for l in range(4)]
    raw[1][0, 5] = 0.0
    text = F.normalize(torch.randn(B, E, 2, device="cuda"), dim=1)
    det = F.normalize(torch.randn(B, G * G, E, device="cuda"), dim=-1).mean(1)
    from oracle import aaclip_oracle as orc
    a, sa = engine.score_maps(raw, text, S, test=True, ksize=7, sigma=1.0, det=det, raw=True)
    ref = orc.anomaly_maps([F.normalize(r, dim=-1).cpu() for r in raw], text.cpu(), S, "Industrial")
    assert torch.isfinite(a).all()
    assert (a.cpu() - ref).abs().max().item() < 1e-4
    b, sb = engine.score_maps([F.normalize(r, dim=-1) for r in raw], text, S, test=True, ksize=7, sigma=1.0, det=det)
    assert (a - b).abs().max().item() < 5e-5 and torch.equal(sa, sb)


@pytest.mark.parametrize("dtype", DTYPES)
def test_hilo_form_of_the_residual_stream(eng, dtype):
    """x -> (hi, lo) -> x: hi is x rounded to the operand dtype (what the GEMMs read), hi + lo reproduces x to the
    square of the format's unit roundoff; exact for values the format holds."""
    _lib, engine = eng
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(1000, 1024, device="cuda", generator=g) * 3.0
    x[::17, 3] = 275.0 + x[::17, 3]                     # massive-activation sized entries
    x[5, :8] = torch.tensor([0.0, 1.0, -2.5, 65504.0, 1e-6, -1e-6, 3.0e-5, 0.125], device="cuda")
    hi, lo = engine.split_hilo(x, dtype)
    assert torch.equal(hi, x.to(dtype))
    back = engine.join_hilo(hi, lo)
    # |hi + lo - x| <= eps^2 |x| (two roundings), with the absolute floor of the format's smallest subnormal
    floor = 2.0 ** -24 if dtype == torch.float16 else 1e-38
    assert ((back - x).abs() <= 4 * EPS[dtype] ** 2 * x.abs() + floor).all()
    exact = torch.randn(64, 1024, device="cuda", generator=g).to(dtype).float()
    h2, l2 = engine.split_hilo(exact, dtype)
    assert torch.equal(engine.join_hilo(h2, l2), exact) and int(l2.float().abs().max().item()) == 0


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K", [(300, 1024, 1024), (2740, 1024, 4096), (77 * 3, 768, 768)])
def test_residual_gemm_on_hilo_stream(eng, M, N, K, dtype):
    """aaclip_gemm16_resid_hl (what the tower launches for out_proj / c_proj): x += A W^T + b on the (hi, lo) planes,
    against fp32 torch on the same 16-bit operands; hi must be exactly the rounded new x; the LayerNorm partials must
    describe the new rows."""
    _lib, engine = eng
    a, w, b = _ab(M, N, K, dtype)
    g = torch.Generator(device="cuda").manual_seed(M + K)
    x = torch.randn(M, N, device="cuda", generator=g) * 2.0 + 0.5
    hi, lo = engine.split_hilo(x, dtype)
    x_in = engine.join_hilo(hi, lo)
    ref = x_in + a.float() @ w.float().t() + b
    stats = engine.gemm16_resid_hl(a, w, b, hi, lo)
    out = engine.join_hilo(hi, lo)
    assert (out - ref).abs().max().item() < 1e-3 + 8 * EPS[dtype] ** 2 * ref.abs().max().item()
    assert (hi.float() - ref).abs().max().item() <= 1.01 * EPS[dtype] * ref.abs().max().item() + 1e-3
    cols = _lib.load().aaclip_stat_columns()
    assert stats.shape == (M, N // cols, 2)
    assert torch.allclose(stats[..., 0].sum(1), ref.sum(1), atol=5e-3, rtol=1e-5)
    assert torch.allclose(stats[..., 1].sum(1), (ref * ref).sum(1), atol=5e-2, rtol=1e-5)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,D", [(300, 1024), (2740, 1024), (77 * 3, 768)])
def test_adapter_gemm_and_blend(eng, M, D, dtype):
    """aaclip_adapter (GEMM with the LeakyReLU + 16-bit store + sum-of-squares epilogue, then the blend kernel on the
    (hi, lo) form of the stream) against model/adapter.py:163-170 in fp32 on the same 16-bit operands; also the
    outputs the next block consumes: the hi plane = the rounded new x, and its LayerNorm partials."""
    _lib, engine = eng
    g = torch.Generator(device="cuda").manual_seed(M + D)
    x = torch.randn(M, D, device="cuda", generator=g) * 1.3 + 0.2
    w = (torch.randn(D, D, device="cuda", generator=g) * D ** -0.5).to(dtype)
    hi, lo = engine.split_hilo(x, dtype)
    x = engine.join_hilo(hi, lo)                          # the value the stream actually holds
    a = F.leaky_relu(hi.float() @ w.float().t(), 0.01)
    a = a * x.norm(dim=-1, keepdim=True) / a.norm(dim=-1, keepdim=True)
    ref = 0.1 * a + 0.9 * x
    hi_in, lo_in = hi.clone(), lo.clone()
    stats = engine.adapter(hi, lo, w, 0.1)
    xo = engine.join_hilo(hi, lo)
    tol = 4 * EPS[dtype] * 0.1 * a.abs().max().item() + 8 * EPS[dtype] ** 2 * ref.abs().max().item() + 1e-5
    assert (xo - ref).abs().max().item() < tol          # a is rounded once to 16 bits, then weighted by 0.1
    assert (hi.float() - xo).abs().max().item() <= 1.01 * EPS[dtype] * xo.abs().max().item()
    cols = _lib.load().aaclip_stat_columns()
    assert stats.shape == (M, D // cols, 2)
    assert torch.allclose(stats[..., 0].sum(1), xo.sum(1), atol=5e-3, rtol=1e-4)
    assert torch.allclose(stats[..., 1].sum(1), (xo * xo).sum(1), atol=5e-2, rtol=1e-4)
    # row-local: a row alone gives the same bits as the row inside the batch
    h1, l1 = hi_in[7:8].clone(), lo_in[7:8].clone()
    engine.adapter(h1, l1, w, 0.1)
    assert torch.equal(h1[0], hi[7]) and torch.equal(l1[0], lo[7])


# ---------------------------------------------------------------------------------------------------------
# LayerNorm folded into the GEMMs that produce / consume the residual stream (model/transformer.py:245-254)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,W,N", [(300, 1024, 3072), (2740, 1024, 4096), (77 * 3, 768, 2304)])
def test_folded_layernorm_gemm_chain(eng, M, W, N, dtype):
    """resid GEMM (x += a w^T + b, emits raw 16-bit x and row partials) followed by the folded-LN GEMM must equal
    Linear(LayerNorm(x)) computed in fp32 torch on the same x -- to one 16-bit rounding of the operands."""
    _lib, engine = eng
    g = torch.Generator(device="cuda").manual_seed(M + W + N)
    a, w_o, b_o = _ab(M, W, W, dtype)
    x = torch.randn(M, W, device="cuda", generator=g) * 2.0 + 0.3          # non-zero row mean on purpose
    x_ref = x + a.float() @ w_o.float().t() + b_o
    x16, stats = engine.gemm16_resid(a, w_o, b_o, x)
    assert (x - x_ref).abs().max().item() < 1e-3
    assert (x16.float() - x).abs().max().item() <= x.abs().max().item() * EPS[dtype]
    # the partials: (sum, sum of squares) over each 128-column slice of the new x
    sc = _lib.load().aaclip_stat_columns()
    want = torch.stack([x.view(M, W // sc, sc).sum(-1), (x * x).view(M, W // sc, sc).sum(-1)], dim=-1)
    assert torch.allclose(stats, want, rtol=1e-5, atol=1e-3)
    # fold + consume
    gamma = torch.randn(W, device="cuda", generator=g) * 0.1 + 1.0
    beta = torch.randn(W, device="cuda", generator=g) * 0.05
    w = torch.randn(N, W, device="cuda", generator=g) * W ** -0.5
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    w16, colsum, bias_f = engine.ln_fold(w, b, gamma, beta, dtype)
    assert torch.allclose(w16.float(), (w * gamma).to(dtype).float())
    assert torch.allclose(colsum, w16.float().sum(-1), atol=1e-3)
    assert torch.allclose(bias_f, b + w @ beta, atol=1e-4)
    y_ref = F.layer_norm(x, (W,), gamma, beta, 1e-5) @ w.t() + b
    tol = y_ref.abs().max().item() * 3 * EPS[dtype] + 2e-3
    y = engine.gemm16_ln(x16, w16, _lib.EPI_LN_BIAS_BF16, bias_f, colsum, stats, W)
    assert (y.float() - y_ref).abs().max().item() < tol
    yg = engine.gemm16_ln(x16, w16, _lib.EPI_LN_BIAS_GELU_BF16, bias_f, colsum, stats, W)
    assert (yg.float() - F.gelu(y_ref)).abs().max().item() < tol
    yq = engine.gemm16_ln(x16, w16, _lib.EPI_LN_BIAS_QGELU_BF16, bias_f, colsum, stats, W)
    assert (yq.float() - y_ref * torch.sigmoid(1.702 * y_ref)).abs().max().item() < tol


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,P,W,N,leaky", [(2, 256, 1024, 768, False), (3, 576, 1024, 768, True), (1, 1369, 1024, 512, False)])
def test_token_tap_projection_with_folded_ln_post(eng, B, P, W, N, leaky, dtype):
    """aaclip_gemm16_ln_tok against model/adapter.py:171-184 in fp32: x[1:] of every image -> ln_post -> SimpleProj
    (optionally LeakyReLU). The GEMM reads all P+1 rows per image from the raw 16-bit stream; the CLS rows must not
    appear in the output and every patch row must land at b*P + t - 1."""
    _lib, engine = eng
    g = torch.Generator(device="cuda").manual_seed(B * 31 + P)
    M = B * (P + 1)
    x = torch.randn(M, W, device="cuda", generator=g) * 2.0 + 0.3
    x[::7, 5] += 40.0                                          # an outlier channel on some rows
    gamma = 1.0 + 0.1 * torch.randn(W, device="cuda", generator=g)
    beta = 0.1 * torch.randn(W, device="cuda", generator=g)
    w = torch.randn(N, W, device="cuda", generator=g) * W ** -0.5
    cols = _lib.load().aaclip_stat_columns()
    stats = torch.stack([x.view(M, W // cols, cols).sum(-1), (x * x).view(M, W // cols, cols).sum(-1)], -1).contiguous()
    w16, cs, bf = engine.ln_fold(w, torch.zeros(N, device="cuda"), gamma, beta, dtype)
    out = engine.gemm16_ln_tok(x.to(dtype), w16, bf, cs, stats, W, P, leaky=leaky)
    ref = F.layer_norm(x, (W,), gamma, beta, 1e-5).view(B, P + 1, W)[:, 1:].reshape(B * P, W) @ w.t()
    if leaky:
        ref = F.leaky_relu(ref, 0.01)
    assert out.shape == (B * P, N)
    # operand rounding: x and W' are rounded once to 16 bits, K = 1024 products of O(1) x O(W^-0.5) terms
    tol = {torch.float16: 2e-2, torch.bfloat16: 1.5e-1}[dtype]
    assert (out - ref).abs().max().item() < tol
    assert ((out - ref).abs().mean() / ref.abs().mean()).item() < 8 * EPS[dtype]


def test_folded_layernorm_is_row_local(eng):
    """Batch invariance: a row's folded-LN output must not depend on which other rows are in the launch."""
    _lib, engine = eng
    g = torch.Generator(device="cuda").manual_seed(5)
    W, N = 1024, 3072
    x = torch.randn(600, W, device="cuda", generator=g)
    x16 = x.half()
    stats = torch.stack([x16.float().view(600, 8, 128).sum(-1), (x16.float() ** 2).view(600, 8, 128).sum(-1)], -1).contiguous()   # any partition of the row is a valid set of partials
    w16, colsum, bias_f = engine.ln_fold(torch.randn(N, W, device="cuda", generator=g) * 0.03,
                                         torch.zeros(N, device="cuda"), torch.ones(W, device="cuda"),
                                         torch.zeros(W, device="cuda"))
    full = engine.gemm16_ln(x16, w16, _lib.EPI_LN_BIAS_BF16, bias_f, colsum, stats, W)
    part = engine.gemm16_ln(x16[130:391].contiguous(), w16, _lib.EPI_LN_BIAS_BF16, bias_f, colsum,
                            stats[130:391].contiguous(), W)
    assert torch.equal(full[130:391], part)
