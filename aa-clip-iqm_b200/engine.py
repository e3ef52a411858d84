"""Torch-tensor front end of the C ABI (include/aaclip.h). PyTorch is used only for device memory and streams.

Everything here requires a CUDA (sm_100) device and the in-tree libaaclip_b200.so; nothing falls back to PyTorch
ops or to the CPU.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*tensors: torch.Tensor):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("aaclip_b200: tensors must live on a CUDA (sm_100a) device; there is no CPU path")


def _ptr_array(tensors: Sequence[Optional[torch.Tensor]]):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = 0 if t is None else t.data_ptr()
    return arr


class StreamGuard:
    """Orders uses of a buffer that one object shares between calls (an engine's workspace, a predictor's staging
    input): a use on a different CUDA stream than the previous one first waits for the event recorded after that use.
    Same-stream use costs one event record. It makes interleaved calls from several streams SAFE (serialised on the
    buffer); it does not make a handle thread-safe."""

    def __init__(self):
        self._event = None
        self._stream = None

    def enter(self):
        if torch.cuda.is_current_stream_capturing():
            return                      # inside a CUDA-graph capture the graph's own edges order the uses
        cur = torch.cuda.current_stream()
        if self._event is not None and self._stream != cur:
            cur.wait_event(self._event)

    def leave(self):
        if torch.cuda.is_current_stream_capturing():
            return
        cur = torch.cuda.current_stream()
        if self._event is None:
            self._event = torch.cuda.Event()
        self._event.record(cur)
        self._stream = cur


# ---------------------------------------------------------------------------------------- building blocks
_TORCH_DT = {torch.float16: _lib.DT_F16, torch.bfloat16: _lib.DT_BF16}


def _dt(t_or_dtype) -> int:
    d = t_or_dtype.dtype if isinstance(t_or_dtype, torch.Tensor) else t_or_dtype
    if d not in _TORCH_DT:
        raise TypeError(f"16-bit operand dtype must be torch.float16 or torch.bfloat16, got {d}")
    return _TORCH_DT[d]


def pack16(src: torch.Tensor, dtype: torch.dtype = torch.float16, padded_cols: Optional[int] = None) -> torch.Tensor:
    """fp32 [R,C] -> fp16|bf16 [R,Cp] zero padded, saturating (aaclip_pack16)."""
    _need_cuda(src)
    src = src.contiguous().float()
    R, C = src.shape
    Cp = C if padded_cols is None else padded_cols
    dst = torch.empty((R, Cp), dtype=dtype, device=src.device)
    _lib.check(_lib.load().aaclip_pack16(_dt(dtype), _ptr(src), _ptr(dst), R, C, Cp, _stream()), "aaclip_pack16")
    return dst


def gemm16(a: torch.Tensor, w: torch.Tensor, epi: int, bias: Optional[torch.Tensor] = None,
           out_f32: Optional[torch.Tensor] = None, out_16: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C = A[M,K] @ W[N,K]^T with fused epilogue `epi` (aaclip_gemm16). Returns the primary output."""
    _need_cuda(a, w)
    assert a.dtype == w.dtype and a.is_contiguous() and w.is_contiguous()
    out_bf16 = out_16
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    bf16_out = epi in (_lib.EPI_BIAS_BF16, _lib.EPI_BIAS_GELU_BF16, _lib.EPI_BIAS_QGELU_BF16)
    if bf16_out and out_bf16 is None:
        out_bf16 = torch.empty((M, N), dtype=a.dtype, device=a.device)
    if not bf16_out and out_f32 is None:
        out_f32 = torch.empty((M, N), dtype=torch.float32, device=a.device)
    _lib.check(
        _lib.load().aaclip_gemm16(_dt(a), _ptr(a), _ptr(w), M, N, K, epi, _ptr(bias), _ptr(out_f32), _ptr(out_bf16), N,
                                  _stream()),
        "aaclip_gemm16")
    return out_bf16 if bf16_out else out_f32


def ln_fold(w: torch.Tensor, b: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
            dtype: torch.dtype = torch.float16) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(W' = W*gamma rounded to `dtype`, colsum(W'), b + W beta): LayerNorm folded into a Linear (aaclip_ln_fold)."""
    _need_cuda(w, b, gamma, beta)
    w = w.contiguous().float()
    N, K = w.shape
    w16 = torch.empty((N, K), dtype=dtype, device=w.device)
    colsum = torch.empty((N,), dtype=torch.float32, device=w.device)
    bias_f = torch.empty((N,), dtype=torch.float32, device=w.device)
    _lib.check(_lib.load().aaclip_ln_fold(_dt(dtype), _ptr(w), _ptr(b.contiguous().float()), _ptr(gamma.contiguous().float()),
                                          _ptr(beta.contiguous().float()), N, K, _ptr(w16), _ptr(colsum), _ptr(bias_f),
                                          _stream()), "aaclip_ln_fold")
    return w16, colsum, bias_f


def gemm16_ln(a: torch.Tensor, w16: torch.Tensor, epi: int, bias_f: torch.Tensor, colsum: torch.Tensor,
              stats: torch.Tensor, ln_width: int) -> torch.Tensor:
    """act(LayerNorm(rows of x) @ W^T + b) from the raw 16-bit rows `a`, folded weights and row partials `stats`
    [M, n_partials, 2] (aaclip_gemm16_ln)."""
    _need_cuda(a, w16, bias_f, colsum, stats)
    assert a.dtype == w16.dtype and a.is_contiguous() and w16.is_contiguous() and stats.is_contiguous()
    M, K = a.shape
    N = w16.shape[0]
    out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    _lib.check(_lib.load().aaclip_gemm16_ln(_dt(a), _ptr(a), _ptr(w16), M, N, K, epi, _ptr(bias_f), _ptr(colsum),
                                            _ptr(stats), stats.shape[1], ln_width, _ptr(out), N, _stream()),
               "aaclip_gemm16_ln")
    return out



def gemm16_ln_tok(a: torch.Tensor, w16: torch.Tensor, bias_f: torch.Tensor, colsum: torch.Tensor, stats: torch.Tensor,
                  ln_width: int, patches: int, leaky: bool = False) -> torch.Tensor:
    """Token-tap projection with ln_post folded in (aaclip_gemm16_ln_tok): `a` holds all patches+1 raw 16-bit rows of
    every image; returns [B*patches, N] fp32 = act(LayerNorm(x)[1:] @ W^T) per image (CLS rows dropped)."""
    _need_cuda(a, w16, bias_f, colsum, stats)
    assert a.dtype == w16.dtype and a.is_contiguous() and w16.is_contiguous() and stats.is_contiguous()
    M, K = a.shape
    N = w16.shape[0]
    out = torch.empty((M // (patches + 1) * patches, N), dtype=torch.float32, device=a.device)
    _lib.check(_lib.load().aaclip_gemm16_ln_tok(_dt(a), _ptr(a), _ptr(w16), M, N, K, int(leaky), _ptr(bias_f), _ptr(colsum),
                                                _ptr(stats), stats.shape[1], ln_width, patches, _ptr(out), N, _stream()),
               "aaclip_gemm16_ln_tok")
    return out


def gemm16_resid(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, x: torch.Tensor, want_x16: bool = True,
                 want_stats: bool = True):
    """x += a @ w^T + bias in place (fp32); returns (x16 | None, stats [M, N/stat_columns, 2] | None)
    (aaclip_gemm16_resid)."""
    _need_cuda(a, w, bias, x)
    assert a.dtype == w.dtype and a.is_contiguous() and w.is_contiguous() and x.is_contiguous() and x.dtype == torch.float32
    M, K = a.shape
    N = w.shape[0]
    x16 = torch.empty((M, N), dtype=a.dtype, device=a.device) if want_x16 else None
    sc = _lib.load().aaclip_stat_columns()
    stats = torch.empty((M, N // sc, 2), dtype=torch.float32, device=a.device) if want_stats else None
    _lib.check(_lib.load().aaclip_gemm16_resid(_dt(a), _ptr(a), _ptr(w), M, N, K, _ptr(bias), _ptr(x), _ptr(x16),
                                               _ptr(stats), N, _stream()), "aaclip_gemm16_resid")
    return x16, stats


def attention(qkv: torch.Tensor, causal: bool = False) -> torch.Tensor:
    """qkv fp16|bf16 [B,N,3D] -> same dtype [B,N,D] (aaclip_attention)."""
    _need_cuda(qkv)
    assert qkv.is_contiguous() and qkv.dim() == 3
    B, N, D3 = qkv.shape
    D = D3 // 3
    out = torch.empty((B, N, D), dtype=qkv.dtype, device=qkv.device)
    _lib.check(_lib.load().aaclip_attention(_dt(qkv), _ptr(qkv), _ptr(out), B, N, D, int(causal), _stream()),
               "aaclip_attention")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, skip_cls_P: int = 0,
              want_16: bool = True, want_f32: bool = False, dtype: torch.dtype = torch.float16) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """LayerNorm(eps=1e-5) over the last dim of fp32 x [rows, D] (aaclip_layernorm)."""
    _need_cuda(x, gamma, beta)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2
    rows_in, D = x.shape
    rows = rows_in if skip_cls_P == 0 else rows_in // (skip_cls_P + 1) * skip_cls_P
    ob = torch.empty((rows, D), dtype=dtype, device=x.device) if want_16 else None
    of = torch.empty((rows, D), dtype=torch.float32, device=x.device) if want_f32 else None
    _lib.check(_lib.load().aaclip_layernorm(_dt(dtype), _ptr(x), _ptr(gamma), _ptr(beta), _ptr(ob), _ptr(of), rows, D,
                                            skip_cls_P, _stream()), "aaclip_layernorm")
    return ob, of


def score_maps(seg: Sequence[torch.Tensor], text: torch.Tensor, img_size: int, test: bool = True, ksize: int = 7,
               sigma: float = 1.0, det: Optional[torch.Tensor] = None, raw: bool = False
               ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Similarity -> (blur) -> bilinear upsample (-> level sum | softmax) and optional image score (aaclip_score_maps).

    seg: list of [B,P,E] fp32 unit-norm patch features; text: [E,2] or [B,E,2].
    raw=True: seg holds the un-normalised projections of a tower with ``seg_normalize(False)``; the similarity kernel
    applies F.normalize itself (aaclip_score_maps_raw).
    test=True  -> maps [B,S,S], summed over the given levels; test=False -> one level, maps [B,2,S,S].
    """
    _need_cuda(text, *seg)
    seg = [s.contiguous().float() for s in seg]
    text = text.contiguous().float()
    B, P, E = seg[0].shape
    G = int(round(P ** 0.5))
    if G * G != P:
        raise ValueError(f"patch count {P} is not a perfect square")
    per_image = int(text.dim() == 3)
    if text.shape[-2:] != (E, 2) or (per_image and text.shape[0] != B):
        raise ValueError(f"text anchors must be [E,2] or [B,E,2], got {tuple(text.shape)}")
    L = len(seg)
    lib = _lib.load()
    ws_bytes = lib.aaclip_score_workspace_bytes(B, L, P)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=text.device)
    maps = torch.empty((B, img_size, img_size) if test else (B, 2, img_size, img_size), dtype=torch.float32,
                       device=text.device)
    score = None
    if det is not None:
        det = det.contiguous().float()
        score = torch.empty((B,), dtype=torch.float32, device=text.device)
    fn = lib.aaclip_score_maps_raw if raw else lib.aaclip_score_maps
    _lib.check(
        fn(_ptr_array(seg), L, _ptr(text), per_image, B, G, E, img_size, int(test), ksize,
           float(sigma), _ptr(ws), ws_bytes, _ptr(maps), _ptr(det), _ptr(score), _stream()),
        "aaclip_score_maps")
    return maps, score


def split_hilo(x: torch.Tensor, dtype: torch.dtype = torch.float16) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 -> the (hi, lo) form of the residual stream: hi = x rounded to `dtype`, lo = the rounded remainder
    (aaclip_split_hilo)."""
    _need_cuda(x)
    x = x.contiguous().float()
    hi = torch.empty(x.shape, dtype=dtype, device=x.device)
    lo = torch.empty(x.shape, dtype=dtype, device=x.device)
    _lib.check(_lib.load().aaclip_split_hilo(_dt(dtype), _ptr(x), _ptr(hi), _ptr(lo), x.numel(), _stream()),
               "aaclip_split_hilo")
    return hi, lo


def join_hilo(hi: torch.Tensor, lo: torch.Tensor) -> torch.Tensor:
    """(hi, lo) -> fp32 (aaclip_join_hilo)."""
    _need_cuda(hi, lo)
    assert hi.dtype == lo.dtype and hi.shape == lo.shape and hi.is_contiguous() and lo.is_contiguous()
    x = torch.empty(hi.shape, dtype=torch.float32, device=hi.device)
    _lib.check(_lib.load().aaclip_join_hilo(_dt(hi), _ptr(hi), _ptr(lo), _ptr(x), hi.numel(), _stream()),
               "aaclip_join_hilo")
    return x


def gemm16_resid_hl(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, x_hi: torch.Tensor, x_lo: torch.Tensor,
                    want_stats: bool = True) -> Optional[torch.Tensor]:
    """x (hi / lo planes, in place) += A W^T + bias (aaclip_gemm16_resid_hl). Returns the LayerNorm partials
    [M, N/stat_columns, 2] of the new rows (or None)."""
    _need_cuda(a, w, bias, x_hi, x_lo)
    assert a.dtype == w.dtype == x_hi.dtype == x_lo.dtype and a.is_contiguous() and w.is_contiguous()
    assert x_hi.is_contiguous() and x_lo.is_contiguous()
    M, K = a.shape
    N = w.shape[0]
    lib = _lib.load()
    stats = torch.empty((M, N // lib.aaclip_stat_columns(), 2), dtype=torch.float32, device=a.device) if want_stats else None
    _lib.check(lib.aaclip_gemm16_resid_hl(_dt(a), _ptr(a), _ptr(w), M, N, K, _ptr(bias.contiguous().float()), _ptr(x_hi),
                                          _ptr(x_lo), _ptr(stats), N, _stream()), "aaclip_gemm16_resid_hl")
    return stats


def adapter(x_hi: torch.Tensor, x_lo: torch.Tensor, w16: torch.Tensor, weight: float) -> torch.Tensor:
    """Residual adapter + blend of one block in place (aaclip_adapter) on the (hi, lo) form of the stream; w16 [D,D]
    packed fc.0.weight. Returns the LayerNorm partials [M, D/stat_columns, 2] of the new x."""
    _need_cuda(x_hi, x_lo, w16)
    assert x_hi.dtype == x_lo.dtype == w16.dtype and x_hi.is_contiguous() and x_lo.is_contiguous()
    M, D = x_hi.shape
    lib = _lib.load()
    np_ = D // lib.aaclip_stat_columns()
    stats = torch.empty((M, np_, 2), dtype=torch.float32, device=x_hi.device)
    ws_bytes = M * D * 2 + M * np_ * 8
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x_hi.device)
    _lib.check(lib.aaclip_adapter(_dt(x_hi), _ptr(x_hi), _ptr(x_lo), _ptr(w16.contiguous()), M, D, float(weight), _ptr(ws),
                                  ws_bytes, _ptr(stats), _stream()), "aaclip_adapter")
    return stats


# ---------------------------------------------------------------------------------------- the tower
DEFAULT_CFG = dict(img_size=518, patch=14, width=1024, heads=16, layers=24, mlp_dim=4096, embed_dim=768,
                   adapt_until=6, levels=(6, 12, 18, 24), act=_lib.ACT_GELU_ERF, proj_relu=0, adapt_weight=0.1,
                   max_batch=64, operand_dtype=_lib.DT_F16, query_dim=0)


class AaclipEngine:
    """Owns one AaclipHandle: packed weights on the device + the forward entry point."""

    def __init__(self, device: torch.device | str = "cuda", **cfg):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("AaclipEngine needs a CUDA (sm_100a) device; there is no CPU path")
        full = dict(DEFAULT_CFG)
        full.update(cfg)
        self.cfg = full
        c = _lib.AaclipCfg()
        for k in ("img_size", "patch", "width", "heads", "layers", "mlp_dim", "embed_dim", "adapt_until", "act",
                  "proj_relu", "max_batch", "operand_dtype", "query_dim"):
            setattr(c, k, int(full[k]))
        c.adapt_weight = float(full["adapt_weight"])
        levels = list(full["levels"])
        c.n_levels = len(levels)
        for i, lv in enumerate(levels):
            c.levels[i] = int(lv)
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.aaclip_create(ctypes.byref(c), ctypes.byref(self._h)), "aaclip_create")
        self.G = full["img_size"] // full["patch"]
        self.P = self.G * self.G
        self.N = self.P + 1
        self.n_levels = len(levels)
        self._ws: Optional[torch.Tensor] = None
        self._ws_guard = StreamGuard()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.aaclip_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- weights
    def set_weight(self, name: str, t: torch.Tensor):
        t = t.detach().to(self.device, torch.float32).contiguous()
        shape = (ctypes.c_int64 * t.dim())(*t.shape)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.aaclip_set_weight(self._h, name.encode(), _ptr(t), shape, t.dim(), _stream()),
                       f"aaclip_set_weight({name})")
            # The engine's copy / pack kernel is enqueued on the current stream, the stream a temporary `t` was
            # allocated on: PyTorch's caching allocator is stream-ordered, so releasing `t` here cannot hand its
            # memory to anything that runs before that copy. No per-tensor synchronisation (there are ~400 tensors).

    def load_state_dict(self, sd: Dict[str, torch.Tensor], prefixes=("image_encoder.", "image_adapter.")):
        """Feed every tensor of an AdaptedCLIP state_dict whose key starts with one of `prefixes` (plus the IQM
        branch's ``query_adapters.*`` when the engine was created with query_dim > 0)."""
        n = 0
        if self.cfg["query_dim"] > 0:
            prefixes = tuple(prefixes) + ("query_adapters.",)
        for k, v in sd.items():
            if k.startswith(prefixes) and not k.endswith("proj") and ".attn_pool" not in k:
                self.set_weight(k, v)
                n += 1
        with torch.cuda.device(self.device):
            if self.missing == 0:
                # load-time work (LayerNorm folding) now, so that forward never allocates or blocks (aaclip.h)
                _lib.check(self._lib.aaclip_finalize_weights(self._h, _stream()), "aaclip_finalize_weights")
            torch.cuda.current_stream().synchronize()      # one synchronisation per load: errors surface here
        return n

    # -- 16-bit range validation (aaclip_saturation_check)
    def saturation_check(self, on: bool = True):
        """While on, every forward also counts stored 16-bit activations that hit the format's ceiling."""
        _lib.check(self._lib.aaclip_saturation_check(self._h, int(on)), "aaclip_saturation_check")

    def saturation_count(self, reset: bool = True) -> int:
        v = ctypes.c_uint64(0)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.aaclip_saturation_count(self._h, ctypes.byref(v), int(reset), _stream()),
                       "aaclip_saturation_count")
        return int(v.value)

    @property
    def missing(self) -> int:
        return self._lib.aaclip_missing_weights(self._h)

    def workspace_bytes(self, B: int) -> int:
        return self._lib.aaclip_workspace_bytes(self._h, B)

    # -- forward
    def forward(self, images: torch.Tensor, want_tokens: bool = False, want_cls: bool = False,
                want_query: bool = False):
        """images [B,3,S,S] fp32 on the engine's device -> (seg list of [B,P,E], det [B,E], tokens|None, cls|None), and
        with want_query a fifth element: the IQM query-adapter projections, a list of [B,P,query_dim] fp32."""
        _need_cuda(images)
        images = images.contiguous().float()
        B = images.shape[0]
        S = self.cfg["img_size"]
        if tuple(images.shape[1:]) != (3, S, S):
            raise ValueError(f"expected images [B,3,{S},{S}], got {tuple(images.shape)}")
        E, W = self.cfg["embed_dim"], self.cfg["width"]
        need = self.workspace_bytes(B)
        self._ws_guard.enter()        # the workspace is shared by every call: order uses from different streams
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        dev = self.device
        seg = [torch.empty((B, self.P, E), dtype=torch.float32, device=dev) for _ in range(self.n_levels)]
        det = torch.empty((B, E), dtype=torch.float32, device=dev)
        tokens = [torch.empty((B, self.P, W), dtype=torch.float32, device=dev) for _ in range(self.n_levels)] \
            if want_tokens else None
        cls = torch.empty((B, W), dtype=torch.float32, device=dev) if want_cls else None
        tok_arr = _ptr_array(tokens) if tokens is not None else None
        query = [torch.empty((B, self.P, self.cfg["query_dim"]), dtype=torch.float32, device=dev)
                 for _ in range(self.n_levels)] if want_query else None
        qry_arr = _ptr_array(query) if query is not None else None
        with torch.cuda.device(dev):
            _lib.check(
                self._lib.aaclip_forward(self._h, _ptr(images), B, _ptr(self._ws), self._ws.numel(), _ptr_array(seg),
                                         _ptr(det), tok_arr, _ptr(cls), qry_arr, _stream()), "aaclip_forward")
        self._ws_guard.leave()
        if want_query:
            return seg, det, tokens, cls, query
        return seg, det, tokens, cls

    def seg_normalize(self, on: bool = True):
        """on=False: forward() returns the raw seg projections (no F.normalize pass); hand them to
        ``score_maps(..., raw=True)``, whose similarity kernel normalises in the one pass it makes anyway."""
        _lib.check(self._lib.aaclip_seg_normalize(self._h, int(on)), "aaclip_seg_normalize")
        self.seg_is_raw = not on

    # -- per-kernel-class device timing (bench.py roofline)
    def profile_enable(self, on: bool = True):
        _lib.check(self._lib.aaclip_profile_enable(self._h, int(on)), "aaclip_profile_enable")

    def profile_read(self) -> Dict[str, Dict[str, float]]:
        n = len(_lib.PROF_CLASSES)
        ms = (ctypes.c_float * n)()
        cnt = (ctypes.c_int32 * n)()
        fl = (ctypes.c_double * n)()
        _lib.check(self._lib.aaclip_profile_read(self._h, ms, cnt, fl), "aaclip_profile_read")
        return {name: dict(ms=float(ms[i]), launches=int(cnt[i]), flops=float(fl[i]))
                for i, name in enumerate(_lib.PROF_CLASSES)}

    @property
    def last_launch_count(self) -> int:
        return self._lib.aaclip_last_launch_count(self._h)


# ---------------------------------------------------------------------------------------- the text tower
DEFAULT_TEXT_CFG = dict(context=77, vocab=49408, width=768, heads=12, layers=12, mlp_dim=3072, embed_dim=768,
                        adapt_until=3, act=_lib.ACT_GELU_ERF, adapt_weight=0.1, max_sentences=512,
                        operand_dtype=_lib.DT_F16)


class AaclipTextEngine:
    """Owns one AaclipTextHandle: the adapted CLIP text tower (AdaptedCLIP.encode_text, model/adapter.py:273-304)."""

    def __init__(self, device: torch.device | str = "cuda", **cfg):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("AaclipTextEngine needs a CUDA (sm_100a) device; there is no CPU path")
        full = dict(DEFAULT_TEXT_CFG)
        full.update(cfg)
        self.cfg = full
        c = _lib.AaclipTextCfg()
        for k in ("context", "vocab", "width", "heads", "layers", "mlp_dim", "embed_dim", "adapt_until", "act",
                  "max_sentences", "operand_dtype"):
            setattr(c, k, int(full[k]))
        c.adapt_weight = float(full["adapt_weight"])
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.aaclip_text_create(ctypes.byref(c), ctypes.byref(self._h)), "aaclip_text_create")
        self._ws: Optional[torch.Tensor] = None
        self._ws_guard = StreamGuard()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.aaclip_text_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_weight(self, name: str, t: torch.Tensor):
        t = t.detach().to(self.device, torch.float32).contiguous()
        shape = (ctypes.c_int64 * t.dim())(*t.shape)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.aaclip_text_set_weight(self._h, name.encode(), _ptr(t), shape, t.dim(), _stream()),
                       f"aaclip_text_set_weight({name})")       # stream-ordered, see AaclipEngine.set_weight

    TEXT_KEYS = ("clipmodel.token_embedding.", "clipmodel.positional_embedding", "clipmodel.transformer.",
                 "clipmodel.ln_final.", "text_adapter.")

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> int:
        """Feed every text-path tensor of an AdaptedCLIP state_dict (clipmodel.* text tower + text_adapter.*)."""
        n = 0
        for k, v in sd.items():
            if k.startswith(self.TEXT_KEYS):
                self.set_weight(k, v)
                n += 1
        with torch.cuda.device(self.device):
            if self.missing == 0:
                _lib.check(self._lib.aaclip_text_finalize_weights(self._h, _stream()), "aaclip_text_finalize_weights")
            torch.cuda.current_stream().synchronize()
        return n

    def saturation_check(self, on: bool = True):
        _lib.check(self._lib.aaclip_text_saturation_check(self._h, int(on)), "aaclip_text_saturation_check")

    def saturation_count(self, reset: bool = True) -> int:
        v = ctypes.c_uint64(0)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.aaclip_text_saturation_count(self._h, ctypes.byref(v), int(reset), _stream()),
                       "aaclip_text_saturation_count")
        return int(v.value)

    @property
    def missing(self) -> int:
        return self._lib.aaclip_text_missing_weights(self._h)

    def encode_text(self, tokens: torch.Tensor) -> torch.Tensor:
        """tokens int [n, context] -> [n, E] fp32 on the device (un-normalised, as the reference returns them)."""
        if tokens.dim() != 2 or tokens.shape[1] != self.cfg["context"]:
            raise ValueError(f"tokens must be [n,{self.cfg['context']}], got {tuple(tokens.shape)}")
        if tokens.numel() and (int(tokens.min()) < 0 or int(tokens.max()) >= self.cfg["vocab"]):
            raise ValueError(f"token ids must lie in [0,{self.cfg['vocab']})")
        n = tokens.shape[0]
        if n == 0:
            return torch.empty((0, self.cfg["embed_dim"]), dtype=torch.float32, device=self.device)
        outs = []
        step = self.cfg["max_sentences"]
        for i in range(0, n, step):
            tk = tokens[i:i + step].to(self.device, torch.int32).contiguous()
            m = tk.shape[0]
            need = self._lib.aaclip_text_workspace_bytes(self._h, m)
            self._ws_guard.enter()
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            out = torch.empty((m, self.cfg["embed_dim"]), dtype=torch.float32, device=self.device)
            with torch.cuda.device(self.device):
                _lib.check(self._lib.aaclip_encode_text(self._h, _ptr(tk), m, _ptr(self._ws), self._ws.numel(), _ptr(out),
                                                        _stream()), "aaclip_encode_text")
            self._ws_guard.leave()
            outs.append(out)
        return outs[0] if len(outs) == 1 else torch.cat(outs)


def text_anchors(emb: torch.Tensor, group_sizes: Sequence[int]) -> torch.Tensor:
    """Prompt ensembling (forward_utils.py:152-161): emb [n,E] fp32, consecutive groups of `group_sizes` rows
    (normal, abnormal, normal, abnormal, ...) -> anchors [n_classes, E, 2] (aaclip_text_anchors)."""
    _need_cuda(emb)
    emb = emb.contiguous().float()
    if len(group_sizes) == 0 or len(group_sizes) % 2 or min(group_sizes) <= 0 or sum(group_sizes) != emb.shape[0]:
        raise ValueError("group_sizes must be a non-empty list of positive sizes, two per class, covering emb")
    off = torch.tensor([0] + list(torch.tensor(list(group_sizes)).cumsum(0)), dtype=torch.int32, device=emb.device)
    E = emb.shape[1]
    out = torch.empty((len(group_sizes) // 2, E, 2), dtype=torch.float32, device=emb.device)
    with torch.cuda.device(emb.device):
        _lib.check(_lib.load().aaclip_text_anchors(_ptr(emb), _ptr(off), len(group_sizes), E, _ptr(out), _stream()),
                   "aaclip_text_anchors")
    return out
