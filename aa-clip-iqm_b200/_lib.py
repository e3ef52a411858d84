"""ctypes binding of libaaclip_b200.so (C ABI declared in include/aaclip.h).

The library is built IN-TREE by :func:`build` (``nvcc -gencode arch=compute_100a,code=sm_100a``) and loaded by
:func:`load`. There is no fallback: if the shared object is missing or fails to load, every product entry point
raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
# AACLIP_LIB_PATH: developer knob for kernel-variant sweeps (tools/); the product default is the in-tree library
LIB_PATH = Path(os.environ.get("AACLIP_LIB_PATH") or (PKG_DIR / "libaaclip_b200.so"))
HEADER = REPO_ROOT / "include" / "aaclip.h"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]

MAX_LEVELS = 4

# epilogue selectors (include/aaclip.h)
EPI_BIAS_BF16 = 0
EPI_BIAS_GELU_BF16 = 1
EPI_BIAS_QGELU_BF16 = 2
EPI_BIAS_RESID_F32 = 3
EPI_F32 = 4
EPI_LEAKY_F32 = 5
EPI_BIAS_F32 = 7
EPI_LN_BIAS_BF16 = 8
EPI_LN_BIAS_GELU_BF16 = 9
EPI_LN_BIAS_QGELU_BF16 = 10

DT_F16 = 0
DT_BF16 = 1
DT_F32 = 2   # fp32-accuracy mode (handles only)

PROF_CLASSES = ("gemm_misc", "attention", "layernorm", "other", "gemm_qkv", "gemm_out", "gemm_fc", "gemm_proj")

ACT_GELU_ERF = 0
ACT_QUICK_GELU = 1


class AaclipCfg(ctypes.Structure):
    """Mirror of ``struct AaclipCfg`` (include/aaclip.h)."""

    _fields_ = [
        ("img_size", ctypes.c_int32),
        ("patch", ctypes.c_int32),
        ("width", ctypes.c_int32),
        ("heads", ctypes.c_int32),
        ("layers", ctypes.c_int32),
        ("mlp_dim", ctypes.c_int32),
        ("embed_dim", ctypes.c_int32),
        ("adapt_until", ctypes.c_int32),
        ("n_levels", ctypes.c_int32),
        ("levels", ctypes.c_int32 * MAX_LEVELS),
        ("act", ctypes.c_int32),
        ("proj_relu", ctypes.c_int32),
        ("adapt_weight", ctypes.c_float),
        ("max_batch", ctypes.c_int32),
        ("operand_dtype", ctypes.c_int32),
        ("query_dim", ctypes.c_int32),
    ]


class AaclipTextCfg(ctypes.Structure):
    """Mirror of ``struct AaclipTextCfg`` (include/aaclip.h)."""

    _fields_ = [
        ("context", ctypes.c_int32),
        ("vocab", ctypes.c_int32),
        ("width", ctypes.c_int32),
        ("heads", ctypes.c_int32),
        ("layers", ctypes.c_int32),
        ("mlp_dim", ctypes.c_int32),
        ("embed_dim", ctypes.c_int32),
        ("adapt_until", ctypes.c_int32),
        ("act", ctypes.c_int32),
        ("adapt_weight", ctypes.c_float),
        ("max_sentences", ctypes.c_int32),
        ("operand_dtype", ctypes.c_int32),
    ]


def _sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [HEADER]


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    if os.environ.get("AACLIP_NO_REBUILD"):   # pre-built kernel variants (tools/build_variants.sh) are used as they are
        return False
    t = LIB_PATH.stat().st_mtime
    return any(p.stat().st_mtime > t for p in _sources())


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu for sm_100a into libaaclip_b200.so next to this file (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libaaclip_b200.so")
    extra = os.environ.get("AACLIP_NVCC_EXTRA", "").split()   # e.g. -DAACLIP_ATTN_POLY_EVERY=0 for a variant sweep
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-o", str(LIB_PATH), str(CSRC / "aaclip_api.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None

_c_void_pp = ctypes.POINTER(ctypes.c_void_p)


def _declare(lib):
    vp, i32, i64, f32, sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t
    lib.aaclip_version.restype = ctypes.c_char_p
    lib.aaclip_version.argtypes = []
    lib.aaclip_last_error.restype = ctypes.c_char_p
    lib.aaclip_last_error.argtypes = []
    lib.aaclip_create.restype = i32
    lib.aaclip_create.argtypes = [ctypes.POINTER(AaclipCfg), _c_void_pp]
    lib.aaclip_destroy.restype = None
    lib.aaclip_destroy.argtypes = [vp]
    lib.aaclip_set_weight.restype = i32
    lib.aaclip_set_weight.argtypes = [vp, ctypes.c_char_p, vp, ctypes.POINTER(i64), i32, vp]
    lib.aaclip_missing_weights.restype = i32
    lib.aaclip_missing_weights.argtypes = [vp]
    lib.aaclip_workspace_bytes.restype = sz
    lib.aaclip_workspace_bytes.argtypes = [vp, i32]
    lib.aaclip_forward.restype = i32
    lib.aaclip_forward.argtypes = [vp, vp, i32, vp, sz, _c_void_pp, vp, _c_void_pp, vp, _c_void_pp, vp]
    lib.aaclip_score_workspace_bytes.restype = sz
    lib.aaclip_score_workspace_bytes.argtypes = [i32, i32, i32]
    lib.aaclip_score_maps.restype = i32
    lib.aaclip_score_maps.argtypes = [_c_void_pp, i32, vp, i32, i32, i32, i32, i32, i32, i32, f32, vp, sz, vp, vp,
                                      vp, vp]
    lib.aaclip_score_maps_raw.restype = i32
    lib.aaclip_score_maps_raw.argtypes = lib.aaclip_score_maps.argtypes
    lib.aaclip_seg_normalize.restype = i32
    lib.aaclip_seg_normalize.argtypes = [vp, i32]
    lib.aaclip_adapter.restype = i32
    lib.aaclip_adapter.argtypes = [i32, vp, vp, vp, i32, i32, f32, vp, sz, vp, vp]
    # (guarded: tools/variants/*.so built from an older commit for an A/B run may predate these entry points;
    # tests/test_cabi_symbols.py checks that the in-tree library exports every declared symbol)
    if hasattr(lib, "aaclip_split_hilo"):
        lib.aaclip_split_hilo.restype = i32
        lib.aaclip_split_hilo.argtypes = [i32, vp, vp, vp, i64, vp]
        lib.aaclip_join_hilo.restype = i32
        lib.aaclip_join_hilo.argtypes = [i32, vp, vp, vp, i64, vp]
        lib.aaclip_gemm16_resid_hl.restype = i32
        lib.aaclip_gemm16_resid_hl.argtypes = [i32, vp, vp, i32, i32, i32, vp, vp, vp, vp, i32, vp]
    lib.aaclip_last_launch_count.restype = i32
    lib.aaclip_last_launch_count.argtypes = [vp]
    lib.aaclip_stat_columns.restype = i32
    lib.aaclip_stat_columns.argtypes = []
    u64p = ctypes.POINTER(ctypes.c_uint64)
    for pre in ("aaclip_", "aaclip_text_"):
        f = getattr(lib, pre + "finalize_weights")
        f.restype, f.argtypes = i32, [vp, vp]
        f = getattr(lib, pre + "saturation_check")
        f.restype, f.argtypes = i32, [vp, i32]
        f = getattr(lib, pre + "saturation_count")
        f.restype, f.argtypes = i32, [vp, u64p, i32, vp]
    lib.aaclip_profile_enable.restype = i32
    lib.aaclip_profile_enable.argtypes = [vp, i32]
    lib.aaclip_profile_read.restype = i32
    lib.aaclip_profile_read.argtypes = [vp, ctypes.POINTER(f32), ctypes.POINTER(i32), ctypes.POINTER(ctypes.c_double)]
    lib.aaclip_text_create.restype = i32
    lib.aaclip_text_create.argtypes = [ctypes.POINTER(AaclipTextCfg), _c_void_pp]
    lib.aaclip_text_destroy.restype = None
    lib.aaclip_text_destroy.argtypes = [vp]
    lib.aaclip_text_set_weight.restype = i32
    lib.aaclip_text_set_weight.argtypes = [vp, ctypes.c_char_p, vp, ctypes.POINTER(i64), i32, vp]
    lib.aaclip_text_missing_weights.restype = i32
    lib.aaclip_text_missing_weights.argtypes = [vp]
    lib.aaclip_text_workspace_bytes.restype = sz
    lib.aaclip_text_workspace_bytes.argtypes = [vp, i32]
    lib.aaclip_encode_text.restype = i32
    lib.aaclip_encode_text.argtypes = [vp, vp, i32, vp, sz, vp, vp]
    lib.aaclip_text_anchors.restype = i32
    lib.aaclip_text_anchors.argtypes = [vp, vp, i32, i32, vp, vp]
    lib.aaclip_metrics_workspace_bytes.restype = sz
    lib.aaclip_metrics_workspace_bytes.argtypes = []
    lib.aaclip_minmax.restype = i32
    lib.aaclip_minmax.argtypes = [vp, i64, vp, vp, i32, vp]
    lib.aaclip_score_histogram.restype = i32
    lib.aaclip_score_histogram.argtypes = [vp, vp, i64, vp, i32, vp, vp]
    lib.aaclip_auc_ap.restype = i32
    lib.aaclip_auc_ap.argtypes = [vp, i32, vp, vp]
    lib.aaclip_row_max.restype = i32
    lib.aaclip_row_max.argtypes = [vp, i32, i64, vp, vp]
    lib.aaclip_resample_ksize.restype = i32
    lib.aaclip_resample_ksize.argtypes = [i32, i32]
    lib.aaclip_resample_coeffs_host.restype = i32
    lib.aaclip_resample_coeffs_host.argtypes = [i32, i32, vp, vp, i32]
    lib.aaclip_preprocess_workspace_bytes.restype = sz
    lib.aaclip_preprocess_workspace_bytes.argtypes = [i32, i32, i32]
    lib.aaclip_preprocess_u8.restype = i32
    lib.aaclip_preprocess_u8.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, vp, ctypes.POINTER(f32), ctypes.POINTER(f32),
                                         vp, sz, vp, vp]
    lib.aaclip_gemm16.restype = i32
    lib.aaclip_gemm16.argtypes = [i32, vp, vp, i32, i32, i32, i32, vp, vp, vp, i32, vp]
    lib.aaclip_ln_fold.restype = i32
    lib.aaclip_ln_fold.argtypes = [i32, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp]
    lib.aaclip_gemm16_ln.restype = i32
    lib.aaclip_gemm16_ln.argtypes = [i32, vp, vp, i32, i32, i32, i32, vp, vp, vp, i32, i32, vp, i32, vp]
    lib.aaclip_gemm16_ln_tok.restype = i32
    lib.aaclip_gemm16_ln_tok.argtypes = [i32, vp, vp, i32, i32, i32, i32, vp, vp, vp, i32, i32, i32, vp, i32, vp]
    lib.aaclip_gemm16_resid.restype = i32
    lib.aaclip_gemm16_resid.argtypes = [i32, vp, vp, i32, i32, i32, vp, vp, vp, vp, i32, vp]
    lib.aaclip_attention.restype = i32
    lib.aaclip_attention.argtypes = [i32, vp, vp, i32, i32, i32, i32, vp]
    lib.aaclip_layernorm.restype = i32
    lib.aaclip_layernorm.argtypes = [i32, vp, vp, vp, vp, vp, i32, i32, i32, vp]
    lib.aaclip_pack16.restype = i32
    lib.aaclip_pack16.argtypes = [i32, vp, vp, i64, i64, i64, vp]


EXPORTED_SYMBOLS = [
    "aaclip_version", "aaclip_last_error", "aaclip_create", "aaclip_destroy", "aaclip_set_weight",
    "aaclip_missing_weights", "aaclip_workspace_bytes", "aaclip_forward", "aaclip_score_workspace_bytes",
    "aaclip_score_maps", "aaclip_last_launch_count", "aaclip_stat_columns", "aaclip_profile_enable", "aaclip_profile_read", "aaclip_gemm16", "aaclip_attention", "aaclip_layernorm",
    "aaclip_pack16",
    "aaclip_metrics_workspace_bytes", "aaclip_minmax", "aaclip_score_histogram", "aaclip_auc_ap", "aaclip_row_max",
    "aaclip_ln_fold", "aaclip_gemm16_ln", "aaclip_gemm16_resid",
    "aaclip_text_create", "aaclip_text_destroy", "aaclip_text_set_weight", "aaclip_text_missing_weights",
    "aaclip_text_workspace_bytes", "aaclip_encode_text", "aaclip_text_anchors",
    "aaclip_resample_ksize", "aaclip_resample_coeffs_host", "aaclip_preprocess_workspace_bytes", "aaclip_preprocess_u8",
    "aaclip_finalize_weights", "aaclip_saturation_check", "aaclip_saturation_count",
    "aaclip_text_finalize_weights", "aaclip_text_saturation_check", "aaclip_text_saturation_count",
    "aaclip_score_maps_raw", "aaclip_seg_normalize", "aaclip_adapter", "aaclip_gemm16_ln_tok",
    "aaclip_split_hilo", "aaclip_join_hilo", "aaclip_gemm16_resid_hl",
]


def load():
    """dlopen the in-tree library (raises RuntimeError if it has not been built)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU or PyTorch fallback for the AA-CLIP hot path."
            )
        lib = ctypes.CDLL(str(LIB_PATH))
        _declare(lib)
        _lib = lib
    return _lib


def check(rc: int, what: str = "aaclip call"):
    if rc != 0:
        msg = load().aaclip_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
