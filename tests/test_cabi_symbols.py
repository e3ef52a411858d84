"""CPU: the C-ABI shared library builds, loads and exports every symbol include/aaclip.h declares."""
import ctypes
import os
import re

from conftest import REPO


def declared_symbols():
    src = open(os.path.join(REPO, "include", "aaclip.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aaclip_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_documented_entry_points():
    syms = declared_symbols()
    for s in ("aaclip_create", "aaclip_destroy", "aaclip_set_weight", "aaclip_forward", "aaclip_score_maps",
              "aaclip_workspace_bytes", "aaclip_last_error", "aaclip_gemm16", "aaclip_attention"):
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    from aaclip_b200 import _lib
    syms = declared_symbols()
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms, "ctypes binding and header disagree"
    raw = ctypes.CDLL(str(_lib.LIB_PATH))
    for s in syms:
        assert hasattr(raw, s), f"{s} not exported by {_lib.LIB_PATH}"
    assert b"sm_100a" in built_lib.aaclip_version()


def test_cfg_struct_layout_matches_header():
    from aaclip_b200 import _lib
    # 9 int32 + 4 levels + act + proj_relu + float + max_batch + operand_dtype + query_dim = 19 4-byte fields
    assert ctypes.sizeof(_lib.AaclipCfg) == 19 * 4
    hdr = open(os.path.join(REPO, "include", "aaclip.h")).read()
    body = hdr[hdr.index("typedef struct AaclipCfg {"):hdr.index("} AaclipCfg;")]
    fields = re.findall(r"^\s*(?:int32_t|float)\s+(\w+)", re.sub(r"/\*.*?\*/", "", body, flags=re.S), flags=re.M)
    assert fields == [n for n, _ in _lib.AaclipCfg._fields_]      # same names, same order as the ctypes mirror


def test_argument_errors_are_reported_without_a_gpu(built_lib):
    """Pure argument validation paths (no CUDA call is reached)."""
    assert built_lib.aaclip_workspace_bytes(None, 4) == 0
    assert built_lib.aaclip_score_workspace_bytes(2, 4, 1369) >= 2 * 4 * 2 * 1369 * 4
    assert built_lib.aaclip_missing_weights(None) == -1
    rc = built_lib.aaclip_create(None, None)
    assert rc == -1 and b"null" in built_lib.aaclip_last_error()


def test_product_path_fails_loudly_without_cuda():
    import pytest
    import torch
    from aaclip_b200 import engine
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        engine.AaclipEngine(device="cpu")
    with pytest.raises(RuntimeError):
        engine.layernorm(torch.zeros(4, 1024), torch.ones(1024), torch.zeros(1024))
