"""Prints the max abs error of the anomaly maps / features / scores against the reference's golden vectors for the
library selected by AACLIP_LIB_PATH / AACLIP_NVCC_EXTRA (kernel-variant A/B runs; the tests assert the tolerances).
    python tools/map_error.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import aaclip_b200  # noqa: F401
from aaclip_b200 import _lib

_lib.build()
from test_forward_gpu import _engine_outputs  # noqa: E402

for name in ("vitl_224_b2", "vitl_224_b1_relu", "vitl_518_b1", "vitl_336_b2"):
    try:
        fx, eng, sd, images, text, seg, det, tokens, cls, maps, score = _engine_outputs(name)
    except Exception as ex:  # noqa: BLE001
        print(name, "skipped:", ex)
        continue
    m = maps.cpu().numpy()
    ref = fx["maps_industrial"]
    print(f"{os.environ.get('AACLIP_NVCC_EXTRA', ''):30s} {name:18s} map err {np.abs(m - ref).max():.3e}  "
          f"score err {np.abs(score.cpu().numpy() - fx['score']).max():.2e}", flush=True)
