"""aa-clip-iqm_b200: B200-native (sm_100a) implementation of the AA-CLIP inference hot path.

Layout
  csrc/            hand-written CUDA kernels (tcgen05/TMEM/TMA GEMM + attention, fused elementwise) and the C ABI
  _lib.py          build + ctypes binding of libaaclip_b200.so (include/aaclip.h)
  engine.py        torch-tensor front end of the C ABI
  model/, forward_utils.py   host-side mirror of the reference call surface (model/adapter.py, forward_utils.py)

The directory name is not a Python identifier; import it as ``aaclip_b200`` (see aaclip_b200.py at the repo root).
"""
from . import _lib  # noqa: F401
from ._lib import build, load  # noqa: F401

__all__ = ["build", "load"]
