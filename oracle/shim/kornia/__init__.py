"""Stand-in for kornia==0.6.9 (requirements.txt:3), which is not installed in this image. Only what the reference
imports on the hot path: kornia.filters.gaussian_blur2d (forward_utils.py:8,208) and hflip/vflip (utils.py)."""
from . import filters, geometry  # noqa: F401
