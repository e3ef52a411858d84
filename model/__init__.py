"""Top-level ``model`` package with the reference's import paths (``from model.adapter import AdaptedCLIP``,
``from model.clip import create_model``, ``from model.tokenizer import tokenize`` -- test_last.py:14-15,
train.py:14-15, forward_utils.py:11): thin re-exports of the sm_100a-backed implementation in ``aa-clip-iqm_b200/``.
Put the repository root on ``sys.path`` (or run a script from it) and the reference's scripts import unchanged."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import aaclip_b200  # noqa: E402,F401  (registers the package directory `aa-clip-iqm_b200` under an importable name)
