"""Import shim: exposes the package directory ``aa-clip-iqm_b200/`` under the importable name ``aaclip_b200``."""
import importlib.util
import sys
from pathlib import Path

_pkg_dir = Path(__file__).resolve().parent / "aa-clip-iqm_b200"
_spec = importlib.util.spec_from_file_location(
    "aaclip_b200", _pkg_dir / "__init__.py", submodule_search_locations=[str(_pkg_dir)]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["aaclip_b200"] = _mod
_spec.loader.exec_module(_mod)
