"""Host-side mirror of the reference's ``model`` package surface for the hot path (model/adapter.py, model/clip.py):
same class / function names, constructor arguments, attribute names, state_dict keys and return values, with the
compute delegated to the sm_100a engine through the C ABI. Nothing here runs the path on the CPU or in PyTorch ops."""
from .adapter import AdaptedCLIP  # noqa: F401
from .adapter_modules import SimpleAdapter, SimpleProj  # noqa: F401
from .clip import create_model, resize_pos_embed  # noqa: F401
