"""Reference import path ``model.adapter_modules`` -> aa-clip-iqm_b200/model/adapter_modules.py."""
from aaclip_b200.model.adapter_modules import SimpleAdapter, SimpleProj  # noqa: F401
