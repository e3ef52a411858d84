"""Reference import path ``model.adapter`` (model/adapter.py) -> aa-clip-iqm_b200/model/adapter.py."""
from aaclip_b200.model.adapter import AdaptedCLIP, IQM, IQMConfig, SimpleAdapter, SimpleProj  # noqa: F401
