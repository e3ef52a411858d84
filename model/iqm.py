"""Reference import path ``model.iqm`` (model/iqm.py) -> aa-clip-iqm_b200/model/iqm.py."""
from aaclip_b200.model.iqm import IQM, IQMConfig, IQMEncoder, IQMLayer  # noqa: F401
