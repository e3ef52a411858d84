"""Reference import path ``dataset.constants`` -> aa-clip-iqm_b200/constants.py (the class / prompt tables)."""
import model  # noqa: F401
from aaclip_b200.constants import BASE_PATH, CLASS_NAMES, DATA_PATH, DOMAINS, PROMPTS, REAL_NAMES  # noqa: F401
