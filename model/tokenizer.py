"""Reference import path ``model.tokenizer`` (model/tokenizer.py) -> aa-clip-iqm_b200/tokenizer.py."""
from aaclip_b200.tokenizer import ClipTokenizer, default_tokenizer, tokenize  # noqa: F401

SimpleTokenizer = ClipTokenizer
