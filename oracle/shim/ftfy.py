"""Stub for `ftfy` (model/tokenizer.py:18,63). The reference prompts are ASCII (dataset/constants.py:78-148)."""


def fix_text(t):
    return t
