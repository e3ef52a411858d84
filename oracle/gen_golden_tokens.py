"""Generates tests/golden/tokenizer_prompts.npz with the UNMODIFIED reference tokenizer (TEST INFRASTRUCTURE; build
container only):   python oracle/gen_golden_tokens.py

Sentences: the full prompt ensemble (16 per class) of every class of every dataset in the reference's
dataset/constants.py (544 sentences; the 432 of MVTec + VisA are BASELINE config 3), plus strings that exercise the
word pattern, the byte fallback, html unescaping, whitespace folding, the special tokens and truncation.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
from oracle import ref_import  # noqa: E402

EXTRA = ["Hello, World!!  it's   a test... don't", "naïve café ÜBER straße 123 4567", "emoji \U0001F600 test ✓",
         "<|startoftext|> hi <|endoftext|>", "a&amp;b &lt;tag&gt; &amp;amp;", "x" * 30, "", "   ", "'s't're've'm'll'd",
         "日本語のテキスト", "tab\tnew\nline", "1234567890 3.14159", "CamelCaseWordsAreFun", "mixed123abc456",
         "under_score-dash/slash\\back", "ÀÉÎÕÜ àéîõü", "a" * 120]


def main():
    ref_import.install()
    os.chdir(ref_import.REF_ROOT)
    from dataset.constants import CLASS_NAMES, PROMPTS, REAL_NAMES
    from model.tokenizer import tokenize

    sentences = []
    for ds in CLASS_NAMES:
        for c in CLASS_NAMES[ds]:
            real = REAL_NAMES[ds][c]
            for states in (PROMPTS["prompt_normal"], PROMPTS["prompt_abnormal"]):
                sentences += [t.format(s.format(real)) for s in states for t in PROMPTS["prompt_templates"]]
    n_prompt = len(sentences)
    sentences += EXTRA
    ids = tokenize(sentences, truncate=True).numpy()
    too_long = []
    for i, s in enumerate(sentences):
        try:
            tokenize([s])
        except RuntimeError:
            too_long.append(i)
    path = os.path.join(REPO, "tests", "golden", "tokenizer_prompts.npz")
    np.savez_compressed(path, ids=ids, n_prompt=np.int64(n_prompt), too_long=np.array(too_long, dtype=np.int64),
                        sentences=np.frombuffer(json.dumps(sentences).encode("utf-8"), dtype=np.uint8))
    print(len(sentences), "sentences,", n_prompt, "prompts,", len(too_long), "need truncation;", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
