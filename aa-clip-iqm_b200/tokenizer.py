"""CLIP byte-pair tokenizer, written from scratch for the text-anchor path.

Replaces the reference's ``model/tokenizer.py:74-186`` (``SimpleTokenizer`` / ``tokenize``): same published algorithm
(GPT-2 style reversible byte-to-unicode table, the CLIP word pattern, lowest-rank-first pair merging over the
48 894 merges of ``bpe_simple_vocab_16e6.txt.gz``) and therefore the same ids, pinned bit for bit against ids the
reference tokenizer produced (tests/golden/text_3classes.npz, tests/golden/tokenizer_prompts.npz).

The merge table is DATA (OpenAI's published CLIP vocabulary); it lives in ``data/bpe_simple_vocab_16e6.txt.gz``.
``ftfy`` is optional: when it is not installed the text-repair step is skipped, which is the identity on the ASCII
prompts of dataset/constants.py.

Implementation notes (how this differs from the reference's loop): words are kept as lists of vocabulary ids, the
merge table maps an id pair to (rank, merged id), and one pass of ``_merge_word`` finds the best-ranked adjacent pair
and rewrites the list in place; results are memoised per word.
"""
from __future__ import annotations

import gzip
import html
import os
from functools import lru_cache
from typing import Dict, List, Sequence, Tuple, Union

import regex
import torch

CONTEXT_LENGTH = 77
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "bpe_simple_vocab_16e6.txt.gz")
_N_MERGES = 49152 - 256 - 2          # merges used by CLIP: vocabulary 49 408 = 2 x 256 bytes + merges + 2 specials
_WORD_RE = regex.compile(
    r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+", regex.IGNORECASE)
_SPACE_RE = regex.compile(r"\s+")


def _byte_alphabet() -> List[str]:
    """The reversible byte -> printable-unicode table of GPT-2 / CLIP: printable latin-1 bytes map to themselves, the
    other 68 bytes to code points 256, 257, ... in byte order. Returned as a list indexed by byte value."""
    keep = set(range(0x21, 0x7F)) | set(range(0xA1, 0xAD)) | set(range(0xAE, 0x100))
    table, spill = [], 0
    for b in range(256):
        if b in keep:
            table.append(chr(b))
        else:
            table.append(chr(256 + spill))
            spill += 1
    return table


def _alphabet_order() -> List[int]:
    """Byte values in the order the vocabulary enumerates them: the kept (printable) bytes first, then the rest."""
    keep = list(range(0x21, 0x7F)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    seen = set(keep)
    return keep + [b for b in range(256) if b not in seen]


class ClipTokenizer:
    def __init__(self, merges_path: str = _DATA):
        if not os.path.exists(merges_path):
            raise FileNotFoundError(f"BPE merge table not found: {merges_path}")
        with gzip.open(merges_path) as f:
            lines = f.read().decode("utf-8").split("\n")
        merges = [tuple(ln.split()) for ln in lines[1:1 + _N_MERGES]]
        alpha = _byte_alphabet()
        order = _alphabet_order()
        # vocabulary: 256 byte symbols, the same 256 with the end-of-word marker, one entry per merge, two specials
        symbols = [alpha[b] for b in order]
        vocab = symbols + [s + "</w>" for s in symbols] + [a + b for a, b in merges]
        self.sot, self.eot = len(vocab), len(vocab) + 1
        vocab += ["<|startoftext|>", "<|endoftext|>"]
        self.vocab = vocab
        self.index: Dict[str, int] = {s: i for i, s in enumerate(vocab)}
        # (left id, right id) -> (rank, merged id)
        self.pair_table: Dict[Tuple[int, int], Tuple[int, int]] = {}
        for rank, (a, b) in enumerate(merges):
            self.pair_table[(self.index[a], self.index[b])] = (rank, self.index[a + b])
        self.byte_id = [self.index[alpha[b]] for b in range(256)]            # mid-word byte symbol ids
        self.byte_end_id = [self.index[alpha[b] + "</w>"] for b in range(256)]  # word-final byte symbol ids
        self._memo: Dict[bytes, Tuple[int, ...]] = {}

    # ---------------------------------------------------------------------------------------------- BPE
    def _merge_word(self, raw: bytes) -> Tuple[int, ...]:
        ids = [self.byte_id[b] for b in raw[:-1]] + [self.byte_end_id[raw[-1]]]
        table = self.pair_table
        while len(ids) > 1:
            best_rank, best = None, None
            for i in range(len(ids) - 1):
                hit = table.get((ids[i], ids[i + 1]))
                if hit is not None and (best_rank is None or hit[0] < best_rank):
                    best_rank, best = hit[0], (ids[i], ids[i + 1], hit[1])
            if best is None:
                break
            left, right, merged = best
            out, i, n = [], 0, len(ids)
            while i < n:                       # every non-overlapping occurrence, left to right
                if i + 1 < n and ids[i] == left and ids[i + 1] == right:
                    out.append(merged)
                    i += 2
                else:
                    out.append(ids[i])
                    i += 1
            ids = out
        return tuple(ids)

    def encode(self, text: str) -> List[int]:
        text = _SPACE_RE.sub(" ", _clean(text)).strip().lower()
        out: List[int] = []
        for word in _WORD_RE.findall(text):
            if word == "<|startoftext|>":
                out.append(self.sot)
                continue
            if word == "<|endoftext|>":
                out.append(self.eot)
                continue
            raw = word.encode("utf-8")
            ids = self._memo.get(raw)
            if ids is None:
                ids = self._memo[raw] = self._merge_word(raw)
            out.extend(ids)
        return out

    def decode(self, ids: Sequence[int]) -> str:
        alpha = _byte_alphabet()
        back = {c: b for b, c in enumerate(alpha)}
        text = "".join(self.vocab[int(i)] for i in ids)
        words = [bytes(back[c] for c in w if c in back).decode("utf-8", errors="replace") for w in text.split("</w>")]
        return " ".join(words)


def _clean(text: str) -> str:
    try:
        import ftfy
        text = ftfy.fix_text(text)
    except ImportError:
        pass
    return html.unescape(html.unescape(text)).strip()


@lru_cache(maxsize=1)
def default_tokenizer() -> ClipTokenizer:
    return ClipTokenizer()


def tokenize(texts: Union[str, List[str]], context_length: int = CONTEXT_LENGTH, truncate: bool = False) -> torch.Tensor:
    """strings -> int32 [n, context_length]: <start> ids <end>, zero padded (model/tokenizer.py:150-186). A sentence
    longer than the context raises RuntimeError unless `truncate` (then it is cut and closed with <end>)."""
    if isinstance(texts, str):
        texts = [texts]
    tk = default_tokenizer()
    out = torch.zeros(len(texts), context_length, dtype=torch.int32)
    for r, t in enumerate(texts):
        ids = [tk.sot] + tk.encode(t) + [tk.eot]
        if len(ids) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {t} is too long for context length {context_length}")
            ids = ids[:context_length]
            ids[-1] = tk.eot
        out[r, :len(ids)] = torch.tensor(ids, dtype=torch.int32)
    return out
