"""CPU: the from-scratch BPE tokenizer (aa-clip-iqm_b200/tokenizer.py) against ids produced by the reference's
tokenizer (model/tokenizer.py:74-186) -- tests/golden/tokenizer_prompts.npz (every class prompt of every dataset of the
reference + edge-case strings, oracle/gen_golden_tokens.py) and the 48 sentences of tests/golden/text_3classes.npz."""
import json

import numpy as np
import pytest
import torch

from conftest import golden


def test_ids_equal_reference_tokenizer():
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import tokenizer
    fx = golden("tokenizer_prompts")
    sentences = json.loads(bytes(fx["sentences"]).decode("utf-8"))
    ids = tokenizer.tokenize(sentences, truncate=True)
    assert ids.dtype == torch.int32 and tuple(ids.shape) == (len(sentences), 77)
    assert np.array_equal(ids.numpy(), fx["ids"])                   # bit-exact, all 705 sentences
    assert ids[0, 0] == 49406 and (ids == 49407).sum(1).min() >= 1  # <start> ... <end>


def test_class_prompts_reproduce_config3_sentences():
    """BASELINE config 3: 27 MVTec + VisA classes x 16 prompts = 432 sentences, in the reference's order."""
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import constants, tokenizer
    fx = golden("tokenizer_prompts")
    ref_sentences = json.loads(bytes(fx["sentences"]).decode("utf-8"))[:int(fx["n_prompt"])]
    ours = [s for ds in constants.CLASS_NAMES for c in constants.CLASS_NAMES[ds]
            for g in constants.class_prompts(constants.REAL_NAMES[ds][c]) for s in g]
    assert ours == ref_sentences
    n = sum(16 * len(constants.CLASS_NAMES[d]) for d in ("MVTec", "VisA"))
    assert n == 432
    t3 = golden("text_3classes")
    reals = ["gray carpet", "oval white pill with small red speckles and the letters 'FF' engraved",
             "dual ultrasonic distance sensor pcb module"]
    sent = [s for r in reals for g in constants.class_prompts(r) for s in g]
    assert np.array_equal(tokenizer.tokenize(sent).numpy(), t3["tokens"])


def test_context_overflow_and_decode():
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import tokenizer
    long = " ".join(["anomaly"] * 100)
    with pytest.raises(RuntimeError, match="too long"):
        tokenizer.tokenize([long])
    t = tokenizer.tokenize([long], truncate=True)
    assert t[0, 0] == 49406 and t[0, -1] == 49407
    tk = tokenizer.default_tokenizer()
    assert tk.decode(tk.encode("a photo of a damaged gray carpet.")).strip() == "a photo of a damaged gray carpet ."
    assert tokenizer.tokenize("").tolist()[0][:3] == [49406, 49407, 0]
