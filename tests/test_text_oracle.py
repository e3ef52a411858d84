"""CPU: the text-anchor restatement (oracle/aaclip_oracle.py encode_text / text_anchor) against golden vectors made
by executing the UNMODIFIED reference tokenizer + AdaptedCLIP.encode_text + get_adapted_single_class_text_embedding
(oracle/gen_golden_text.py). fp32 vs fp32."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import aaclip_oracle as orc


def test_text_oracle_matches_reference_golden():
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import weights
    fx = golden("text_3classes")
    sd = weights.make_text_state_dict(seed=int(fx["seed"]))
    assert sum(float(v.double().sum()) for v in sd.values()) == pytest.approx(float(fx["w_checksum"]), rel=0, abs=1e-4)
    tokens = torch.from_numpy(fx["tokens"])
    # reference tokenizer known-answer: SOT 49406 ... EOT 49407, int32, context 77
    assert tokens.dtype == torch.int32 and tokens.shape[1] == 77
    assert tokens[0, 0].item() == 49406 and tokens.max().item() == 49407
    with torch.no_grad():
        emb = orc.encode_text(sd, tokens)
    np.testing.assert_allclose(emb.numpy(), fx["emb"], rtol=0, atol=2e-5)
    sizes = fx["sizes"].tolist()
    r = 0
    for c in range(len(sizes) // 2):
        a, b = sizes[2 * c], sizes[2 * c + 1]
        anc = orc.text_anchor(emb[r:r + a], emb[r + a:r + a + b])
        r += a + b
        np.testing.assert_allclose(anc.numpy(), fx["anchors"][c], rtol=0, atol=2e-6)
        np.testing.assert_allclose(anc.norm(dim=0).numpy(), np.ones(2), atol=1e-6)   # unit-norm columns
    assert float(fx["oracle_vs_ref_emb"]) < 2e-5 and float(fx["oracle_vs_ref_anchor"]) < 2e-6
