"""GPU parity of the text-anchor path (BASELINE config 3) through the C ABI (aaclip_encode_text /
aaclip_text_anchors) against golden vectors produced by the unmodified reference (oracle/gen_golden_text.py).
Tolerances: 16-bit tensor-core operands -> anchors (unit vectors) within 2e-3 abs."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu

ANCHOR_TOL = 2e-3
EMB_REL_TOL = 5e-3   # un-normalised sentence embeddings, relative to their largest magnitude


@pytest.fixture(scope="module")
def text_engine():
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import engine, weights
    fx = golden("text_3classes")
    sd = weights.make_text_state_dict(seed=int(fx["seed"]))
    eng = engine.AaclipTextEngine("cuda", max_sentences=64)
    n = eng.load_state_dict(sd)
    assert eng.missing == 0 and n == len(sd)
    return fx, eng, sd


def test_encode_text_matches_reference_golden(text_engine):
    from aaclip_b200 import engine
    fx, eng, sd = text_engine
    tokens = torch.from_numpy(fx["tokens"]).cuda()
    eng.saturation_check(True)
    emb = eng.encode_text(tokens)
    torch.cuda.synchronize()
    assert eng.saturation_count() == 0
    eng.saturation_check(False)
    err = np.abs(emb.cpu().numpy() - fx["emb"]).max()
    scale = np.abs(fx["emb"]).max()
    print(f"sentence embeddings: max abs err {err:.3e} (|max| {scale:.2f})")
    assert err < EMB_REL_TOL * scale
    anchors = engine.text_anchors(emb, fx["sizes"].tolist())
    torch.cuda.synchronize()
    aerr = np.abs(anchors.cpu().numpy() - fx["anchors"]).max()
    print(f"anchors: max abs err {aerr:.3e}")
    assert aerr < ANCHOR_TOL
    n = anchors.norm(dim=1)
    assert torch.allclose(n, torch.ones_like(n), atol=1e-5)


def test_encode_text_is_batch_invariant_and_chunks(text_engine):
    """A sentence's embedding must not depend on which other sentences share the call (ragged last chunk included)."""
    fx, eng, sd = text_engine
    tokens = torch.from_numpy(fx["tokens"]).cuda()
    full = eng.encode_text(tokens)
    one = eng.encode_text(tokens[5:6])
    part = eng.encode_text(tokens[3:20])
    torch.cuda.synchronize()
    assert torch.equal(full[5:6], one) and torch.equal(full[3:20], part)


def test_text_anchor_ensemble_against_oracle(text_engine):
    """The ensembling kernel alone, fp32 in / fp32 out, against the restated forward_utils.py:152-161 on the box."""
    from aaclip_b200 import engine
    from oracle import aaclip_oracle as orc
    g = torch.Generator().manual_seed(3)
    emb = torch.randn(23, 768, generator=g)
    sizes = [6, 10, 1, 6]
    got = engine.text_anchors(emb.cuda(), sizes).cpu()
    want = torch.stack([orc.text_anchor(emb[0:6], emb[6:16]), orc.text_anchor(emb[16:17], emb[17:23])])
    assert (got - want).abs().max().item() < 1e-6


def test_text_argument_errors(text_engine):
    from aaclip_b200 import engine
    fx, eng, sd = text_engine
    with pytest.raises(ValueError):
        eng.encode_text(torch.zeros(2, 76, dtype=torch.int32))
    with pytest.raises(ValueError):
        eng.encode_text(torch.full((1, 77), 49408, dtype=torch.int32))
    with pytest.raises(ValueError):
        engine.text_anchors(torch.randn(4, 768).cuda(), [1, 2])            # does not cover emb
    assert eng.encode_text(torch.zeros(0, 77, dtype=torch.int32)).shape == (0, 768)   # empty input
    fresh = engine.AaclipTextEngine("cuda", max_sentences=4)
    with pytest.raises(RuntimeError, match="weight tensors not set"):
        fresh.encode_text(torch.zeros(1, 77, dtype=torch.int32))
