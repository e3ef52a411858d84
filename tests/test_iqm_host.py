"""CPU: the host-side IQM branch (aa-clip-iqm_b200/model/iqm.py + AdaptedCLIP._iqm_branch) against the reference's
IQM branch on the same seeded weights and inputs (tests/golden/iqm_*.npz `branch_alone`, oracle/gen_golden_iqm.py)."""
import numpy as np
import pytest
import torch

from conftest import golden


@pytest.mark.parametrize("name", ["iqm_224_h768", "iqm_224_h512"])
def test_iqm_branch_matches_reference(name):
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import weights
    from aaclip_b200.model import AdaptedCLIP, create_model
    from aaclip_b200.model.clip import ClipWeights, MODEL_CONFIGS
    fx = golden(name)
    S, B, seed, hidden = int(fx["img_size"]), int(fx["batch"]), int(fx["seed"]), int(fx["hidden"])
    clip = ClipWeights(MODEL_CONFIGS["ViT-L-14-336"], init="empty")           # the branch never touches the towers
    model = AdaptedCLIP(clip_model=clip, relu=False, iqm_hidden_size=hidden).eval()
    assert weights.fill_iqm_branch_(model, seed) == int(fx["n_iqm_params"])
    assert abs(sum(float(p.double().sum()) for p in model.iqm.parameters()) - float(fx["w_checksum"])) < 1e-6
    g = torch.Generator().manual_seed(seed + 100)
    P = (S // 14) ** 2
    tok = [torch.randn(B, P, 1024, generator=g) for _ in range(4)]
    cls = torch.randn(B, 1024, generator=g)
    text = torch.from_numpy(fx["text"])
    assert model.iqm_factored
    with torch.no_grad():
        out = model._iqm_branch(tok, cls, text.unsqueeze(0).repeat(B, 1, 1))        # factored cross-attention
        model.iqm_factored = False
        lit = model._iqm_branch(tok, cls, text.unsqueeze(0).repeat(B, 1, 1))        # the reference's literal order
        model.iqm_factored = True
    assert tuple(out.last_hidden_state.shape) == (B, 2, hidden)
    err_f = np.abs(out.last_hidden_state.numpy() - fx["branch_alone"]).max()
    err_l = np.abs(lit.last_hidden_state.numpy() - fx["branch_alone"]).max()
    print(f"{name}: IQM branch vs reference: factored {err_f:.2e}, literal {err_l:.2e}")
    assert err_l < 1e-5 and err_f < 1e-5
    # a [768,2] anchor (no batch axis) takes the reference's transpose route: 2 text tokens of width 768 -> new lazy layer
    with torch.no_grad():
        out2 = model._iqm_branch(tok, cls, text)
    assert model.text_feature_proj.in_features == 768 and tuple(out2.last_hidden_state.shape) == (B, 2, hidden)
