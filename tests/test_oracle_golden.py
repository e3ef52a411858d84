"""CPU: the oracle restatement (oracle/aaclip_oracle.py) against the golden vectors produced by executing the
UNMODIFIED reference (oracle/gen_golden.py). fp32 vs fp32: tolerances are rounding-order only."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import aaclip_oracle as orc


def _run(name):
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import weights
    fx = golden(name)
    S, B, seed, relu = int(fx["img_size"]), int(fx["batch"]), int(fx["seed"]), bool(fx["relu"])
    sd = weights.make_visual_state_dict(seed=seed, img_size=S, relu=relu)
    images = weights.make_images(seed, B, S)
    text = weights.make_text_anchor(seed)
    # the generators must reproduce the tensors the reference saw
    assert float(images.double().sum()) == pytest.approx(float(fx["img_checksum"]), rel=0, abs=1e-6)
    assert sum(float(v.double().sum()) for v in sd.values()) == pytest.approx(float(fx["w_checksum"]), rel=0, abs=1e-5)
    assert np.array_equal(text.numpy(), fx["text"])
    with torch.no_grad():
        out = orc.predict(sd, images, text, S, domain="Industrial")
        out["maps_med"] = orc.anomaly_maps(out["seg"], text, S, "Medical")
        out["train0"] = orc.calculate_similarity_map(out["seg"][0], text, S, test=False)
    return fx, out


@pytest.mark.parametrize("name", ["vitl_224_b2", "vitl_224_b1_relu", "vitl_518_b1", "vitl_336_b2", "vitl_518_b2"])
def test_oracle_matches_reference_golden(name):
    fx, out = _run(name)
    rows = fx["rows"]
    for l in range(4):
        np.testing.assert_allclose(out["seg"][l][:, rows, :].numpy(), fx["seg_rows"][l], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out["det"].numpy(), fx["det"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["score"].numpy(), fx["score"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["maps"].numpy(), fx["maps_industrial"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(out["maps_med"].numpy()[:, ::2, ::2], fx["maps_medical"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(out["train0"].numpy()[:, :, ::4, ::4], fx["train0"], rtol=0, atol=1e-5)
    # recorded at generation time: restatement vs reference, same process
    assert float(fx["oracle_vs_ref_maps"]) < 2e-5 and float(fx["oracle_vs_ref_seg"]) < 2e-6


def test_blur_restatement_matches_torchvision():
    """kornia is not installable here ('parity unpinned' for that one call); torchvision's gaussian_blur implements the
    same published definition (normalised exp(-x^2/2s^2) taps, reflect border) and must agree."""
    import torchvision.transforms.functional as TF
    x = torch.randn(2, 1, 37, 37)
    for k, s in ((7, 1.0), (9, 1.5)):
        a = orc.gaussian_blur2d(x, k, s)
        b = TF.gaussian_blur(x, [k, k], [s, s])
        assert (a - b).abs().max().item() < 1e-6


def test_similarity_map_shapes_and_modes():
    f = torch.nn.functional.normalize(torch.randn(2, 16 * 16, 768), dim=-1)
    t = torch.nn.functional.normalize(torch.randn(768, 2), dim=0)
    m = orc.calculate_similarity_map(f, t, 224, test=True, domain="Industrial")
    assert m.shape == (2, 1, 224, 224)
    p = orc.calculate_similarity_map(f, t, 224, test=False)
    assert p.shape == (2, 2, 224, 224)
    assert torch.allclose(p.sum(1), torch.ones(2, 224, 224), atol=1e-6)
    # per-image anchors give the same answer as a shared one when they are equal
    m2 = orc.calculate_similarity_map(f, t[None].repeat(2, 1, 1), 224, test=True, domain="Industrial")
    assert torch.equal(m, m2)
