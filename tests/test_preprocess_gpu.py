"""GPU: aaclip_preprocess_u8 (bicubic resize + ToTensor + Normalize on the device) through the C ABI, bit-exact against
the PIL + torchvision golden vectors and against the oracle on seeded ragged sizes."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import golden
from oracle import pil_resample as pr
from oracle.gen_golden_preprocess import LARGE, SMALL, make_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pre(built_lib):
    from aaclip_b200 import preprocess
    return preprocess


def _bits(t):
    return t.detach().cpu().numpy().view(np.uint32)


def test_small_cases_equal_pil_golden(pre):
    g = golden("preprocess")
    for i, (H, W, S) in enumerate(SMALL):
        out = pre.ImagePreprocessor(S)(g[f"small{i}_src"])
        assert out.shape == (1, 3, S, S)
        assert np.array_equal(_bits(out[0]), g[f"small{i}_out"].view(np.uint32)), (H, W, S)


def test_baseline_size_checksum_equals_pil(pre):
    g = golden("preprocess")
    for i, (H, W, S) in enumerate(LARGE):
        out = pre.ImagePreprocessor(S)(make_image(200 + i, H, W))
        assert hashlib.sha256(out[0].cpu().numpy().tobytes()).digest() == g[f"large{i}_sha256"].tobytes()


@pytest.mark.parametrize("H,W,S", [(1, 1, 8), (2, 3, 5), (17, 301, 37), (301, 17, 37), (518, 518, 518), (64, 100, 64),
                                   (100, 64, 64), (999, 1001, 336), (13, 4099, 24), (200, 40, 24), (40, 200, 24)])
def test_ragged_sizes_equal_the_oracle(pre, H, W, S):
    rng = np.random.default_rng(H * 7919 + W)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    out = pre.ImagePreprocessor(S)(img)
    assert np.array_equal(_bits(out[0]), pr.preprocess(img, S).view(np.uint32))


def test_batch_rows_start_at_unaligned_offsets(pre):
    """W * 3 = 393 bytes per row: every row of every image starts at a different offset modulo 16."""
    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, (5, 77, 131, 3), dtype=np.uint8)
    out = pre.ImagePreprocessor(56)(imgs)
    for b in range(5):
        assert np.array_equal(_bits(out[b]), pr.preprocess(imgs[b], 56).view(np.uint32))


def test_output_feeds_the_tower_layout_and_errors(pre):
    p = pre.ImagePreprocessor(56)
    out = torch.empty((2, 3, 56, 56), device="cuda")
    assert p(np.zeros((2, 40, 40, 3), np.uint8), out=out) is out
    ref = pr.preprocess(np.zeros((40, 40, 3), np.uint8), 56)
    assert np.array_equal(_bits(out[1]), ref.view(np.uint32))
    with pytest.raises(ValueError):
        p(np.zeros((2, 40, 40, 4), np.uint8))
    with pytest.raises(ValueError):
        p(np.zeros((40, 40, 3), np.float32))


def test_predict_u8_equals_the_reference_transform_then_the_tower(built_lib):
    """decoded uint8 images -> maps through AnomalyPredictor.predict_u8 == predict_device on the oracle's transform."""
    from aaclip_b200 import pipeline, weights
    S = 224
    sd = weights.make_visual_state_dict(seed=0, img_size=S)
    pred = pipeline.AnomalyPredictor(sd, weights.make_text_anchor(0), img_size=S, max_batch=2)
    imgs = np.stack([make_image(7 + i, 300, 410) for i in range(2)])
    maps_a, scores_a = pred.predict_u8(imgs)
    maps_a, scores_a = maps_a.clone(), scores_a.clone()
    x = torch.from_numpy(np.stack([pr.preprocess(im, S) for im in imgs])).cuda()
    maps_b, scores_b = pred.predict_device(x)
    assert torch.equal(maps_a, maps_b) and torch.equal(scores_a, scores_b)
