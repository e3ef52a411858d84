"""GPU: fp32-accuracy mode (BASELINE config 4, AACLIP_DT_F32) -- the same launch sequence with fp32 SIMT GEMM /
attention kernels -- against golden vectors produced by the unmodified fp32 reference.
north_star tolerances for fp32 mode: anomaly maps within 1e-4 abs, image scores within 1e-3."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu

MAP_TOL_F32 = 1e-4
SCORE_TOL = 1e-3


def _run(name, max_batch=4):
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import _lib, engine, weights
    fx = golden(name)
    S, B, seed, relu = int(fx["img_size"]), int(fx["batch"]), int(fx["seed"]), bool(fx["relu"])
    sd = weights.make_visual_state_dict(seed=seed, img_size=S, relu=relu)
    eng = engine.AaclipEngine("cuda", img_size=S, proj_relu=int(relu), max_batch=max_batch, operand_dtype=_lib.DT_F32)
    eng.load_state_dict(sd)
    assert eng.missing == 0
    images = weights.make_images(seed, B, S).cuda()
    text = weights.make_text_anchor(seed).cuda()
    seg, det, tokens, cls = eng.forward(images, want_tokens=True, want_cls=True)
    maps, score = engine.score_maps(seg, text, S, test=True, ksize=7, sigma=1.0, det=det)
    torch.cuda.synchronize()
    return fx, eng, seg, det, tokens, cls, maps, score


@pytest.mark.parametrize("name", ["vitl_336_b2", "vitl_224_b2", "vitl_224_b1_relu", "vitl_518_b1"])
def test_fp32_mode_matches_reference_golden(name):
    fx, eng, seg, det, tokens, cls, maps, score = _run(name)
    rows = torch.as_tensor(fx["rows"]).cuda()
    ferr = max(np.abs(seg[l][:, rows, :].cpu().numpy() - fx["seg_rows"][l]).max() for l in range(4))
    merr = np.abs(maps.cpu().numpy() - fx["maps_industrial"]).max()
    serr = np.abs(score.cpu().numpy() - fx["score"]).max()
    print(f"{name}: fp32 mode map max abs err {merr:.3e}, feature err {ferr:.3e}, score err {serr:.3e}")
    assert merr < MAP_TOL_F32
    assert serr < SCORE_TOL and serr < 1e-5
    assert ferr < 2e-6
    assert np.abs(det.cpu().numpy() - fx["det"]).max() < 1e-6
    assert np.abs(cls.cpu().numpy() - fx["cls"]).max() < 5e-5


def test_fp32_mode_config4_batch_and_agreement_with_fp16():
    """Config 4 shape: @336 in one larger batch (images 0..1 are the golden ones); the fp16 tensor-core path must agree
    with fp32 mode to its own tolerance on every image of the batch."""
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import _lib, engine, weights
    fx = golden("vitl_336_b2")
    S, seed, B = 336, int(fx["seed"]), 8
    sd = weights.make_visual_state_dict(seed=seed, img_size=S)
    images = weights.make_images(seed, B, S).cuda()
    text = weights.make_text_anchor(seed).cuda()
    out = {}
    for dt in (_lib.DT_F32, _lib.DT_F16):
        eng = engine.AaclipEngine("cuda", img_size=S, max_batch=B, operand_dtype=dt)
        eng.load_state_dict(sd)
        seg, det, _, _ = eng.forward(images)
        out[dt] = engine.score_maps(seg, text, S, test=True, ksize=7, sigma=1.0, det=det)
        eng.close()
    torch.cuda.synchronize()
    m32, s32 = out[_lib.DT_F32]
    m16, s16 = out[_lib.DT_F16]
    assert np.abs(m32[:2].cpu().numpy() - fx["maps_industrial"]).max() < MAP_TOL_F32     # batch-invariant too
    assert (m32 - m16).abs().max().item() < 1e-2 and (s32 - s16).abs().max().item() < 1e-3


def test_fp32_text_anchors():
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import _lib, engine, weights
    fx = golden("text_3classes")
    sd = weights.make_text_state_dict(seed=int(fx["seed"]))
    eng = engine.AaclipTextEngine("cuda", max_sentences=64, operand_dtype=_lib.DT_F32)
    eng.load_state_dict(sd)
    assert eng.missing == 0
    emb = eng.encode_text(torch.from_numpy(fx["tokens"]).cuda())
    anchors = engine.text_anchors(emb, fx["sizes"].tolist())
    torch.cuda.synchronize()
    eerr = np.abs(emb.cpu().numpy() - fx["emb"]).max()
    aerr = np.abs(anchors.cpu().numpy() - fx["anchors"]).max()
    print(f"fp32 text: embedding err {eerr:.3e}, anchor err {aerr:.3e}")
    assert eerr < 5e-5 and aerr < 1e-5      # suggested fp32 tolerance for anchors: 1e-4
