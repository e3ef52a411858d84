"""CPU tests of the autograd (training) route of the mirror against a reference-generated fixture.

tests/golden/train_step_224.npz holds the loss and (sub-sampled) gradients that the UNMODIFIED reference's
train.train_image_adapter (train.py:115-236) leaves after one batch (oracle/gen_golden_train.py). Here the same step
is replayed through this repository's AdaptedCLIP (plain-PyTorch route, model/torch_path.py) and forward_utils losses.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden


def _train_step_loss(model, fu, images, mask, label, text_feature, img_size):
    """The loss assembly of train.train_image_adapter (train.py:147-207), on this repository's call surface."""
    iqm_weight, text_weight = 0.4, 0.6
    patch_features, det_feature, iqm_outputs = model(images, text_embeddings=text_feature)
    cls_preds = torch.matmul(det_feature.unsqueeze(1), text_feature)[:, 0]
    loss = F.cross_entropy(cls_preds, label) * 0.5
    for f in patch_features:
        loss = loss + fu.calculate_seg_loss(fu.calculate_similarity_map(f, text_feature, img_size), mask) * text_weight * 0.5
    q = iqm_outputs.last_hidden_state
    norm_q, abn_q = q[:, 0, :], q[:, 1, :]
    for f in patch_features:
        diff = F.cosine_similarity(f, abn_q.unsqueeze(1), dim=-1) - F.cosine_similarity(f, norm_q.unsqueeze(1), dim=-1)
        p = torch.sigmoid(diff)
        B, L = p.shape
        H = int(np.sqrt(L))
        two = torch.cat([(1 - p).view(B, 1, H, H), p.view(B, 1, H, H)], dim=1)
        two = F.interpolate(two, size=(img_size, img_size), mode="bilinear", align_corners=False)
        loss = loss + fu.calculate_seg_loss(two, mask) * iqm_weight * 0.5
    return loss


@pytest.mark.timeout(600)
def test_one_training_step_matches_reference():
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import forward_utils as fu
    from aaclip_b200 import weights
    from aaclip_b200.model import AdaptedCLIP, create_model
    g = golden("train_step_224")
    S, B, seed, hidden = int(g["img_size"]), int(g["batch"]), int(g["seed"]), int(g["hidden"])
    clip = create_model("ViT-L-14-336", img_size=S, pretrained=None, force_image_size=S, device="cpu")
    model = AdaptedCLIP(clip_model=clip, relu=False, iqm_hidden_size=hidden).eval()      # eval: train.py:333
    sd = weights.make_visual_state_dict(seed=seed, img_size=S, relu=False)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected
    weights.fill_iqm_branch_(model, seed)
    images = weights.make_images(seed, B, S)
    mask = torch.from_numpy(g["mask"].astype(np.float32))
    label = torch.from_numpy(g["label"])
    text = torch.stack([weights.make_text_anchor(seed), weights.make_text_anchor(seed + 1)], dim=0)
    assert model._training_route()                       # adapters / IQM require grad -> plain PyTorch route, any device
    loss = _train_step_loss(model, fu, images, mask, label, text, S)
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 2e-4 * max(1.0, abs(float(g["loss"])))
    named = dict(model.named_parameters())
    for key in [k[5:] for k in g.files if k.startswith("grad:")]:
        gr = named[key].grad
        assert gr is not None and torch.isfinite(gr).all(), key
        ref = torch.from_numpy(g["grad:" + key])
        got = gr[:16, :64] if gr.dim() == 2 else gr[:64]
        scale = float(g["gnorm:" + key])
        assert abs(float(gr.double().norm()) - scale) < 2e-3 * scale, key
        assert float((got - ref).abs().max()) < 2e-3 * float(ref.abs().max()) + 1e-9, key
    # CLIP backbone parameters are frozen holders: the route must not have produced gradients for them
    assert all(p.grad is None for n, p in named.items() if n.startswith("clipmodel."))


def test_inference_route_still_refuses_cpu():
    """no_grad => engine route => no CPU / PyTorch fallback (there is no GPU in the CPU test environment)."""
    import aaclip_b200  # noqa: F401
    from aaclip_b200.model import AdaptedCLIP, create_model
    clip = create_model("ViT-L-14-336", img_size=224, pretrained=None, force_image_size=224, device="cpu")
    model = AdaptedCLIP(clip_model=clip, relu=False).eval()
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="CUDA"):
            model(torch.zeros(1, 3, 224, 224))
        with pytest.raises(RuntimeError, match="CUDA"):
            model.encode_text(torch.zeros(1, 77, dtype=torch.int32))


def test_similarity_map_autograd_matches_formula():
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import forward_utils as fu
    torch.manual_seed(0)
    f = F.normalize(torch.randn(2, 16 * 16, 768), dim=-1).requires_grad_()
    t = F.normalize(torch.randn(2, 768, 2), dim=1)
    out = fu.calculate_similarity_map(f, t, 224)                   # train mode: softmax over the two channels
    assert out.shape == (2, 2, 224, 224) and torch.allclose(out.sum(1), torch.ones(2, 224, 224), atol=1e-6)
    out.sum().backward()
    assert f.grad is not None
    m = fu.calculate_similarity_map(f, t, 224, test=True, domain="Industrial")
    assert m.shape == (2, 1, 224, 224)
