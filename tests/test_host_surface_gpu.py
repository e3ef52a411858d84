"""GPU: the reference-facing Python surface (AdaptedCLIP.forward / encode_text, forward_utils.calculate_similarity_map,
get_adapted_single_class_text_embedding) driven the way test_last.get_predictions drives the reference
(test_last.py:70-151), against golden vectors produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu


def _model(S, relu, seed, text_seed=None):
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import model, weights
    clip = model.create_model("ViT-L-14-336", S, pretrained=None, force_image_size=S, device="cuda")
    m = model.AdaptedCLIP(clip_model=clip, relu=relu, iqm_hidden_size=512, iqm_num_layers=2, iqm_num_heads=8).to("cuda")
    m.eval()
    sd = dict(weights.make_visual_state_dict(seed=seed, img_size=S, relu=relu))
    if text_seed is not None:
        sd.update(weights.make_text_state_dict(seed=text_seed))
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected
    return m


@pytest.mark.parametrize("name", ["vitl_224_b2", "vitl_224_b1_relu"])
def test_get_predictions_style_loop_matches_reference(name):
    from aaclip_b200 import forward_utils as fu
    from aaclip_b200 import weights
    fx = golden(name)
    S, B, seed, relu = int(fx["img_size"]), int(fx["batch"]), int(fx["seed"]), bool(fx["relu"])
    model = _model(S, relu, seed)
    image = weights.make_images(seed, B, S).to("cuda")
    class_text_embeddings = weights.make_text_anchor(seed).to("cuda")
    with torch.no_grad():
        # --- the body of test_last.get_predictions (no-IQM branch), verbatim in structure
        epoch_text_feature = class_text_embeddings.unsqueeze(0).repeat(image.size(0), 1, 1)
        patch_features, det_feature, iqm_outputs = model(image, text_embeddings=epoch_text_feature)
        assert tuple(iqm_outputs.last_hidden_state.shape) == (B, 2, 512)      # the IQM branch ran (text given)
        text_anomaly_maps = [fu.calculate_similarity_map(f, epoch_text_feature, S, test=True, domain="Industrial")
                             for f in patch_features]
        final_map = torch.cat(text_anomaly_maps, dim=1).sum(1)
        train0 = fu.calculate_similarity_map(patch_features[0], class_text_embeddings, S, test=False)
        fused_map, score = fu.anomaly_maps(patch_features, epoch_text_feature, det_feature, S, "Industrial")
    assert text_anomaly_maps[0].shape == (B, 1, S, S) and train0.shape == (B, 2, S, S)
    assert np.abs(final_map.cpu().numpy() - fx["maps_industrial"]).max() < 1e-2
    assert np.abs(fused_map.cpu().numpy() - fx["maps_industrial"]).max() < 1e-2
    assert np.abs(train0.cpu().numpy()[:, :, ::4, ::4] - fx["train0"]).max() < 1e-2
    assert np.abs(score.cpu().numpy() - fx["score"]).max() < 1e-3
    assert np.abs(det_feature.cpu().numpy() - fx["det"]).max() < 2e-3
    # reloading adapter weights (as a later checkpoint would) must reach the engine
    with torch.no_grad():
        model.image_adapter["det_proj"].fc[0].weight.mul_(-1.0) if relu else model.image_adapter["det_proj"].fc.weight.mul_(-1.0)
        _, det2, _ = model(image)
    assert not torch.allclose(det2, det_feature)


def test_text_embedding_surface_matches_reference():
    from aaclip_b200 import forward_utils as fu
    fx = golden("text_3classes")
    model = _model(224, False, 1, text_seed=int(fx["seed"]))
    tokens, sizes = torch.from_numpy(fx["tokens"]), fx["sizes"].tolist()
    reals = {"carpet": "gray carpet", "pill": "oval white pill with small red speckles and the letters 'FF' engraved",
             "pcb1": "dual ultrasonic distance sensor pcb module"}   # dataset/constants.py REAL_NAMES of the fixture's classes
    # the built-in BPE tokenizer must reproduce the ids the reference tokenizer produced for these 48 sentences
    from aaclip_b200 import constants, tokenizer
    sentences = [s for c in reals.values() for g in constants.class_prompts(c) for s in g]
    assert torch.equal(tokenizer.tokenize(sentences), tokens.to(torch.int32))
    cursor = {"i": 0}

    def replay_tokenize(sentences):
        i = cursor["i"]
        cursor["i"] += len(sentences)
        return tokens[i:i + len(sentences)]

    with torch.no_grad():
        a0 = fu.get_adapted_single_class_text_embedding(model, "MVTec", "carpet", "cuda", tokenize=replay_tokenize,
                                                        real_name=reals["carpet"])
        assert a0.shape == (768, 2)
        assert np.abs(a0.cpu().numpy() - fx["anchors"][0]).max() < 2e-3
        cursor["i"] = 0
        all_ = fu.get_adapted_text_embedding(model, "MVTec", "cuda", tokenize=replay_tokenize, real_names=reals)
        assert list(all_) == list(reals)
        for i, c in enumerate(reals):
            assert np.abs(all_[c].cpu().numpy() - fx["anchors"][i]).max() < 2e-3
        # default tokenizer + REAL_NAMES lookup (no overrides): same anchors
        a1 = fu.get_adapted_single_class_text_embedding(model, "VisA", "pcb1", "cuda")
        assert np.abs(a1.cpu().numpy() - fx["anchors"][2]).max() < 2e-3
        # the un-adapted CLIP text tower (test_last.py:264-271 without a text-adapter checkpoint): host PyTorch
        plain = model.encode_text(tokens[:2].cuda(), adapt_text=False)
        assert plain.shape == (2, 768) and torch.isfinite(plain).all()


def test_grad_mode_takes_the_autograd_route():
    """train.py:150 calls the model with autograd on: the plain-PyTorch route must agree with the engine route
    (same parameters) and produce gradients for the adapters."""
    from aaclip_b200 import weights
    model = _model(224, False, 1)
    x = weights.make_images(1, 1, 224).to("cuda")
    with torch.no_grad():
        seg_e, det_e, _ = model(x)
    for p in model.image_adapter.parameters():
        p.requires_grad_(True)
    seg_t, det_t, _ = model(x)
    assert det_t.requires_grad
    assert (det_t - det_e).abs().max().item() < 2e-3 and (seg_t[3] - seg_e[3]).abs().max().item() < 2e-2
    det_t.sum().backward()
    g = model.image_adapter["layer_adapters"][0].fc[0].weight.grad
    assert g is not None and torch.isfinite(g).all() and g.abs().max().item() > 0


@pytest.mark.parametrize("name", ["iqm_224_h768", "iqm_224_h512"])
def test_get_predictions_with_iqm_matches_reference(name):
    """The body of the reference's evaluation loop, test_last.get_predictions (test_last.py:70-151), IQM branch
    included, against what the unmodified reference produced for the same seeded weights and batch
    (oracle/gen_golden_iqm.py): the fused 0.6 * text + 0.4 * IQM maps and the fork's [B,2] image predictions."""
    import torch.nn as nn
    import torch.nn.functional as F
    from aaclip_b200 import forward_utils as fu
    from aaclip_b200 import model as M
    from aaclip_b200 import weights
    fx = golden(name)
    S, B, seed, hidden = int(fx["img_size"]), int(fx["batch"]), int(fx["seed"]), int(fx["hidden"])
    clip = M.create_model("ViT-L-14-336", S, pretrained=None, force_image_size=S, device="cuda")
    model = M.AdaptedCLIP(clip_model=clip, relu=False, iqm_hidden_size=hidden, iqm_num_layers=2, iqm_num_heads=8).to("cuda")
    model.eval()
    missing, unexpected = model.load_state_dict(weights.make_visual_state_dict(seed=seed, img_size=S), strict=False)
    assert not unexpected
    assert weights.fill_iqm_branch_(model, seed) == int(fx["n_iqm_params"])
    device, img_size = "cuda", S
    class_text_embeddings = weights.make_text_anchor(seed).to(device)
    iqm_weight, text_weight = 0.4, 0.6
    with torch.no_grad():
        torch.manual_seed(1234)
        image = weights.make_images(seed, B, S).to(device)
        epoch_text_feature = class_text_embeddings.unsqueeze(0).repeat(image.size(0), 1, 1)
        patch_features, det_feature, iqm_outputs = model(image, text_embeddings=epoch_text_feature)
        pred = det_feature @ epoch_text_feature
        pred = (pred[:, 1] + 1) / 2
        text_anomaly_maps = [fu.calculate_similarity_map(f, epoch_text_feature, img_size, test=True, domain="Industrial")
                             for f in patch_features]
        iqm_anomaly_maps = []
        final_query_embedding = iqm_outputs.last_hidden_state
        norm_query, abnorm_query = final_query_embedding[:, 0, :], final_query_embedding[:, 1, :]
        for f in patch_features:
            if norm_query.shape[-1] != f.shape[-1]:
                proj_layer = nn.Linear(norm_query.shape[-1], f.shape[-1]).to(f.device)   # seeded CPU init, as the reference
                norm_query, abnorm_query = proj_layer(norm_query), proj_layer(abnorm_query)
            norm_sim = F.cosine_similarity(f, norm_query.unsqueeze(1), dim=-1)
            abnorm_sim = F.cosine_similarity(f, abnorm_query.unsqueeze(1), dim=-1)
            iqm_pred = torch.sigmoid(abnorm_sim - norm_sim)
            Bq, L = iqm_pred.shape
            H = int(L ** 0.5)
            iqm_pred = F.interpolate(iqm_pred.view(Bq, 1, H, H), size=(img_size, img_size), mode="bilinear",
                                     align_corners=False)
            iqm_anomaly_maps.append(iqm_pred)
        text_map = torch.cat(text_anomaly_maps, dim=1).sum(1, keepdim=True)
        iqm_map = torch.cat(iqm_anomaly_maps, dim=1).sum(1, keepdim=True)
        final_map = text_map * text_weight + iqm_map * iqm_weight
    assert final_map.shape == (B, 1, S, S) and pred.shape == (B, 2)
    assert np.abs(iqm_outputs.last_hidden_state.cpu().numpy() - fx["iqm_last_hidden"]).max() < 2e-2
    assert np.abs(final_map.cpu().numpy() - fx["preds"]).max() < 1e-2      # north_star map tolerance, fused map
    assert np.abs(pred.cpu().numpy() - fx["preds_image"]).max() < 1e-3
