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
        assert iqm_outputs is None
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
    # stand-in for the reference BPE tokenizer (out of scope): replays the ids the reference tokenizer produced
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
        with pytest.raises(NotImplementedError):
            model.encode_text(tokens[:1].cuda(), adapt_text=False)


def test_grad_mode_is_rejected():
    model = _model(224, False, 1)
    for p in model.image_adapter.parameters():
        p.requires_grad_(True)
    with pytest.raises(RuntimeError, match="no_grad"):
        model(torch.zeros(1, 3, 224, 224, device="cuda"))
