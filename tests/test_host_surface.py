"""CPU: the host-side mirror of the reference call surface (aa-clip-iqm_b200/model, forward_utils) against fixtures
made by executing the unmodified reference (oracle/gen_golden_host.py): state_dict contract, positional-embedding
resize, checkpoint loading as test_last.py:231-247 does it, metrics_eval semantics (row a11), and
the "no fallback" rule."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import REPO, golden


@pytest.fixture(scope="module")
def clip224():
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import model
    return model.create_model("ViT-L-14-336", 518, pretrained=None, force_image_size=518)


@pytest.mark.parametrize("relu", [False, True])
def test_state_dict_keys_and_shapes_match_reference(clip224, relu):
    from aaclip_b200 import model
    ref = json.load(open(os.path.join(REPO, "tests", "golden", "ref_state_dict_keys.json")))["relu" if relu else "norelu"]
    m = model.AdaptedCLIP(clip224, relu=relu, iqm_hidden_size=512)
    ours = {k: list(v.shape) for k, v in m.state_dict().items()}
    # the fixture holds EVERY key of the reference's AdaptedCLIP.state_dict() (CLIP towers, adapters, the IQM branch:
    # iqm.*, class_query_mlp.*, query_adapters.*, pos_embedding, visual_weight, text_weight, iqm_layer_norm.*)
    assert len(ref) == 856
    missing = [k for k in ref if k not in ours]
    assert not missing, missing[:5]
    assert all(ours[k] == ref[k] for k in ref), [k for k in ref if ours[k] != ref[k]][:5]
    extra = [k for k in ours if k not in ref]
    assert not extra, extra[:5]


def test_resize_pos_embed_matches_reference():
    import types
    from aaclip_b200 import model
    fx = golden("pos_embed_resize")
    g = torch.Generator().manual_seed(int(fx["seed"]))
    old = torch.randn(577, 1024, generator=g) * 0.03
    for S in (518, 224):
        gs = S // 14
        sd = {"visual.positional_embedding": old.clone()}
        model.resize_pos_embed(sd, types.SimpleNamespace(visual=types.SimpleNamespace(grid_size=(gs, gs))))
        new = sd["visual.positional_embedding"]
        assert new.shape == (gs * gs + 1, 1024)
        assert torch.equal(new[0], old[0])                      # class-token row untouched
        np.testing.assert_allclose(new[:, ::16].numpy(), fx[f"pos_{S}"], rtol=0, atol=1e-6)
    same = {"visual.positional_embedding": old.clone()}
    model.resize_pos_embed(same, types.SimpleNamespace(visual=types.SimpleNamespace(grid_size=(24, 24))))
    assert torch.equal(same["visual.positional_embedding"], old)   # same grid -> untouched


def test_create_model_routes(tmp_path, monkeypatch):
    from aaclip_b200 import model
    from aaclip_b200.model import clip as clip_mod
    # random-init route: img_size alone is ignored, force_image_size applies (model/clip.py:159-161)
    m = model.create_model("ViT-L-14-336", 518, pretrained=None)
    assert m.visual.positional_embedding.shape == (577, 1024)
    with pytest.raises(RuntimeError, match="not found"):
        model.create_model("ViT-B-16-plus-240", 240)
    with pytest.raises(RuntimeError, match="required"):
        model.create_model("ViT-L-14-336", 518, pretrained=None, require_pretrained=True)
    with pytest.raises(RuntimeError, match="not found"):
        model.create_model("ViT-L-14-336", 518, pretrained="openai", require_pretrained=True)
    # openai route from a fabricated 336-px checkpoint: fp16 tensors promoted, positional embedding resampled to 518
    src = model.create_model("ViT-L-14-336", 336, pretrained=None)
    sd = {k: v.half() if v.dim() > 1 else v for k, v in src.state_dict().items()}
    path = tmp_path / "fab.pt"
    torch.save(sd, path)
    monkeypatch.setitem(clip_mod._MODEL_CKPT_PATHS, "ViT-L-14-336", str(path))
    m = model.create_model("ViT-L-14-336", 518, pretrained="openai", require_pretrained=True)
    assert m.visual.positional_embedding.shape == (1370, 1024) and m.visual.positional_embedding.dtype == torch.float32
    assert torch.equal(m.visual.positional_embedding[0], sd["visual.positional_embedding"][0].float())
    assert torch.equal(m.visual.conv1.weight, sd["visual.conv1.weight"].float())


def test_adapter_checkpoints_load_like_test_last(clip224):
    """test_last.py:231-247: model.text_adapter.load_state_dict(ckpt['text_adapter']), same for image_adapter."""
    from aaclip_b200 import model, weights
    m = model.AdaptedCLIP(clip224, relu=False)
    sd = weights.make_visual_state_dict(seed=1, img_size=518)
    img_ckpt = {k[len("image_adapter."):]: v for k, v in sd.items() if k.startswith("image_adapter.")}
    m.image_adapter.load_state_dict(img_ckpt)                     # strict
    tsd = {k[len("text_adapter."):]: v for k, v in weights.make_text_state_dict(seed=1).items()
           if k.startswith("text_adapter.")}
    m.text_adapter.load_state_dict(tsd)
    assert torch.equal(m.image_adapter["seg_proj"][2].fc.weight, sd["image_adapter.seg_proj.2.fc.weight"])
    # whole-model load through the reference's key names (image_encoder.* aliases clipmodel.visual.*)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected
    assert torch.equal(m.clipmodel.visual.conv1.weight, sd["image_encoder.conv1.weight"])


def test_no_cpu_or_pytorch_fallback(clip224):
    from aaclip_b200 import model
    m = model.AdaptedCLIP(clip224, relu=False)
    x = torch.zeros(1, 3, 518, 518)
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="CUDA"):
            m(x)
        with pytest.raises(RuntimeError, match="CUDA"):
            m.encode_text(torch.zeros(1, 77, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="ClipWeights"):
        clip224(x)
    from aaclip_b200 import engine, forward_utils as fu
    with pytest.raises(RuntimeError, match="CUDA"):       # scoring under no_grad = engine route: no torch fallback
        with torch.no_grad():
            fu.calculate_similarity_map(torch.zeros(1, 1369, 768), torch.zeros(768, 2), 518, test=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        engine.text_anchors(torch.zeros(16, 768), [6, 10])


def test_metrics_eval_semantics():
    """Against an independent numpy/sklearn evaluation of the formulae in forward_utils.py:233-308."""
    from sklearn.metrics import average_precision_score, roc_auc_score
    from aaclip_b200 import forward_utils as fu
    rng = np.random.default_rng(0)
    n, S = 6, 32
    masks = (rng.random((n, 1, S, S)) > 0.9).astype(np.int64)
    labels = np.array([0, 1, 1, 0, 1, 0])
    preds = rng.normal(size=(n, S, S)).astype(np.float32) + 2.0 * masks[:, 0]
    scores = rng.random(n).astype(np.float32)
    for domain in ("Industrial", "Medical"):
        res = fu.metrics_eval(masks, labels, preds, scores, "carpet", domain)
        pp = (preds - preds.min()) / (preds.max() - preds.min())
        ip = (scores - scores.min()) / (scores.max() - scores.min())
        img = pp.max(axis=(1, 2)) * 0.5 + ip * 0.5 if domain == "Industrial" else pp.max(axis=(1, 2))
        assert res["pixel AUC"] == round(roc_auc_score(masks.flatten(), pp.flatten()), 4) * 100
        assert res["pixel AP"] == round(average_precision_score(masks.flatten(), pp.flatten()), 4) * 100
        assert res["image AUC"] == round(roc_auc_score(labels, img), 4) * 100
        assert res["class name"] == "carpet"
    # the fork's [B,2] image-score quirk: column 0 is used (forward_utils.py:269-270); single-label class -> 0
    res = fu.metrics_eval(masks, np.ones(n), preds, np.stack([scores, scores[::-1]], 1), "x", "Industrial")
    assert res["image AUC"] == 0 and res["image AP"] == 0
    res4 = fu.metrics_eval(masks, labels, preds[:, None], scores, "x", "Industrial")      # [B,1,S,S] maps accepted
    assert res4["pixel AUC"] == fu.metrics_eval(masks, labels, preds, scores, "x", "Industrial")["pixel AUC"]


def test_reference_import_paths_and_cli_defaults():
    """The repository root answers the reference's import block (test_last.py:13-22, train.py:13-22) and the
    evaluation driver keeps the reference's command line (test_last.py:161-190)."""
    from dataset import DOMAINS, get_dataset  # noqa: F401
    from forward_utils import (calculate_seg_loss, calculate_similarity_map, get_adapted_single_class_text_embedding,  # noqa: F401
                               get_adapted_text_embedding, metrics_eval, visualize)
    from model.adapter import AdaptedCLIP
    from model.clip import create_model  # noqa: F401
    from model.tokenizer import tokenize
    from utils import cos_sim, setup_seed
    import aaclip_b200.model
    assert AdaptedCLIP is aaclip_b200.model.AdaptedCLIP
    assert DOMAINS["MVTec"] == "Industrial" and DOMAINS["Brain"] == "Medical"
    assert tokenize("a photo of a damaged gray carpet.")[0, :10].tolist() == [49406, 320, 1125, 539, 320, 13568, 7048,
                                                                             6922, 269, 49407]
    setup_seed(111)
    a = torch.nn.functional.normalize(torch.randn(3, 8), dim=-1)
    assert torch.allclose(cos_sim(a, a[:1]), a[:1] @ a.t())
    with pytest.raises(NotImplementedError):
        get_dataset("MVTec", 518, "few_shot", 4, "train")
    with pytest.raises(NotImplementedError):
        visualize()
    import test_last as drv
    d = drv.build_parser().parse_args([])
    assert (d.model_name, d.img_size, d.dataset, d.image_batch_size, d.seed, d.iqm_hidden_size, d.iqm_num_layers,
            d.iqm_num_heads, d.image_adapt_until, d.text_adapt_until) == ("ViT-L-14-336", 518, "MVTec", 32, 111, 512, 2, 8,
                                                                            6, 3)
