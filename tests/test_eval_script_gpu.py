"""GPU: the evaluation driver (test_last.py at the repository root: the role and command line of the reference's
test_last.py:160-330) end to end on a tiny synthetic dataset written to disk in the reference's format
(dataset/metadata/<name>/full-shot.jsonl records + image / mask files): loader -> device transform -> tower + IQM branch
-> fused maps -> metrics table."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make_dataset(root, per_class=4, size=(96, 80)):
    from PIL import Image
    rng = np.random.default_rng(7)
    meta_dir = os.path.join(root, "metadata", "BTAD")
    os.makedirs(meta_dir)
    rows = []
    for cls in ("01", "02", "03"):
        for i in range(per_class):
            label = int(i % 2 == 1)
            ipath = f"{cls}/test/{'ko' if label else 'ok'}/{i:03d}.png"
            os.makedirs(os.path.dirname(os.path.join(root, "data", ipath)), exist_ok=True)
            Image.fromarray(rng.integers(0, 256, (size[1], size[0], 3), dtype=np.uint8)).save(os.path.join(root, "data", ipath))
            row = {"image_path": ipath, "label": label, "class_name": cls}
            if label:
                mpath = f"{cls}/ground_truth/ko/{i:03d}_mask.png"
                os.makedirs(os.path.dirname(os.path.join(root, "data", mpath)), exist_ok=True)
                m = np.zeros((size[1], size[0]), dtype=np.uint8)
                m[10:40, 20:60] = 255
                Image.fromarray(m).save(os.path.join(root, "data", mpath))
                row["mask_path"] = mpath
            rows.append(row)
    with open(os.path.join(meta_dir, "full-shot.jsonl"), "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
    return os.path.join(root, "data"), os.path.join(root, "metadata")


def test_evaluation_driver_end_to_end(tmp_path):
    import test_last as drv
    data, meta = _make_dataset(str(tmp_path))
    common = ["--dataset", "BTAD", "--img_size", "224", "--image_batch_size", "2", "--num_workers", "0",
              "--iqm_hidden_size", "768", "--synthetic_weights", "--data_path", data, "--metadata_root", meta]
    df_dev = drv.main(common + ["--save_path", str(tmp_path / "a")])
    df_cpu = drv.main(common + ["--save_path", str(tmp_path / "b"), "--no_device_transform"])
    assert list(df_dev["class name"]) == ["01", "02", "03", "Average"]
    vals = df_dev[["pixel AUC", "pixel AP", "image AUC", "image AP"]].to_numpy(dtype=np.float64)
    assert np.isfinite(vals).all() and (vals >= 0).all() and (vals <= 100).all()
    # the device-side transform is bit-exact with PIL + torchvision: both runs must print the same table
    assert df_dev.equals(df_cpu)
    assert os.path.exists(tmp_path / "a" / "test.log")
    # device metrics + the well-posed image score: same pixel metrics to 3 decimals (AUROC / AP of the same maps)
    df_m = drv.main(common + ["--save_path", str(tmp_path / "c"), "--device_metrics", "--well_posed_image_score"])
    a = df_dev[["pixel AUC", "pixel AP"]].to_numpy(dtype=np.float64)[:3]
    b = df_m[["pixel AUC", "pixel AP"]].to_numpy(dtype=np.float64)[:3]
    assert np.abs(a - b).max() <= 0.011        # percent points: identical to 3 decimals of the [0,1] metric after rounding


def test_get_predictions_signature_and_shapes():
    """get_predictions keeps the reference's return values (test_last.py:153-158)."""
    import test_last as drv
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import weights
    from aaclip_b200.model import AdaptedCLIP, create_model
    S, B = 224, 2
    clip = create_model("ViT-L-14-336", S, pretrained=None, force_image_size=S, device="cuda")
    model = AdaptedCLIP(clip_model=clip, relu=False, iqm_hidden_size=512).to("cuda").eval()
    model.load_state_dict(weights.make_visual_state_dict(seed=3, img_size=S), strict=False)
    loader = [{"image": weights.make_images(3, B, S), "mask": torch.zeros(B, 1, S, S), "label": torch.tensor([0, 1]),
               "file_name": ["a", "b"], "class_name": ["01", "01"]}]
    with torch.no_grad():
        masks, labels, preds, preds_image, names = drv.get_predictions(model, weights.make_text_anchor(3).cuda(), loader,
                                                                       "cuda", S, dataset="BTAD")
    assert masks.shape == (B, 1, S, S) and preds.shape == (B, 1, S, S) and preds_image.shape == (B, 2)
    assert labels.tolist() == [0, 1] and names == ["a", "b"]
