"""Reference import path ``dataset`` (test_last.py:16: ``from dataset import get_dataset, DOMAINS``).

Only the EVALUATION side of the reference's data layer is provided (dataset/__init__.py:124-196,245-256): per-class test
sets read from the same ``dataset/metadata/<name>/full-shot.jsonl`` records ({"image_path", "label", "mask_path",
"class_name"}), yielding the same sample dictionaries. The image transform (bicubic resize to img_size, ToTensor,
CLIP mean / std) is the reference's; on the GPU the same arithmetic is available bit-exactly through
``aaclip_b200.preprocess.ImagePreprocessor`` (aaclip_preprocess_u8). The training datasets and their augmentation
pipeline are outside this build.
"""
import json
import os

import model  # noqa: F401
import torch
from aaclip_b200.constants import CLASS_NAMES, DATA_PATH, DOMAINS  # noqa: F401
from torch.utils.data import Dataset

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
METADATA_ROOT = os.environ.get("AACLIP_METADATA_ROOT", "./dataset/metadata")


class SingleClassTestSet(Dataset):
    """All records of one class; items are {"image" [3,S,S], "mask" [1,S,S], "label", "file_name", "class_name"}."""

    def __init__(self, data_path: str, meta_path: str, img_size: int, class_name: str, logger=None,
                 raw_uint8: bool = False):
        from PIL import Image
        from torchvision import transforms
        self.data_path, self.img_size, self.raw_uint8 = data_path, img_size, raw_uint8
        with open(meta_path) as f:
            rows = [json.loads(line) for line in f if line.strip()]
        self.meta = [r for r in rows if r["class_name"] == class_name]
        self._open = Image.open
        self.to_image = transforms.Compose([
            transforms.Resize((img_size, img_size), Image.BICUBIC), transforms.ToTensor(),
            transforms.Normalize(mean=CLIP_MEAN, std=CLIP_STD)])
        self.to_mask = transforms.Compose([transforms.Resize((img_size, img_size), Image.NEAREST),
                                           transforms.ToTensor()])
        if logger:
            logger.info(f"Class name: {class_name}")
            logger.info(f"Sample number: {len(self.meta)}")

    def __len__(self):
        return len(self.meta)

    def __getitem__(self, i):
        m = self.meta[i]
        rgb = self._open(os.path.join(self.data_path, m["image_path"])).convert("RGB")
        if m["label"]:
            mask = self.to_mask(self._open(os.path.join(self.data_path, m["mask_path"])).convert("L"))
            mask = (mask != 0).float()
        else:
            mask = torch.zeros(1, self.img_size, self.img_size)
        item = {"mask": mask, "label": m["label"], "file_name": m["image_path"], "class_name": m["class_name"]}
        if self.raw_uint8:
            import numpy as np
            item["image_u8"] = torch.from_numpy(np.asarray(rgb).copy())      # [H,W,3]; resized + normalised on the GPU
        else:
            item["image"] = self.to_image(rgb)
        return item


def get_dataset(dataset_name: str, img_size: int, training_mode, shot: int = -1, stage: str = "train", logger=None,
                data_path: str = None, metadata_root: str = None, raw_uint8: bool = False):
    """stage "test" / "visualize" -> {class_name: Dataset} (dataset/__init__.py:245-268). `data_path` / `metadata_root`
    override DATA_PATH[dataset_name] and ./dataset/metadata; `raw_uint8` makes the items carry the decoded image as
    uint8 [H,W,3] ("image_u8") instead of the transformed tensor, for the device-side transform (aaclip_preprocess_u8)."""
    if stage not in ("test", "visualize"):
        raise NotImplementedError("only the evaluation datasets (stage='test') are provided; the training datasets "
                                  "and their augmentations are outside the inference hot path")
    assert dataset_name in DATA_PATH, f"Dataset {dataset_name} not found; available datasets: {list(DATA_PATH.keys())}"
    meta_path = os.path.join(metadata_root or METADATA_ROOT, dataset_name, "full-shot.jsonl")
    return {c: SingleClassTestSet(data_path or DATA_PATH[dataset_name], meta_path, img_size, c,
                                  logger if stage == "test" else None, raw_uint8=raw_uint8)
            for c in CLASS_NAMES[dataset_name]}
