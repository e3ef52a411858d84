"""Generates tests/golden/preprocess.npz by running the reference's transform (dataset/__init__.py:150-159) through
the libraries it calls -- PIL ``Image.resize(BICUBIC)`` and torchvision ``ToTensor`` / ``Normalize`` -- on seeded
synthetic RGB images. Small cases are stored in full; the BASELINE-size cases (1024x1024 and 700x900 -> 518) as a
SHA-256 of the float32 output bytes (a checksum the oracle and the CUDA path must reproduce bit for bit).
    python oracle/gen_golden_preprocess.py
"""
import hashlib
import os
import sys

import numpy as np
from PIL import Image
from torchvision import transforms

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pil_resample as pr  # noqa: E402

SMALL = [(97, 131, 56), (40, 30, 56), (56, 80, 56), (300, 200, 37), (33, 33, 70)]
LARGE = [(1024, 1024, 518), (700, 900, 518)]


def make_image(seed: int, H: int, W: int) -> np.ndarray:
    """Seeded synthetic RGB image: a smooth ramp plus noise plus saturated patches (exercises clip8 at 0 and 255)."""
    rng = np.random.default_rng(seed)
    base = np.add.outer(np.arange(H) * (255.0 / max(H - 1, 1)), np.arange(W) * (128.0 / max(W - 1, 1)))[..., None] * 0.6
    img = base + rng.normal(0, 40, (H, W, 3))
    img[: H // 5, : W // 4] = 255.0
    img[-(H // 6) - 1:, -(W // 3) - 1:] = 0.0
    img[H // 2, :] = rng.integers(0, 2, (W, 3)) * 255.0   # a 1-pixel line of full-swing pixels: ringing -> clipping
    return np.clip(img, 0, 255).astype(np.uint8)


def reference(img: np.ndarray, S: int) -> np.ndarray:
    tf = transforms.Compose([transforms.Resize((S, S), Image.BICUBIC), transforms.ToTensor(),
                             transforms.Normalize(mean=pr.CLIP_MEAN, std=pr.CLIP_STD)])
    return tf(Image.fromarray(img, "RGB")).numpy()


def main():
    out = {}
    for i, (H, W, S) in enumerate(SMALL):
        img = make_image(100 + i, H, W)
        ref = reference(img, S)
        assert np.array_equal(ref.view(np.uint32), pr.preprocess(img, S).view(np.uint32)), (H, W, S)
        out[f"small{i}_src"] = img
        out[f"small{i}_out"] = ref
    for i, (H, W, S) in enumerate(LARGE):
        img = make_image(200 + i, H, W)
        ref = reference(img, S)
        assert np.array_equal(ref.view(np.uint32), pr.preprocess(img, S).view(np.uint32)), (H, W, S)
        out[f"large{i}_shape"] = np.array([H, W, S], np.int64)
        out[f"large{i}_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(ref).tobytes()).digest(), np.uint8)
    import PIL
    import torchvision
    out["versions"] = np.array([f"Pillow {PIL.__version__}", f"torchvision {torchvision.__version__}"])
    path = os.path.join(ROOT, "tests", "golden", "preprocess.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
