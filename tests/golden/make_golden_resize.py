"""Generate tests/golden/resize_cases.npz: outputs of cv2.resize itself (the reference's resampling dependency,
Testing/dataloader.py:63 and Testing/test.py:64) on seeded random inputs.  Run where cv2 is installed:
    python tests/golden/make_golden_resize.py
Nothing here calls oracle/ or the CUDA path."""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LINEAR = [((37, 53), (97, 161)), ((97, 161), (37, 53)), ((64, 96), (49, 97)), ((5, 7), (33, 21)), ((128, 256), (128, 256))]
NEAREST = [((97, 161), (24, 40)), ((64, 96), (16, 24)), ((33, 21), (50, 9))]


def main():
    rec = {"cv2_version": np.array(cv2.__version__)}
    rng = np.random.default_rng(2024)
    for i, ((h, w), (H, W)) in enumerate(LINEAR):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        rec[f"lin{i}_in"], rec[f"lin{i}_out"] = img, cv2.resize(img, (W, H))
    # the deployment shape: only a band of rows / columns of the 1024x2048 -> 769x1537 result is stored
    img = rng.integers(0, 256, (1024, 2048, 3), dtype=np.uint8)
    out = cv2.resize(img, (1537, 769))
    rec["big_seed"] = np.array(77)
    big = np.random.default_rng(77).integers(0, 256, (1024, 2048, 3), dtype=np.uint8)
    out = cv2.resize(big, (1537, 769))
    rec["big_rows"] = out[[0, 1, 384, 767, 768]]
    rec["big_cols"] = out[:, [0, 1, 768, 1535, 1536]]
    for i, ((h, w), (H, W)) in enumerate(NEAREST):
        lab = rng.integers(0, 19, (h, w), dtype=np.int8)
        rec[f"nn{i}_in"], rec[f"nn{i}_out"] = lab, cv2.resize(lab, (W, H), interpolation=cv2.INTER_NEAREST)
    path = os.path.join(HERE, "resize_cases.npz")
    np.savez_compressed(path, **rec)
    print(path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
