"""TEST INFRASTRUCTURE -- stand-in for the `imageio` package (not installed in this image) for running the unmodified
Testing/test.py and Testing/dataloader.py: the two calls they make, backed by OpenCV.
  imageio.imread(path)        (dataloader.py:61)  -> uint8 HxWx3 RGB
  imageio.imwrite(path, arr)  (test.py:72)        <- uint8 HxWx3 RGB
"""
import cv2 as _cv2


def imread(path):
    img = _cv2.imread(path, _cv2.IMREAD_COLOR)
    if img is None:
        raise FileNotFoundError(path)
    return _cv2.cvtColor(img, _cv2.COLOR_BGR2RGB)


def imwrite(path, arr):
    if arr.ndim == 3:
        arr = _cv2.cvtColor(arr, _cv2.COLOR_RGB2BGR)
    if not _cv2.imwrite(path, arr):
        raise OSError(f"could not write {path}")
