"""TEST INFRASTRUCTURE -- runs the reference's own Testing/test.py, byte for byte as shipped (oracle/_ref/Testing, see
make_ref.py), against either model package:

    python oracle/run_reference_script.py --package dropin    -- --model td4-psp18 --output_path OUT ...
    python oracle/run_reference_script.py --package reference -- --model td4-psp18 --output_path OUT ...

--package dropin    : `from model import ...` resolves to tdnet_b200/dropin/model (the B200-native path)
--package reference : ... to oracle/_ref/Testing/model (the reference on torch/cuDNN, TF32 switched off so that it
                      computes in fp32 like its CPU path)
The script itself is not edited.  What is supplied from outside, because this image differs from the stack the reference
pins (requirements.txt: opencv-python 4.1.1, NumPy 1.x era): the `imageio` module (ref_shims/imageio.py, two calls backed
by OpenCV), no-op cv2.namedWindow / imshow / waitKey (headless box), and NumPy-1 integer semantics for
dataloader.decode_segmap -- it writes colour values up to 250 into an int8 copy of the label map, which NumPy 1 wrapped
silently (the wrap is undone by the script's final astype(uint8)) and NumPy 2 refuses with OverflowError; the label map
is handed over as int16 instead, which yields the same RGB image without the wrap.
The working directory is oracle/_ref/Testing, so the script's default --img_path ./data/vid1 is the reference's own clip.
"""
import argparse
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "Testing")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--package", choices=["dropin", "reference"], required=True)
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    rest = a.rest[1:] if a.rest[:1] == ["--"] else a.rest
    if not os.path.isfile(os.path.join(REF, "test.py")):
        sys.exit("oracle/_ref/Testing is missing: run `python oracle/make_ref.py` where /root/reference exists")
    paths = [os.path.join(ROOT, "oracle", "ref_shims")]
    if a.package == "dropin":
        paths += [os.path.join(ROOT, "tdnet_b200", "dropin"), ROOT]
    paths += [REF]                       # dataloader.py (and, for --package reference, model/)
    sys.path[:0] = paths

    import cv2
    cv2.namedWindow = lambda *args, **kw: None
    cv2.imshow = lambda *args, **kw: None
    cv2.waitKey = lambda *args, **kw: -1
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    import numpy as np
    import dataloader                      # the reference's, unmodified; test.py imports the same module object
    _decode = dataloader.cityscapesLoader.decode_segmap
    dataloader.cityscapesLoader.decode_segmap = lambda self, temp: _decode(self, np.asarray(temp).astype(np.int16))

    os.chdir(REF)
    sys.argv = ["test.py"] + rest
    runpy.run_path(os.path.join(REF, "test.py"), run_name="__main__")


if __name__ == "__main__":
    main()
