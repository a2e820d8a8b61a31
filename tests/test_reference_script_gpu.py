"""GPU: the reference's own Testing/test.py -- unmodified, from oracle/_ref/Testing (oracle/make_ref.py) -- driven
end to end on the drop-in package, on the reference's own 15 frames data/vid1/*.png at its own 769x1537, and the PNGs
it writes compared with the PNGs the same script writes with the reference's model package on the same box.

Both packages load the same synthetic checkpoint through the script's --_td4_psp18_path (strict=True in both,
td4_psp18.py:232-240); the reference runs on torch/cuDNN with TF32 switched off.  What is injected from outside is
listed in oracle/run_reference_script.py (an `imageio` stand-in and no-op cv2 GUI calls); test.py, dataloader.py and
the frames are byte-identical copies (oracle/_ref/MANIFEST.json).
"""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from common import make_weights

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "Testing")
# quarter-size label maps of 192 x 384 = 73 728 pixels: argmax near-ties may differ between two fp32 implementations
# (even fp32-vs-fp64 runs of the reference flip ~1e-5 of the pixels)
MAX_MISMATCH_PER_FRAME = 8


def run_script(package, ckpt, out_dir, model="td4-psp18"):
    os.makedirs(out_dir, exist_ok=True)
    key = {"td4-psp18": "--_td4_psp18_path", "td2-psp50": "--_td2_psp50_path", "psp101": "--_psp101_path"}[model]
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "run_reference_script.py"), "--package", package, "--",
           "--model", model, "--output_path", out_dir + os.sep, key, ckpt]
    env = dict(os.environ)
    env.pop("TDNET_B200_ENGINE", None)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "Loading pretrained model" in p.stdout
    m = re.search(r"Average\s+RunningTime/Latency=([0-9.]+) s", p.stdout)
    assert m, p.stdout[-1000:]
    per_frame = [float(x) for x in re.findall(r"Frame\s+\d+\s+RunningTime/Latency=([0-9.]+) s", p.stdout)]
    return float(m.group(1)), per_frame


@pytest.mark.skipif(not (os.path.isfile(os.path.join(REF, "test.py")) and os.path.isdir(os.path.join(REF, "data", "vid1"))),
                    reason="oracle/_ref/Testing absent (python oracle/make_ref.py needs /root/reference)")
@pytest.mark.parametrize("model,arch,backbone", [("td4-psp18", "td4_psp18", "resnet18"), ("td2-psp50", "td2_psp50", "resnet50"),
                                                 ("psp101", "pspnet", "resnet101")])
def test_unmodified_reference_script_on_dropin_matches_reference_package(tmp_path, model, arch, backbone):
    """All three `--model` choices of Testing/test.py:21-38 (td4-psp18, td2-psp50, psp101)."""
    import cv2
    ckpt = str(tmp_path / "ckpt.pkl")
    if arch == "pspnet":
        from common import make_pspnet_oracle
        torch.save(make_pspnet_oracle(backbone)[1], ckpt)
    else:
        torch.save(make_weights(arch, backbone, 97, 193), ckpt)   # LayerNorm([97,193]) as the reference constructs it
    lat_ref, frames_ref = run_script("reference", ckpt, str(tmp_path / "reference"), model)
    lat_ours, frames_ours = run_script("dropin", ckpt, str(tmp_path / "dropin"), model)
    names = sorted(os.listdir(tmp_path / "reference" / "vid1"))
    assert len(names) == 15 and names == sorted(os.listdir(tmp_path / "dropin" / "vid1"))
    mismatches = []
    for n in names:
        a = cv2.imread(str(tmp_path / "reference" / "vid1" / n))
        b = cv2.imread(str(tmp_path / "dropin" / "vid1" / n))
        assert a is not None and b is not None and a.shape == b.shape == (769 // 4, 1537 // 4, 3), (n, a.shape)
        assert len(np.unique(a.reshape(-1, 3), axis=0)) > 3      # a real segmentation, not a constant map
        mismatches.append(int((a != b).any(axis=2).sum()))
    rec = {"model": model, "frames": len(names), "pixels_per_frame": 192 * 384, "mismatching_pixels": mismatches,
           "latency_s_reference_package_on_cudnn_fp32": lat_ref, "latency_s_dropin": lat_ours,
           "per_frame_s_dropin": frames_ours, "per_frame_s_reference": frames_ref}
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "reference_script_run.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass
    print(rec)
    assert max(mismatches) <= MAX_MISMATCH_PER_FRAME, mismatches
