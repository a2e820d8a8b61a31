"""Shared helpers for the parity tests: golden cases, oracle construction, error metrics."""
import os

import numpy as np
import torch

from oracle.td2fa_oracle import TD2FAOracle, fa_feature_hw, td2fa_state_dict_template
from oracle.tdnet_oracle import PSPNetOracle, TDOracle, state_dict_template
from tdnet_b200.synth import synth_state_dict

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CH_STRIDE = 4  # tests/golden/make_golden.py stores every 4th channel of the wide taps

# name -> (arch, backbone)
GOLDEN_CASES = {
    "td4_r18_97x161": ("td4_psp18", "resnet18"),
    "td4_r18_128x256": ("td4_psp18", "resnet18"),
    "td2_r50_64x128": ("td2_psp50", "resnet50"),
    "td2_r34_80x112": ("td2_psp50", "resnet34"),
    "td4_r50_64x64_n2": ("td4_psp18", "resnet50"),
    "td4_r18_769x1537_chk": ("td4_psp18", "resnet18"),
}
# single-path PSPNet comparison model (Testing/model/pspnet/pspnet.py): name -> backbone
PSPNET_GOLDEN_CASES = {"psp_r101_64x96_n2": "resnet101", "psp_r18_97x161": "resnet18"}
# TD2-FANet (Training/ptsemseg/models/td2_fanet/td2_fa.py; tests/golden/make_golden_fanet.py): name -> backbone
FANET_GOLDEN_CASES = {"td2fa_r18_128x192": "resnet18", "td2fa_r34_97x161_n2": "resnet34",
                      "td2fa_r50_64x96": "resnet50"}


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    H, W, batch, n_frames, h8, w8 = (int(v) for v in g["meta"])
    return g, dict(H=H, W=W, batch=batch, n_frames=n_frames, h8=h8, w8=w8)


def feat_hw(h, w):
    for _ in range(3):
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    return h, w


def make_weights(arch, backbone, h8, w8, seed=0):
    return synth_state_dict(state_dict_template(arch, backbone, ln_shape=(h8, w8)), seed=seed)


def make_oracle(arch, backbone, H, W, seed=0):
    h8, w8 = feat_hw(H, W)
    sd = make_weights(arch, backbone, h8, w8, seed)
    return TDOracle(arch, sd, backbone), sd


def make_pspnet_oracle(backbone, seed=0):
    sd = synth_state_dict(state_dict_template("pspnet", backbone), seed=seed)
    return PSPNetOracle(sd, backbone), sd


def make_fanet_oracle(backbone, H, W, seed=0):
    h4, w4 = fa_feature_hw(H, W)
    sd = synth_state_dict(td2fa_state_dict_template(backbone, ln_shape=(h4, w4)), seed=seed)
    return TD2FAOracle(sd, backbone), sd


def max_abs(a, b):
    return float((torch.as_tensor(a).double() - torch.as_tensor(b).double()).abs().max())


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def argmax_report(test_logits, ref_logits, err):
    """Argmax agreement the way SURVEY.md 8(c) states the gate: labels must be equal on every pixel
    whose reference top-1/top-2 margin exceeds 2*err (err = measured max-abs logit error); pixels
    inside that band are near-ties that even fp64-vs-fp32 runs of the reference flip."""
    ref = torch.as_tensor(ref_logits)
    tst = torch.as_tensor(test_logits)
    top2 = ref.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    same = tst.argmax(1) == ref.argmax(1)
    decided = margin > 2.0 * err
    return dict(mismatch_total=int((~same).sum()), mismatch_decided=int((~same & decided).sum()),
                near_ties=int((~decided).sum()), pixels=int(same.numel()))


def record(name, **metrics):
    """Append measured parity numbers to gpurun_out/parity_metrics.jsonl (copied into profiles/ and cited
    in DESIGN.md); never fails a test."""
    import json
    try:
        out = os.path.join(os.path.dirname(GOLDEN_DIR), "..", "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_metrics.jsonl"), "a") as f:
            f.write(json.dumps(dict(name=name, **metrics)) + "\n")
    except OSError:
        pass
