"""Generate tests/golden/td2fa_*.npz by running the UNMODIFIED reference TD2-FANet on CPU (build container only).

    python tests/golden/make_golden_fanet.py        # needs /root/reference (absent on the GPU box)

The model is Training/ptsemseg/models/td2_fanet/td2_fa.py (class td2_fa) in eval mode.  Nothing of its code is
changed; three import-time obstacles of that tree are handled from the outside:
  * `ptsemseg/models/__init__.py:4` imports SyncBatchNorm from the `encoding` package (PyTorch-Encoding, not
    installed, CUDA-only).  A stub `encoding.nn` module is registered whose SyncBatchNorm is the reference's own
    single-process stand-in for it, Testing/model/pspnet/td4_psp18.py:11-24 (`BatchNorm2d(activation=...)`:
    eval-mode BatchNorm + optional LeakyReLU(0.01)); that class is what is passed as `norm_layer`.
  * `td2_fa.__init__` stops in `pdb.set_trace()` (td2_fa.py:81): pdb.set_trace is replaced by a no-op.
  * `resnet18(pretrained=True)` downloads ImageNet weights (resnet.py:161-165): model_zoo.load_url returns {} (the
    loop in ResNet.init_weight then copies nothing); all weights are overwritten by tdnet_b200.synth anyway.
The hard-coded LayerNorm([96, 192]) (td2_fa.py:71-72) is re-created at the feat4 size of the test input, as for
the other models.  Outputs of the reference are stored; nothing here calls the CUDA path, and oracle/ is only used
for the state-dict key/shape cross-check.
"""
import os
import pdb
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_TRAIN = "/root/reference/Training"
REF_TEST = "/root/reference/Testing"


def import_reference():
    sys.path.insert(0, REF_TEST)
    from model import td4_psp18 as ref_td4            # Testing/model/__init__.py:1 -> pspnet/td4_psp18.py
    norm_layer = ref_td4.BatchNorm2d

    enc = types.ModuleType("encoding")
    enc_nn = types.ModuleType("encoding.nn")
    enc_nn.SyncBatchNorm = norm_layer
    enc.nn = enc_nn
    sys.modules["encoding"], sys.modules["encoding.nn"] = enc, enc_nn
    pdb.set_trace = lambda *a, **k: None
    import torch.utils.model_zoo as model_zoo
    model_zoo.load_url = lambda *a, **k: {}
    sys.path.insert(0, REF_TRAIN)
    from ptsemseg.models.td2_fanet.td2_fa import td2_fa
    return td2_fa, norm_layer


from oracle.td2fa_oracle import fa_feature_hw, td2fa_state_dict_template  # noqa: E402  (key/shape check only)
from tdnet_b200.synth import synth_clip, synth_state_dict  # noqa: E402

# name, backbone, H, W, batch, calls (call i: frames (i, i+1), pos_id = i % 2)
CASES = [
    ("td2fa_r18_128x192", "resnet18", 128, 192, 1, 3),     # feat4 16x24 ... feat32 2x3
    ("td2fa_r34_97x161_n2", "resnet34", 97, 161, 2, 2),    # ragged maps 13x21 / 7x11 / 4x6 / 2x3, batch 2
    ("td2fa_r50_64x96", "resnet50", 64, 96, 1, 2),         # Bottleneck backbone, 2048-channel feat32
]
# the only size the UNPATCHED reference accepts: LayerNorm([96, 192]) (td2_fa.py:71-72) = a 768x1536 input; checksums only
NATIVE = ("td2fa_r18_768x1536_chk", "resnet18", 768, 1536, 1, 2)
CH_STRIDE = 4


def run_native(td2_fa, norm_layer, name, backbone, H, W, batch, calls):
    """No patch at all: the hard-coded LayerNorm([96, 192]) is used as constructed."""
    torch.manual_seed(0)
    net = td2_fa(nclass=19, backbone=backbone, norm_layer=norm_layer, path_num=2).eval()
    assert fa_feature_hw(H, W) == (96, 192)
    net.load_state_dict(synth_state_dict(td2fa_state_dict_template(backbone), seed=0), strict=True)
    cur = {}
    for idx in (1, 2):
        getattr(net, f"head{idx}").register_forward_hook(
            lambda _m, _i, out, k=f"head{idx}": cur.setdefault(k, []).append(out))
    frames = synth_clip(calls + 1, H, W, batch=batch, clip_id=0)
    rec = {}
    with torch.no_grad():
        for i in range(calls):
            cur.clear()
            out = net([frames[i], frames[i + 1]], pos_id=i % 2)
            head = cur[f"head{i % 2 + 1}"][0]
            rec[f"head_mean_{i}"] = np.float64(head.double().mean().item())
            rec[f"head_absmean_{i}"] = np.float64(head.double().abs().mean().item())
            rec[f"head_sub_{i}"] = head[:, :, ::8, ::16].numpy().copy()
            rec[f"logits_sub_{i}"] = out[:, :, ::64, ::128].numpy().copy()
    rec["meta"] = np.array([H, W, batch, calls, 96, 192], dtype=np.int64)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB, {len(rec)} arrays")


def run_case(td2_fa, norm_layer, name, backbone, H, W, batch, calls):
    torch.manual_seed(0)
    net = td2_fa(nclass=19, backbone=backbone, norm_layer=norm_layer, path_num=2).eval()
    ref_sd = net.state_dict()
    h4, w4 = fa_feature_hw(H, W)
    tmpl = td2fa_state_dict_template(backbone)
    assert set(tmpl) == set(ref_sd), sorted(set(tmpl) ^ set(ref_sd))[:8]
    for k in ref_sd:
        assert tuple(tmpl[k].shape) == tuple(ref_sd[k].shape) and tmpl[k].dtype == ref_sd[k].dtype, k
    net.layer_norm1.ln = nn.LayerNorm([h4, w4])
    net.layer_norm2.ln = nn.LayerNorm([h4, w4])
    tmpl = td2fa_state_dict_template(backbone, ln_shape=(h4, w4))
    net.load_state_dict(synth_state_dict(tmpl, seed=0), strict=True)

    cur = {}

    def hook(key):
        def f(_m, _i, out):
            cur.setdefault(key, []).append(out)
        return f

    for cname, m in net.named_children():
        if cname.startswith(("pretrained", "ffm_", "enc", "atn", "layer_norm", "head")):
            m.register_forward_hook(hook(cname))
    frames = synth_clip(calls + 1, H, W, batch=batch, clip_id=0)
    rec = {}
    with torch.no_grad():
        for i in range(calls):
            cur.clear()
            pos = i % 2
            out = net([frames[i], frames[i + 1]], pos_id=pos)
            a, b = (1, 2) if pos == 0 else (2, 1)            # a: sub-network of the current frame
            if i == calls - 1:
                rec[f"logits_{i}"] = out.numpy().copy()
            rec[f"head_{i}"] = cur[f"head{a}"][0].numpy().copy()     # first call: head(LN(atn + v)); second: out_sub
            if i == calls - 1:
                feats = cur[f"pretrained{a}"][0]
                rec["tap_feat4"] = feats[0].numpy().copy()
                rec["tap_feat32"] = feats[3][:, ::CH_STRIDE].numpy().copy()
                rec["tap_up32"] = cur[f"ffm_32_{a}"][0][:, ::CH_STRIDE].numpy().copy()
                up16, sm16 = cur[f"ffm_16_{a}"][0]
                rec["tap_up16"], rec["tap_sm16"] = up16[:, ::CH_STRIDE].numpy().copy(), sm16.numpy().copy()
                rec["tap_up8"] = cur[f"ffm_8_{a}"][0].numpy().copy()
                rec["tap_sm4"] = cur[f"ffm_4_{a}"][0].numpy().copy()
                q, v = cur[f"enc{a}"][0]
                k_, v_ = cur[f"enc{b}"][0]
                rec["tap_q"], rec["tap_v"] = q.numpy().copy(), v[:, ::CH_STRIDE].numpy().copy()
                rec["tap_k_sub"], rec["tap_v_sub"] = k_.numpy().copy(), v_.numpy().copy()
                rec["tap_atn"] = cur[f"atn{a}"][0][:, ::CH_STRIDE].numpy().copy()
                rec["tap_normed"] = cur[f"layer_norm{a}"][0][:, ::CH_STRIDE].numpy().copy()
    rec["meta"] = np.array([H, W, batch, calls, h4, w4], dtype=np.int64)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB, {len(rec)} arrays")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    td2_fa, norm_layer = import_reference()
    only = sys.argv[1:]
    for case in CASES:
        if not only or case[0] in only:
            run_case(td2_fa, norm_layer, *case)
    if not only or NATIVE[0] in only:
        run_native(td2_fa, norm_layer, *NATIVE)
