"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU (build container only).

    python tests/golden/make_golden.py            # needs /root/reference (absent on the GPU box)

For every case: import /root/reference/Testing/model, build the reference module, patch only the
hard-coded LayerNorm([97,193]) (td4_psp18.py:107-110) to the feature-map size, load the synthetic
state dict from tdnet_b200.synth, feed the seeded clip, and record with forward hooks what the
reference itself produced.  Nothing in here calls oracle/ or the CUDA path; the outputs are the pin
for both.  Also asserts that oracle.state_dict_template() reproduces the reference's state-dict
keys and shapes exactly.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/Testing")

from model import pspnet, td2_psp50, td4_psp18  # noqa: E402  (the reference)

from oracle.tdnet_oracle import state_dict_template  # noqa: E402  (key/shape check only)
from tdnet_b200.synth import synth_clip, synth_state_dict  # noqa: E402

# name, arch, backbone, H, W, batch, frames, full-res logits kept for these frame indices
CASES = [
    ("td4_r18_97x161", "td4_psp18", "resnet18", 97, 161, 1, 9, (8,)),    # ragged 13x21 map, P'=24
    ("td4_r18_128x256", "td4_psp18", "resnet18", 128, 256, 1, 6, (5,)),    # 16x32 map, P'=32
    ("td2_r50_64x128", "td2_psp50", "resnet50", 64, 128, 1, 4, (3,)),      # frame pairs x2
    ("td2_r34_80x112", "td2_psp50", "resnet34", 80, 112, 1, 3, (2,)),      # 'bise34' stand-in, 10x14
    ("td4_r50_64x64_n2", "td4_psp18", "resnet50", 64, 64, 2, 5, (4,)),     # batch 2 streams
    ("td4_r18_769x1537_chk", "td4_psp18", "resnet18", 769, 1537, 1, 5, ()),  # reference-native size
    # single-path PSPNet comparison model (pspnet.py); batch 2: only x[-1:] is segmented
    ("psp_r101_64x96_n2", "pspnet", "resnet101", 64, 96, 2, 2, (0, 1)),
    ("psp_r18_97x161", "pspnet", "resnet18", 97, 161, 1, 2, (1,)),
]


CH_STRIDE = 4


def feat_hw(h, w):
    for _ in range(3):
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    return h, w


def run_pspnet_case(name, backbone, H, W, batch, n_frames, keep):
    """pspnet.pspnet(nclass=19, backbone=...) (pspnet.py:31-89): stateless, segments x[-1:] only."""
    torch.manual_seed(0)
    net = pspnet.pspnet(nclass=19, backbone=backbone).eval()
    ref_sd = net.state_dict()
    tmpl = state_dict_template("pspnet", backbone)
    assert set(tmpl) == set(ref_sd), sorted(set(tmpl) ^ set(ref_sd))[:8]
    for k in ref_sd:
        assert tuple(tmpl[k].shape) == tuple(ref_sd[k].shape) and tmpl[k].dtype == ref_sd[k].dtype, k
    net.load_state_dict(synth_state_dict(tmpl, seed=0), strict=True)
    cur = {}
    net.head.conv5[0].register_forward_hook(lambda m, i, o: cur.__setitem__("z", o))
    net.head.register_forward_hook(lambda m, i, o: cur.__setitem__("head", o))
    rec = {}
    with torch.no_grad():
        for i, f in enumerate(synth_clip(n_frames, H, W, batch=batch, clip_id=0)):
            cur.clear()
            # pspnet.forward runs the backbone piecewise (no pretrained.forward hook fires): tap layer4 instead
            hk = net.pretrained.layer4.register_forward_hook(lambda m, i_, o: cur.__setitem__("c4", o))
            out = net(f, pos_id=i % 4)
            hk.remove()
            rec[f"head_{i}"] = cur["head"].numpy().copy()
            if i in keep:
                rec[f"logits_{i}"] = out.numpy().copy()
            if i == n_frames - 1:
                rec["tap_c4"] = cur["c4"][:, ::CH_STRIDE].numpy().copy()
                rec["tap_z"] = cur["z"][:, ::CH_STRIDE].numpy().copy()
    h8, w8 = feat_hw(H, W)
    rec["meta"] = np.array([H, W, batch, n_frames, h8, w8], dtype=np.int64)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB, {len(rec)} arrays")


def run_case(name, arch, backbone, H, W, batch, n_frames, keep):
    if arch == "pspnet":
        return run_pspnet_case(name, backbone, H, W, batch, n_frames, keep)
    torch.manual_seed(0)
    mod = td4_psp18.td4_psp18 if arch == "td4_psp18" else td2_psp50.td2_psp50
    paths = 4 if arch == "td4_psp18" else 2
    net = mod(nclass=19, path_num=paths, backbone=backbone).eval()
    ref_sd = net.state_dict()
    tmpl = state_dict_template(arch, backbone)
    assert set(tmpl) == set(ref_sd), sorted(set(tmpl) ^ set(ref_sd))[:8]
    for k in ref_sd:
        assert tuple(tmpl[k].shape) == tuple(ref_sd[k].shape) and tmpl[k].dtype == ref_sd[k].dtype, k
    h8, w8 = feat_hw(H, W)
    if (h8, w8) != (97, 193):
        for cname, m in net.named_children():
            if cname.startswith("layer_norm"):
                m.ln = nn.LayerNorm([h8, w8])
    tmpl = state_dict_template(arch, backbone, ln_shape=(h8, w8))
    net.load_state_dict(synth_state_dict(tmpl, seed=0), strict=True)

    rec = {}
    cur = {}

    def hook(key):
        def f(_m, _i, out):
            cur.setdefault(key, []).append(out)
        return f

    for cname, m in net.named_children():
        if cname.startswith(("pretrained", "psp", "enc", "atn", "layer_norm", "head")):
            m.register_forward_hook(hook(cname))
    frames = synth_clip(n_frames, H, W, batch=batch, clip_id=0)
    big = name.endswith("_chk")
    with torch.no_grad():
        for i, f in enumerate(frames):
            cur.clear()
            pos = i % paths
            out = net(f, pos_id=pos)
            p = pos + 1
            head = cur[f"head{p}"][0]
            if big:  # only checksums for the big native-size case
                rec[f"head_mean_{i}"] = np.float64(head.double().mean().item())
                rec[f"head_absmean_{i}"] = np.float64(head.double().abs().mean().item())
                rec[f"head_sub_{i}"] = head[:, :, ::8, ::16].numpy().copy()
                rec[f"logits_sub_{i}"] = out[:, :, ::64, ::128].numpy().copy()
                if i == n_frames - 1:
                    # the reference's arg-max label map of the last (steady-state) frame at its native size
                    # (Testing/test.py:61) and the pixels whose top-1 / top-2 margin is below 1e-3 (near-ties), bit-packed
                    top2 = out.topk(2, dim=1).values
                    rec["argmax_last"] = out.argmax(1).to(torch.uint8).numpy().copy()
                    rec["near_tie_last"] = np.packbits(((top2[:, 0] - top2[:, 1]) < 1e-3).numpy())
                continue
            rec[f"head_{i}"] = head.numpy().copy()
            if i in keep:
                rec[f"logits_{i}"] = out.numpy().copy()
            if i == n_frames - 1:  # per-stage taps of the last (steady-state) frame
                # wide maps are stored for every CH_STRIDE-th channel only (fixture size)
                rec["tap_c4"] = cur[f"pretrained{p}"][0][:, ::CH_STRIDE].numpy().copy()
                rec["tap_z"] = cur[f"psp{p}"][0][:, ::CH_STRIDE].numpy().copy()
                q_cur, v_cur = cur[f"enc{p}"][0]
                q_sub, k_sub, v_sub = cur[f"enc{p}"][1]
                rec["tap_q_cur"], rec["tap_v_cur"] = q_cur.numpy().copy(), v_cur[:, ::CH_STRIDE].numpy().copy()
                rec["tap_q_sub"], rec["tap_k_sub"] = q_sub.numpy().copy(), k_sub.numpy().copy()
                rec["tap_v_sub"] = v_sub.numpy().copy()
                rec["tap_normed"] = cur[f"layer_norm{p}"][0][:, ::CH_STRIDE].numpy().copy()
                hops = [k for k in cur if k.startswith("atn")]
                # hooks fire in call order; the last hop is the propagated feature map
                rec["tap_v_prop"] = cur[hops[-1]][0][:, ::CH_STRIDE].numpy().copy()
                for j, k in enumerate(hops[:-1]):
                    rec[f"tap_hop{j}"] = cur[k][0].numpy().copy()
    rec["meta"] = np.array([H, W, batch, n_frames, h8, w8], dtype=np.int64)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB, {len(rec)} arrays")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    only = sys.argv[1:]
    for case in CASES:
        if not only or case[0] in only:
            run_case(*case)
