"""CPU: the oracle restatement vs. what the unmodified reference produced (tests/golden/*.npz)."""
import numpy as np
import pytest

from common import (CH_STRIDE, FANET_GOLDEN_CASES, GOLDEN_CASES, PSPNET_GOLDEN_CASES, load_golden, make_fanet_oracle,
                    make_oracle, make_pspnet_oracle, max_abs)
from oracle.tdnet_oracle import stage_plan, state_dict_template
from tdnet_b200.synth import synth_clip

# Same machine + same torch primitives in the same order: expected bit-equal; the tolerance only
# absorbs oneDNN choosing a different kernel when the thread count differs from generation time.
TOL = 2e-5


@pytest.mark.parametrize("name", [n for n in GOLDEN_CASES if not n.endswith("_chk")])
def test_oracle_matches_reference_outputs(name):
    arch, backbone = GOLDEN_CASES[name]
    g, m = load_golden(name)
    oracle, _ = make_oracle(arch, backbone, m["H"], m["W"])
    frames = synth_clip(m["n_frames"], m["H"], m["W"], batch=m["batch"], clip_id=0)
    paths = oracle.paths
    for i, f in enumerate(frames):
        out = oracle(f, pos_id=i % paths)
        assert max_abs(oracle.taps["head"], g[f"head_{i}"]) <= TOL, (name, i)
        if f"logits_{i}" in g:
            assert out.shape == g[f"logits_{i}"].shape
            assert max_abs(out, g[f"logits_{i}"]) <= TOL
            assert (out.argmax(1).numpy() == g[f"logits_{i}"].argmax(1)).mean() > 0.9999
    t = oracle.taps
    s = CH_STRIDE
    assert max_abs(t["c4"][:, ::s], g["tap_c4"]) <= 5 * TOL
    assert max_abs(t["z"][:, ::s], g["tap_z"]) <= 5 * TOL
    assert max_abs(t["q_cur"], g["tap_q_cur"]) <= 5 * TOL
    assert max_abs(t["v_cur"][:, ::s], g["tap_v_cur"]) <= 5 * TOL
    assert max_abs(t["q_sub"], g["tap_q_sub"]) <= 5 * TOL
    assert max_abs(t["k_sub"], g["tap_k_sub"]) <= 5 * TOL
    assert max_abs(t["v_sub"], g["tap_v_sub"]) <= 5 * TOL
    assert max_abs(t["v_prop"][:, ::s], g["tap_v_prop"]) <= 5 * TOL
    assert max_abs(t["normed"][:, ::s], g["tap_normed"]) <= 5 * TOL
    if arch == "td4_psp18":
        assert max_abs(t["v2"], g["tap_hop0"]) <= 5 * TOL
        assert max_abs(t["v3"], g["tap_hop1"]) <= 5 * TOL
    # FIFO semantics of buffer_contral (td4_psp18.py:123-134 / td2_psp50.py:98-109)
    assert len(oracle.Q_queue) == len(oracle.K_queue) == len(oracle.V_queue) == oracle.depth


@pytest.mark.parametrize("name", sorted(PSPNET_GOLDEN_CASES))
def test_pspnet_oracle_matches_reference_outputs(name):
    """PSPNetOracle vs the unmodified pspnet.pspnet (pspnet.py:31-89) run on CPU; the batch-2 case pins
    `x = x[-1:]` (only the last image of the batch is segmented, :74)."""
    g, m = load_golden(name)
    oracle, _ = make_pspnet_oracle(PSPNET_GOLDEN_CASES[name])
    frames = synth_clip(m["n_frames"], m["H"], m["W"], batch=m["batch"], clip_id=0)
    scale = 1.0
    for i, f in enumerate(frames):
        out = oracle(f, pos_id=i % 4)
        scale = max(1.0, float(np.abs(g[f"head_{i}"]).max()))
        assert max_abs(oracle.taps["head"], g[f"head_{i}"]) <= TOL * scale, (name, i)
        if f"logits_{i}" in g:
            assert out.shape == g[f"logits_{i}"].shape == (1, 19, m["H"], m["W"])
            assert max_abs(out, g[f"logits_{i}"]) <= TOL * scale
    s = CH_STRIDE
    assert max_abs(oracle.taps["c4"][:, ::s], g["tap_c4"]) <= 5 * TOL * scale
    assert max_abs(oracle.taps["z"][:, ::s], g["tap_z"]) <= 5 * TOL * scale


@pytest.mark.parametrize("name", sorted(FANET_GOLDEN_CASES))
def test_fanet_oracle_matches_reference_outputs(name):
    """TD2FAOracle vs the unmodified td2_fa (Training/ptsemseg/models/td2_fanet/td2_fa.py) run on CPU: call i feeds
    the frame pair (i, i+1) with pos_id = i % 2; every FAModule output of the current frame's sub-network is pinned."""
    g, m = load_golden(name)
    oracle, _ = make_fanet_oracle(FANET_GOLDEN_CASES[name], m["H"], m["W"])
    calls = m["n_frames"]
    frames = synth_clip(calls + 1, m["H"], m["W"], batch=m["batch"], clip_id=0)
    for i in range(calls):
        out = oracle([frames[i], frames[i + 1]], pos_id=i % 2)
        scale = max(1.0, float(np.abs(g[f"head_{i}"]).max()))
        assert max_abs(oracle.taps["head"], g[f"head_{i}"]) <= TOL * scale, (name, i)
        if f"logits_{i}" in g:
            assert out.shape == g[f"logits_{i}"].shape == (m["batch"], 19, m["H"], m["W"])
            assert max_abs(out, g[f"logits_{i}"]) <= TOL * scale
    t, s = oracle.taps, CH_STRIDE
    a = 1 if (calls - 1) % 2 == 0 else 2
    for key, tap, stride in (("feat4", f"feat4_{a}", 1), ("feat32", f"feat32_{a}", s), ("up32", f"up32_{a}", s),
                             ("up16", f"up16_{a}", s), ("sm16", f"sm16_{a}", 1), ("up8", f"up8_{a}", 1),
                             ("sm4", f"sm4_{a}", 1), ("q", "q", 1), ("v", "v", s), ("k_sub", "k_sub", 1),
                             ("v_sub", "v_sub", 1), ("atn", "atn", s), ("normed", "normed", s)):
        ref = g["tap_" + key]
        got = t[tap][:, ::stride] if stride > 1 else t[tap]
        assert tuple(got.shape) == ref.shape, key
        assert max_abs(got, ref) <= 5 * TOL * max(1.0, float(np.abs(ref).max())), key
    # `up` is a 1x1 conv with padding 1 (td2_fa.py:348): the map it returns is 2 pixels larger than its input
    h32, w32 = g["tap_feat32"].shape[2:]
    assert g["tap_up32"].shape[2:] == (h32 + 2, w32 + 2)


def test_fanet_oracle_native_size_checksums():
    """768x1536: the only size the unpatched td2_fa accepts (hard-coded LayerNorm([96, 192]), td2_fa.py:71-72); the
    fixture was produced without touching the module at all."""
    g, m = load_golden("td2fa_r18_768x1536_chk")
    assert (m["h8"], m["w8"]) == (96, 192)
    oracle, _ = make_fanet_oracle("resnet18", m["H"], m["W"])
    frames = synth_clip(m["n_frames"] + 1, m["H"], m["W"], batch=m["batch"], clip_id=0)
    for i in range(m["n_frames"]):
        out = oracle([frames[i], frames[i + 1]], pos_id=i % 2)
        head = oracle.taps["head"]
        assert max_abs(head[:, :, ::8, ::16], g[f"head_sub_{i}"]) <= 2 * TOL
        assert max_abs(out[:, :, ::64, ::128], g[f"logits_sub_{i}"]) <= 2 * TOL
        assert abs(head.double().mean().item() - float(g[f"head_mean_{i}"])) <= 1e-6
    assert oracle.taps["k_sub"].shape == (1, 32 * 64, 64)          # stride-3 keys of a 96x192 map


def test_oracle_native_size_checksums():
    """769x1537 (the only size the unpatched reference accepts, LayerNorm([97,193]))."""
    name = "td4_r18_769x1537_chk"
    arch, backbone = GOLDEN_CASES[name]
    g, m = load_golden(name)
    assert (m["h8"], m["w8"]) == (97, 193)
    oracle, _ = make_oracle(arch, backbone, m["H"], m["W"])
    frames = synth_clip(m["n_frames"], m["H"], m["W"], batch=m["batch"], clip_id=0)
    for i, f in enumerate(frames):
        out = oracle(f, pos_id=i % 4)
        head = oracle.taps["head"]
        assert max_abs(head[:, :, ::8, ::16], g[f"head_sub_{i}"]) <= TOL
        assert max_abs(out[:, :, ::64, ::128], g[f"logits_sub_{i}"]) <= TOL
        assert abs(head.double().mean().item() - float(g[f"head_mean_{i}"])) <= 1e-6
    assert oracle.Q_queue[0].shape == (1, 25 * 49, 64)  # P' = 1225 keys


def test_warmup_frames_skip_attention():
    """While the FIFO is not full the output is head(LN(v_cur)) (td4_psp18.py:142-143)."""
    oracle, _ = make_oracle("td4_psp18", "resnet18", 64, 64)
    frames = synth_clip(4, 64, 64)
    for i in range(3):
        oracle(frames[i], pos_id=i)
        assert "v_prop" not in oracle.taps
    oracle(frames[3], pos_id=3)
    assert "v_prop" in oracle.taps


def test_dilation_plan_matches_reference_table():
    """layer3 block0 conv1 d1 / conv2 d2, rest d2; layer4 conv1 d=[4,8,16][i], conv2 d4."""
    kind, plan = stage_plan("resnet18")
    assert kind == "basic"
    assert [(b["d1"], b["d2"]) for b in plan[2]] == [(1, 2), (2, 2)]
    assert [(b["d1"], b["d2"]) for b in plan[3]] == [(4, 4), (8, 4)]
    assert plan[1][0]["stride"] == 2 and plan[1][0]["downsample"]
    kind, plan = stage_plan("resnet50")
    assert kind == "bottleneck" and plan[0][0]["downsample"] and not plan[0][1]["downsample"]
    assert [b["d1"] for b in plan[3]] == [4, 8, 16]


def test_state_dict_template_sizes():
    """728 tensors / 54.9 M elements for td4-psp18, 776 / 65.5 M for td2-psp50."""
    sd = state_dict_template("td4_psp18")
    assert len(sd) == 728 and abs(sum(v.numel() for v in sd.values()) / 1e6 - 54.9) < 0.1
    sd = state_dict_template("td2_psp50")
    assert len(sd) == 776 and abs(sum(v.numel() for v in sd.values()) / 1e6 - 65.5) < 0.1
    sd = state_dict_template("pspnet")   # PSPNet-101: 670 tensors / 67.9 M (checked against the reference in make_golden.py)
    assert len(sd) == 670 and abs(sum(v.numel() for v in sd.values()) / 1e6 - 67.9) < 0.1
    from oracle.td2fa_oracle import td2fa_state_dict_template
    sd = td2fa_state_dict_template("resnet18")   # checked against the reference module in make_golden_fanet.py
    assert sd["ffm_32_1.up.conv.weight"].shape == (256, 512, 1, 1) and sd["head_aux2.conv_out.weight"].shape == (19, 64, 1, 1)


def test_oracle_ref_is_a_byte_identical_copy_of_the_reference_tree():
    """oracle/_ref/Testing (oracle/make_ref.py) must be the reference's Testing/ tree byte for byte: the manifest
    written at copy time is re-derived from the copy, and from /root/reference when that exists here."""
    import json
    import os
    from oracle import make_ref
    man = os.path.join(os.path.dirname(make_ref.DST), "MANIFEST.json")
    if not os.path.isfile(man):
        pytest.skip("oracle/_ref absent")
    files = json.load(open(man))["files"]
    assert "test.py" in files and "dataloader.py" in files and "model/pspnet/td4_psp18.py" in files
    assert make_ref.manifest(make_ref.DST) == files
    if os.path.isdir(make_ref.SRC):
        assert make_ref.manifest(make_ref.SRC) == files
