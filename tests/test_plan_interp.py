"""CPU: the host side of the frame engine (plan construction: op order, views, weight packing, folded BatchNorm,
algebraic rewrites) executed by the CPU plan interpreter (tests/plan_interp.py) and compared with what the unmodified
reference produced (tests/golden/*.npz).  No CUDA kernel runs here; kernels are checked by the -m gpu tests."""
import numpy as np
import pytest
import torch

from common import CH_STRIDE, FANET_GOLDEN_CASES, GOLDEN_CASES, PSPNET_GOLDEN_CASES, load_golden, max_abs
from plan_interp import read, run_plan
from tdnet_b200.engine import Engine
from tdnet_b200.model import arch as A
from tdnet_b200.synth import synth_clip, synth_state_dict


def _nchw(view):
    return read(view.ct()).permute(0, 3, 1, 2)


@pytest.mark.parametrize("mode", ["tc", "simt"])
@pytest.mark.parametrize("name", sorted(FANET_GOLDEN_CASES))
def test_fanet_plan_matches_reference_golden(name, mode):
    g, meta = load_golden(name)
    H, W, n, calls = meta["H"], meta["W"], meta["batch"], meta["n_frames"]
    m = A.build_arch("td2_fa", FANET_GOLDEN_CASES[name], 19)
    h4, w4 = A.feature_hw(H, W)
    tmpl = {k: torch.zeros(shape, dtype=torch.long if kind == "long_buffer" else torch.float32)
            for k, (shape, kind) in A.parameter_table(m, (h4, w4)).items()}
    sd = synth_state_dict(tmpl, seed=0)
    eng = Engine(m, sd, n, H, W, torch.device("cpu"), (h4, w4), mode=mode)
    frames = synth_clip(calls + 1, H, W, batch=n, clip_id=0)
    tol = 2e-4
    for i in range(calls):
        plan = eng.plan(i % 2 + 1, True)
        prev, cur = frames[i].contiguous(), frames[i + 1].contiguous()
        out = torch.empty(n, 19, H, W)
        run_plan(plan, {"img": cur.data_ptr(), "img2": prev.data_ptr(), "out": out.data_ptr()})
        head = _nchw(plan.taps["head"])
        assert max_abs(head, g[f"head_{i}"]) <= tol, (name, mode, i, max_abs(head, g[f"head_{i}"]))
        if f"logits_{i}" in g:
            assert max_abs(out, g[f"logits_{i}"]) <= tol
    t, s = plan.taps, CH_STRIDE
    for key, tap, stride in (("feat4", "feat4", 1), ("feat32", "feat32", s), ("up32", "up32", s), ("up16", "up16", s),
                             ("sm16", "sm16", 1), ("up8", "up8", 1), ("sm4", "sm4", 1), ("v", "v", s),
                             ("normed", "normed", s)):
        got = _nchw(t[tap])[:, ::stride]
        ref = g["tap_" + key]
        assert tuple(got.shape) == ref.shape, key
        assert max_abs(got, ref) <= tol * max(1.0, float(np.abs(ref).max())), (key, max_abs(got, ref))
    q = read(t["q"].ct()).reshape(n, -1, 64)
    assert max_abs(q, g["tap_q"]) <= tol
    assert max_abs(read(t["k_sub"].ct()).reshape(n, -1, 64), g["tap_k_sub"]) <= tol
    assert max_abs(read(t["v_sub"].ct()).reshape(n, -1, 256), g["tap_v_sub"]) <= tol
    fused = _nchw(t["fused"])[:, ::s]
    ref_fused = g["tap_atn"] + g["tap_v"]
    assert max_abs(fused, ref_fused) <= tol * max(1.0, float(np.abs(ref_fused).max()))
    assert int(eng.range_flag.item()) == 0


def _synth_weights(m, ln_shape):
    tmpl = {k: torch.zeros(shape, dtype=torch.long if kind == "long_buffer" else torch.float32)
            for k, (shape, kind) in A.parameter_table(m, ln_shape).items()}
    return synth_state_dict(tmpl, seed=0)


@pytest.mark.parametrize("mode", ["tc", "tc_nofold", "simt"])
@pytest.mark.parametrize("name", ["td4_r18_97x161", "td4_r18_128x256", "td2_r50_64x128", "td2_r34_80x112", "td4_r50_64x64_n2"])
def test_td_plans_match_reference_golden(name, mode, monkeypatch):
    """Warm-up and steady plans of the TD models frame by frame, FIFO included (push order, shifts, slot views).  'tc' runs
    the pyramid fold (Encoding convs over [c4 slice | interpolation channels], no z); 'tc_nofold' the materialised z."""
    if mode == "tc_nofold":
        monkeypatch.setenv("TDNET_B200_PSP_FOLD", "0")
        mode = "tc"
    arch, backbone = GOLDEN_CASES[name]
    g, meta = load_golden(name)
    H, W, n = meta["H"], meta["W"], meta["batch"]
    m = A.build_arch(arch, backbone, 19)
    ln = (meta["h8"], meta["w8"])
    eng = Engine(m, _synth_weights(m, ln), n, H, W, torch.device("cpu"), ln, mode=mode)
    frames = synth_clip(meta["n_frames"], H, W, batch=n, clip_id=0)
    for i, f in enumerate(frames):
        plan = eng.plan(i % m.paths + 1, i >= m.depth)
        img, out = f.contiguous(), torch.empty(n, 19, H, W)
        run_plan(plan, {"img": img.data_ptr(), "out": out.data_ptr()})
        err = max_abs(_nchw(plan.taps["head"]), g[f"head_{i}"])
        assert err <= 2e-4, (name, mode, i, err)
        if f"logits_{i}" in g:
            assert max_abs(out, g[f"logits_{i}"]) <= 2e-4
    s = CH_STRIDE
    assert ("z" in plan.taps) == (not eng.psp_fold)
    if "z" in plan.taps:
        assert max_abs(_nchw(plan.taps["z"])[:, ::s], g["tap_z"]) <= 1e-3
    assert max_abs(_nchw(plan.taps["normed"])[:, ::s], g["tap_normed"]) <= 1e-3
    # the newest FIFO entry is what the reference queued (Encoding(pre=True), transformer.py:34-50)
    assert max_abs(read(eng.k_slots[-1].ct()).reshape(n, -1, 64), g["tap_k_sub"]) <= 1e-3
    assert max_abs(read(eng.v_slots[-1].ct()).reshape(n, -1, m.d_v), g["tap_v_sub"]) <= 1e-3
    assert max_abs(read(eng.q_slots[-1].ct()).reshape(n, -1, 64), g["tap_q_sub"]) <= 1e-3
    assert int(eng.range_flag.item()) == 0


@pytest.mark.parametrize("name", sorted(PSPNET_GOLDEN_CASES))
def test_pspnet_plan_matches_reference_golden(name):
    g, meta = load_golden(name)
    H, W, n = meta["H"], meta["W"], meta["batch"]
    m = A.build_arch("pspnet", PSPNET_GOLDEN_CASES[name], 19)
    eng = Engine(m, _synth_weights(m, (0, 0)), 1, H, W, torch.device("cpu"), (0, 0), mode="tc")
    plan = eng.plan(1, True)
    for i, f in enumerate(synth_clip(meta["n_frames"], H, W, batch=n, clip_id=0)):
        img, out = f[-1:].contiguous(), torch.empty(1, 19, H, W)      # pspnet.py:74 `x = x[-1:]`
        run_plan(plan, {"img": img.data_ptr(), "out": out.data_ptr()})
        scale = max(1.0, float(np.abs(g[f"head_{i}"]).max()))
        assert max_abs(_nchw(plan.taps["head"]), g[f"head_{i}"]) <= 2e-4 * scale
        if f"logits_{i}" in g:
            assert max_abs(out, g[f"logits_{i}"]) <= 2e-4 * scale


def test_label_and_preview_output_stages():
    """forward_labels / forward_preview variants of the last op: labels == arg-max of the logits; the preview equals the
    cv2.INTER_NEAREST resize of that label map (Testing/test.py:61-64) -- host-built coordinate tables included."""
    from oracle import cv2_resize_oracle as O
    name = "td4_r18_97x161"
    arch, backbone = GOLDEN_CASES[name]
    g, meta = load_golden(name)
    H, W = meta["H"], meta["W"]
    m = A.build_arch(arch, backbone, 19)
    ln = (meta["h8"], meta["w8"])
    eng = Engine(m, _synth_weights(m, ln), 1, H, W, torch.device("cpu"), ln, mode="tc")
    plan = eng.plan(1, False)
    img = synth_clip(1, H, W)[0].contiguous()
    logits, labels = torch.empty(1, 19, H, W), torch.empty(1, H, W, dtype=torch.uint8)
    run_plan(plan, {"img": img.data_ptr(), "out": logits.data_ptr()})
    run_plan(plan, {"img": img.data_ptr(), "out": labels.data_ptr()}, last_op=plan.labels_op)
    assert torch.equal(labels.long(), logits.argmax(1))
    for ph, pw in ((H // 4, W // 4), (37, 53)):
        prev = torch.empty(1, ph, pw, dtype=torch.uint8)
        run_plan(plan, {"img": img.data_ptr(), "out": prev.data_ptr()}, last_op=eng.preview_op(plan, ph, pw))
        want = O.resize_nearest(labels[0].numpy().astype(np.int8), pw, ph)
        assert np.array_equal(prev[0].numpy().astype(np.int8), want)


@pytest.mark.parametrize("arch,backbone,H,W", [
    ("td4_psp18", "resnet18", 65, 97), ("td4_psp18", "resnet18", 120, 200), ("td2_psp50", "resnet34", 72, 136),
    ("td2_fa", "resnet18", 96, 160), ("td2_fa", "resnet18", 130, 210), ("pspnet", "resnet18", 70, 110)])
def test_plans_at_sizes_without_fixtures_match_the_oracle(arch, backbone, H, W):
    """Odd input sizes that no fixture covers (ragged maps, key grids that do not divide, `up` frames of FANet): the
    interpreted plan against the CPU oracle on the same synthetic weights, warm-up into steady state."""
    from common import make_fanet_oracle, make_oracle, make_pspnet_oracle
    m = A.build_arch(arch, backbone, 19)
    ln = A.feature_hw(H, W) if arch != "pspnet" else (0, 0)
    if arch == "td2_fa":
        oracle, sd = make_fanet_oracle(backbone, H, W)
    elif arch == "pspnet":
        oracle, sd = make_pspnet_oracle(backbone)
    else:
        oracle, sd = make_oracle(arch, backbone, H, W)
    eng = Engine(m, sd, 1, H, W, torch.device("cpu"), ln, mode="tc")
    frames = synth_clip(m.depth + 3, H, W, clip_id=6)
    for i in range(m.depth + 2):
        cur = frames[i + 1].contiguous()
        out = torch.empty(1, 19, H, W)
        if arch == "td2_fa":
            prev = frames[i].contiguous()
            ref = oracle([prev, cur], pos_id=i % 2)
            run_plan(eng.plan(i % 2 + 1, True), {"img": cur.data_ptr(), "img2": prev.data_ptr(), "out": out.data_ptr()})
        else:
            ref = oracle(cur, pos_id=i % m.paths)
            run_plan(eng.plan(i % m.paths + 1, i >= m.depth), {"img": cur.data_ptr(), "out": out.data_ptr()})
        scale = max(1.0, float(ref.abs().max()))
        assert max_abs(out, ref) <= 2e-4 * scale, (arch, H, W, i, max_abs(out, ref))
    assert int(eng.range_flag.item()) == 0
