"""GPU: TD2-FANet (tdnet_b200.model.td2_fa) -- the new kernels against torch, the model against
the fixtures the reference produced (tests/golden/td2fa_*.npz) and against the CPU oracle at larger sizes.

Tolerances as in test_model_gpu.py: logits max-abs error <= LOGIT_TOL (logits std ~1.6, |max| ~9); arg-max labels
identical on every pixel whose reference top-1/top-2 margin exceeds twice the measured error.
"""
import ctypes as C

import pytest
import torch

from common import (CH_STRIDE, FANET_GOLDEN_CASES, argmax_report, load_golden, make_fanet_oracle, max_abs, record,
                    rel_l2)
from tdnet_b200.synth import synth_clip

pytestmark = pytest.mark.gpu

LOGIT_TOL = 2e-4
TAP_TOL = 1e-3


def build_fanet(backbone, h4, w4, sd, mode="tc"):
    from tdnet_b200.model import td2_fa
    net = td2_fa.td2_fa(nclass=19, backbone=backbone, path_num=2, ln_shape=(h4, w4))
    net.load_state_dict(sd, strict=True)
    net.engine_mode = mode
    return net.eval().to("cuda:0")


def tap(view):
    return view.torch().permute(0, 3, 1, 2).contiguous().cpu()


@pytest.fixture(scope="module")
def env():
    import __graft_entry__ as g
    g.build()
    from tdnet_b200 import _cabi
    from tdnet_b200.engine import View
    return _cabi.load(), _cabi, View, torch.device("cuda:0")


def _fill(view, x):
    """Write the dense fp32 tensor x [n,h,w,c] into an engine View of either dtype."""
    flat = x.reshape(-1)
    if view.split:
        hi = flat.half()
        view.base.copy_(hi)
        view.lo.copy_((flat - hi.float()).half())
        return view.torch()           # what the kernel actually reads (hi + lo)
    view.base.copy_(flat)
    return x


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("n,h,w,c", [(1, 16, 24, 64), (2, 13, 21, 128), (1, 2, 3, 512), (1, 37, 53, 2048),
                                     (1, 128, 256, 64)])
def test_fa_linear_attention_kernels_against_torch(env, n, h, w, c, split):
    """tdn_fa_context + tdn_fa_apply == F.normalize / matmul / matmul of FAModule.forward (td2_fa.py:358-367)."""
    lib, cabi, View, dev = env
    g = torch.Generator(device="cuda").manual_seed(n * 1000 + h * 10 + c)
    q = torch.randn(n, h, w, 32, generator=g, device="cuda") * 3 + 0.5
    k = torch.randn(n, h, w, 32, generator=g, device="cuda") * 2 - 0.3
    k[0, 0, 0] = 0                                              # a zero key row: normalize's eps clamp
    v = torch.randn(n, h, w, c, generator=g, device="cuda").abs()
    qv, kv = View.alloc(n, h, w, 32, dev), View.alloc(n, h, w, 32, dev)
    vv, yv = View.alloc(n, h, w, c, dev, split=split), View.alloc(n, h, w, c, dev, split=split)
    _fill(qv, q), _fill(kv, k)
    v_seen = _fill(vv, v)
    f = torch.full((n, 32, c), float("nan"), device="cuda")
    ws_bytes = int(lib.tdn_fa_context_workspace_bytes(n, h, w, c))
    ws = torch.empty(ws_bytes // 4 + 4, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    kt, vt, qt, yt = kv.ct(), vv.ct(), qv.ct(), yv.ct()
    cabi.check(lib.tdn_fa_context(C.byref(kt), C.byref(vt), f.data_ptr(), ws.data_ptr(), ws_bytes, None), "fa_context")
    out_scale = 1.0 if h * w < 4096 else 2.0 ** -6          # an exact power of two applied to the stored output
    cabi.check(lib.tdn_fa_apply(C.byref(qt), f.data_ptr(), C.byref(yt), out_scale, flag.data_ptr(), None), "fa_apply")
    torch.cuda.synchronize()
    kn = torch.nn.functional.normalize(k.double().reshape(n, h * w, 32), p=2, dim=2, eps=1e-12)
    qn = torch.nn.functional.normalize(q.double().reshape(n, h * w, 32), p=2, dim=2, eps=1e-12)
    f_ref = kn.transpose(1, 2) @ v_seen.double().reshape(n, h * w, c)
    y_ref = (qn @ f_ref).reshape(n, h, w, c)
    scale = float(f_ref.abs().max())
    assert max_abs(f.cpu(), f_ref.cpu()) <= 2e-6 * scale + 1e-6, max_abs(f.cpu(), f_ref.cpu())
    yscale = float(y_ref.abs().max())
    assert max_abs(yv.torch().cpu() / out_scale, y_ref.cpu()) <= 4e-6 * yscale + 1e-6
    assert int(flag.item()) == 0
    # bit-reproducible: a second run gives identical sums (fixed-order reduction, no atomics)
    f2 = torch.empty_like(f)
    cabi.check(lib.tdn_fa_context(C.byref(kt), C.byref(vt), f2.data_ptr(), ws.data_ptr(), ws_bytes, None), "fa_context")
    torch.cuda.synchronize()
    assert torch.equal(f, f2)


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("with_up", [True, False])
def test_add_upsampled_against_torch(env, split, with_up):
    """tdn_add_upsampled == F.interpolate(up, size, bilinear, align_corners=True) + (a + b) (td2_fa.py:373, 398-402)."""
    lib, cabi, View, dev = env
    n, h, w, c, hu, wu = 2, 13, 21, 64, 9, 13                 # 9x13 = a 7x11 map after the `up` conv (+2)
    g = torch.Generator(device="cuda").manual_seed(7)
    a, b = torch.randn(n, h, w, c, generator=g, device="cuda"), torch.randn(n, h, w, c, generator=g, device="cuda") * 5
    u = torch.randn(n, hu, wu, c, generator=g, device="cuda") * 3
    av, bv, uv = (View.alloc(*s, dev, split=split) for s in ((n, h, w, c), (n, h, w, c), (n, hu, wu, c)))
    ov = View.alloc(n, h, w, c, dev, split=split)
    a_s, b_s, u_s = _fill(av, a), _fill(bv, b), _fill(uv, u)
    at, bt, ut, ot = av.ct(), bv.ct(), uv.ct(), ov.ct()
    cabi.check(lib.tdn_add_upsampled(C.byref(at), C.byref(bt), C.byref(ut) if with_up else None, C.byref(ot), None),
               "add_upsampled")
    torch.cuda.synchronize()
    ref = a_s.double() + b_s.double()
    if with_up:
        ref = ref + torch.nn.functional.interpolate(u_s.double().permute(0, 3, 1, 2), (h, w), mode="bilinear",
                                                    align_corners=True).permute(0, 2, 3, 1)
    assert max_abs(ov.torch().cpu(), ref.cpu()) <= 4e-6 * float(ref.abs().max())


def test_fa_kernel_argument_errors(env):
    lib, cabi, View, dev = env
    k16, v = View.alloc(1, 4, 4, 16, dev), View.alloc(1, 4, 4, 64, dev)
    f = torch.empty(32 * 64, device="cuda")
    kt, vt = k16.ct(), v.ct()
    assert lib.tdn_fa_context(C.byref(kt), C.byref(vt), f.data_ptr(), f.data_ptr(), 1 << 20, None) == -2   # 32 key channels only
    k32 = View.alloc(1, 4, 4, 32, dev).ct()
    assert lib.tdn_fa_context(C.byref(k32), C.byref(vt), f.data_ptr(), f.data_ptr(), 16, None) == -5         # workspace
    assert lib.tdn_add_upsampled(C.byref(vt), C.byref(kt), None, C.byref(vt), None) == -1                     # dims


@pytest.mark.parametrize("mode", ["tc", "simt"])
@pytest.mark.parametrize("name", sorted(FANET_GOLDEN_CASES))
def test_fanet_matches_reference_golden(name, mode):
    backbone = FANET_GOLDEN_CASES[name]
    g, m = load_golden(name)
    _, sd = make_fanet_oracle(backbone, m["H"], m["W"])
    net = build_fanet(backbone, m["h8"], m["w8"], sd, mode)
    calls = m["n_frames"]
    frames = [f.cuda() for f in synth_clip(calls + 1, m["H"], m["W"], batch=m["batch"], clip_id=0)]
    for i in range(calls):
        out = net([frames[i], frames[i + 1]], pos_id=i % 2)
        torch.cuda.synchronize()
        assert out.shape == (m["batch"], 19, m["H"], m["W"]) and out.dtype == torch.float32 and out.is_cuda
        eng, plan = net._last
        err = max_abs(tap(plan.taps["head"]), g[f"head_{i}"])
        assert err <= LOGIT_TOL, (name, mode, i, err)
        if f"logits_{i}" in g:
            ref = torch.from_numpy(g[f"logits_{i}"])
            e = max_abs(out.cpu(), ref)
            rep = argmax_report(out.cpu(), ref, max(e, 1e-6))
            record(f"golden/{name}/{mode}/call{i}", max_abs=e, rel_l2=rel_l2(out.cpu(), ref), **rep)
            assert e <= LOGIT_TOL, (name, i, e)
            assert rep["mismatch_decided"] == 0, rep
    t, s = plan.taps, CH_STRIDE
    for key, stride in (("feat4", 1), ("feat32", s), ("up32", s), ("up16", s), ("sm16", 1), ("up8", 1), ("sm4", 1),
                        ("v", s), ("normed", s)):
        got = tap(t[key])[:, ::stride]
        assert tuple(got.shape) == g["tap_" + key].shape, key
        assert max_abs(got, g["tap_" + key]) <= TAP_TOL, (key, max_abs(got, g["tap_" + key]))
    assert max_abs(t["q"].torch().reshape(m["batch"], -1, 64).cpu(), g["tap_q"]) <= TAP_TOL
    assert max_abs(t["k_sub"].torch().reshape(m["batch"], -1, 64).cpu(), g["tap_k_sub"]) <= TAP_TOL
    assert max_abs(t["v_sub"].torch().reshape(m["batch"], -1, 256).cpu(), g["tap_v_sub"]) <= TAP_TOL
    assert max_abs(tap(t["fused"])[:, ::s], g["tap_atn"] + g["tap_v"]) <= TAP_TOL
    net.check_numeric_range()


@pytest.mark.parametrize("backbone,H,W,n", [("resnet18", 512, 1024, 1), ("resnet34", 360, 480, 2)])
def test_fanet_against_oracle(backbone, H, W, n):
    """Larger maps (softmax neither flat nor one-hot there, P' = 946 / 300 keys) against the oracle on the host CPU; the
    fourth call repeats the third through the captured CUDA graph and must reproduce it bit for bit."""
    from tdnet_b200.model.arch import feature_hw
    oracle, sd = make_fanet_oracle(backbone, H, W)
    h4, w4 = feature_hw(H, W)
    net = build_fanet(backbone, h4, w4, sd)
    frames = synth_clip(4, H, W, batch=n, clip_id=9)
    dev = [f.cuda() for f in frames]
    outs = []
    for i in range(3):
        ref = oracle([frames[i], frames[i + 1]], pos_id=i % 2)
        out = net([dev[i], dev[i + 1]], pos_id=i % 2).cpu()
        e = max_abs(out, ref)
        rep = argmax_report(out, ref, max(e, 1e-6))
        record(f"oracle/td2fa_{backbone}_{H}x{W}/call{i}", max_abs=e, rel_l2=rel_l2(out, ref), **rep)
        assert e <= LOGIT_TOL, (i, e)
        assert rep["mismatch_decided"] == 0, rep
        assert rep["near_ties"] <= 0.002 * rep["pixels"] + 2, rep
        outs.append(out)
    again = net([dev[2], dev[3]], pos_id=0).cpu()       # plan (1, steady) used for the 3rd time: graph replay
    assert torch.equal(again, outs[2])
    labels = net.forward_labels([dev[2], dev[3]], pos_id=0).cpu()
    assert labels.dtype == torch.uint8 and torch.equal(labels.long(), outs[2].argmax(1))
    net.check_numeric_range()


def test_fanet_native_size_768x1536_against_reference_checksums():
    """The reference's own size (LayerNorm([96, 192]) untouched, default ln_shape) against what the unmodified module
    produced there: sub-sampled head / logits and the head mean."""
    from common import load_golden
    g, m = load_golden("td2fa_r18_768x1536_chk")
    _, sd = make_fanet_oracle("resnet18", m["H"], m["W"])
    from tdnet_b200.model import td2_fa
    net = td2_fa.td2_fa(nclass=19, backbone="resnet18", path_num=2)            # default ln_shape = (96, 192)
    net.load_state_dict(sd, strict=True)
    net.eval().to("cuda:0")
    frames = [f.cuda() for f in synth_clip(m["n_frames"] + 1, m["H"], m["W"], clip_id=0)]
    for i in range(m["n_frames"]):
        out = net([frames[i], frames[i + 1]], pos_id=i % 2)
        e = max_abs(out[:, :, ::64, ::128].cpu(), g[f"logits_sub_{i}"])
        record(f"golden/td2fa_r18_768x1536_chk/call{i}", max_abs=e)
        assert e <= LOGIT_TOL, (i, e)
        head = tap(net._last[1].taps["head"])
        assert max_abs(head[:, :, ::8, ::16], g[f"head_sub_{i}"]) <= LOGIT_TOL
        assert abs(head.double().mean().item() - float(g[f"head_mean_{i}"])) <= 1e-5
    net.check_numeric_range()


def test_fanet_full_size_1024x2048_properties():
    """Cityscapes-sized frame pair: runs, finite, deterministic run to run, labels == arg-max of the logits."""
    from tdnet_b200.model.arch import feature_hw
    H, W = 1024, 2048
    _, sd = make_fanet_oracle("resnet18", H, W)
    net = build_fanet("resnet18", *feature_hw(H, W), sd)
    f = [x.cuda() for x in synth_clip(2, H, W, clip_id=2)]
    a = net(f, pos_id=0)
    b = net(f, pos_id=0)
    c = net(f, pos_id=0)
    torch.cuda.synchronize()
    assert a.shape == (1, 19, H, W) and torch.isfinite(a).all()
    assert torch.equal(a, b) and torch.equal(a, c)
    assert torch.equal(net.forward_labels(f, pos_id=0).long(), a.argmax(1))
    d = net(f, pos_id=1)
    assert torch.isfinite(d).all() and not torch.equal(a, d)
    net.check_numeric_range()


def test_fanet_api_errors():
    from tdnet_b200.model import td2_fa
    with pytest.raises(AssertionError):
        td2_fa.td2_fa(nclass=19, backbone="resnet18", path_num=4)
    with pytest.raises(AssertionError):
        td2_fa.td2_fa(nclass=19, backbone="resnet101", path_num=2)
    net = td2_fa.td2_fa(nclass=19, backbone="resnet18", path_num=2, ln_shape=(8, 12)).to("cuda:0")
    x = torch.zeros(1, 3, 64, 96, device="cuda")
    with pytest.raises(RuntimeError, match="inference path only"):
        net([x, x], pos_id=0)
    net.eval()
    with pytest.raises(RuntimeError, match="Only Two Paths"):
        net([x, x], pos_id=2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        net([x.cpu(), x.cpu()], pos_id=0)
    with pytest.raises(RuntimeError, match="normalized_shape"):
        net([torch.zeros(1, 3, 128, 96, device="cuda")] * 2, pos_id=0)
    assert net([x, x], pos_id=1).shape == (1, 19, 64, 96)
