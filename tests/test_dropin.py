"""CPU: the drop-in `model` package exposes the reference's import surface (Testing/test.py:9)."""
import importlib
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dropin_import_surface():
    sys.path.insert(0, os.path.join(ROOT, "tdnet_b200", "dropin"))
    try:
        sys.modules.pop("model", None)
        model = importlib.import_module("model")
        from model import pspnet, td2_psp50, td4_psp18  # noqa: F401  (the line in Testing/test.py:9)
        net = model.td4_psp18.td4_psp18(nclass=19, path_num=4, model_path="/nonexistent.pth")  # test.py:26
        assert isinstance(net, torch.nn.Module) and net.path_num == 4
        net.eval()
        assert len(net.state_dict()) == 728
        net2 = model.td2_psp50.td2_psp50(nclass=19, path_num=2)
        assert len(net2.state_dict()) == 776 and net2.Q_queue == []
        net3 = model.pspnet.pspnet(nclass=19)                      # PSPNet-101 comparison model (test.py:29-31)
        assert len(net3.state_dict()) == 670 and net3.backbone == "resnet101"
        with pytest.raises(RuntimeError, match="unknown backbone"):
            model.pspnet.pspnet(nclass=19, backbone="resnet152")
    finally:
        sys.path.pop(0)
        sys.modules.pop("model", None)


def test_checkpoint_roundtrip_strict(tmp_path):
    """pretrained_mp_load: torch.load + load_state_dict(strict=True) (td4_psp18.py:232-240)."""
    from tdnet_b200.model import td2_psp50
    from tdnet_b200.synth import synth_state_dict
    a = td2_psp50.td2_psp50(nclass=19, path_num=2, backbone="resnet18")
    sd = synth_state_dict(a.state_dict(), seed=3)
    path = tmp_path / "ckpt.pth"
    torch.save(sd, path)
    b = td2_psp50.td2_psp50(nclass=19, path_num=2, backbone="resnet18", model_path=str(path))
    for k, v in b.state_dict().items():
        assert torch.equal(v, sd[k]), k
    with pytest.raises(RuntimeError, match="no CPU path"):
        b(torch.zeros(1, 3, 32, 32), pos_id=0)


def test_engine_plans_build_without_gpu():
    """Host logic only: plan construction (buffer pool, descriptors, op order) for both engine modes."""
    from collections import Counter
    from tdnet_b200.engine import Engine
    from tdnet_b200.model import arch as A
    m = A.build_arch("td4_psp18", "resnet18", 19)
    h8, w8 = A.feature_hw(97, 161)
    sd = {k: torch.full(shape, 0.01) if kind != "long_buffer" else torch.zeros(shape, dtype=torch.long)
          for k, (shape, kind) in A.parameter_table(m, (h8, w8)).items()}
    for k in sd:
        if k.endswith("running_var"):
            sd[k].fill_(1.0)
    # 28 = every conv but the 3-ch stem and the PSP branch convs (with the pyramid fold the first key projection is a
    # tensor-core GEMM over [c4 slice | interpolation channels] even on the 4 x 6 key grid; the 19-class classifier runs
    # as a GEMM with 24 output channels)
    for mode, tc_ops in (("simt", 0), ("tc", 28)):
        eng = Engine(m, sd, 1, 97, 161, torch.device("cpu"), (h8, w8), mode=mode)
        warm, steady = eng.plan(1, False), eng.plan(1, True)
        c = Counter(fn.__name__ for fn, _ in steady.ops)
        assert c["tdn_conv2d_tc"] == tc_ops and c["tdn_upsample_logits"] == 1
        assert (c["tdn_attention_tc"] == 3) == (mode == "tc")
        assert len(steady.ops) > len(warm.ops)           # warm-up frames skip the attention hops
        assert eng.pk == 4 * 6 and len(eng.k_slots) == 3  # P' = ceil(13/4) x ceil(21/4), FIFO depth 3
    with pytest.raises(RuntimeError, match="normalized_shape"):
        Engine(m, sd, 1, 64, 64, torch.device("cpu"), (h8, w8))


def test_td2_fa_constructor_state_dict_and_pretrained_init(tmp_path):
    """td2_fa (Training/ptsemseg/models/td2_fanet/td2_fa.py:16-85): constructor contract, state-dict layout, and
    pretrained_init() (:246-274) which fills BOTH sub-networks from one single-path FANet checkpoint
    (ptsemseg/utils.py:35-66: resnet. / ffm_*. / clslayer_8. / clslayer_32. key groups)."""
    from tdnet_b200.model import td2_fa
    from tdnet_b200.synth import synth_state_dict
    net = td2_fa.td2_fa(nclass=19, backbone="resnet18", path_num=2)
    sd = net.state_dict()
    assert len(sd) == 616 and sd["layer_norm1.ln.weight"].shape == (96, 192)         # LayerNorm([96, 192]), :71-72
    assert sd["ffm_32_1.up.conv.weight"].shape == (256, 512, 1, 1) and "pretrained1.fc.weight" not in sd
    with pytest.raises(AssertionError):
        td2_fa.td2_fa(nclass=19, backbone="resnet18", path_num=4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        net.eval()([torch.zeros(1, 3, 64, 64)] * 2, pos_id=0)
    # a single-path FANet checkpoint: sub-network 1 of a synthetic model, renamed to the single-path key groups
    rename = {"pretrained1": "resnet", "ffm_32_1": "ffm_32", "ffm_16_1": "ffm_16", "ffm_8_1": "ffm_8",
              "ffm_4_1": "ffm_4", "head1": "clslayer_8", "head_aux1": "clslayer_32"}
    full = synth_state_dict(sd, seed=5)
    single = {rename[k.split(".")[0]] + "." + k.split(".", 1)[1]: v for k, v in full.items() if k.split(".")[0] in rename}
    path = tmp_path / "fanet.pth"
    torch.save(single, path)
    b = td2_fa.td2_fa(nclass=19, backbone="resnet18", path_num=2, mdl_path=str(path))
    got = b.state_dict()
    for k, v in full.items():
        head = k.split(".")[0]
        if head in rename:
            twin = k.replace(head, head[:-1] + "2", 1)
            assert torch.equal(got[k], v) and torch.equal(got[twin], v), k
    assert not torch.equal(got["enc1.w_vs.0.conv.weight"], full["enc1.w_vs.0.conv.weight"])   # not in the checkpoint
    del single["resnet.conv1.weight"]
    torch.save(single, path)
    with pytest.raises(RuntimeError, match="missing"):
        td2_fa.td2_fa(nclass=19, backbone="resnet18", path_num=2, mdl_path=str(path))


def test_forward_path_methods_mirror_the_reference_surface():
    """forward_path1..4 (td4_psp18.py:137-212) / forward_path1..2 (td2_psp50.py:112-143, td2_fa.py:87-186) exist with the
    reference's arity; like everything else they refuse CPU tensors instead of falling back."""
    from tdnet_b200.model import td2_fa, td2_psp50, td4_psp18
    x = torch.zeros(1, 3, 32, 32)
    td4 = td4_psp18.td4_psp18(nclass=19, path_num=4).eval()
    for k in range(1, 5):
        with pytest.raises(RuntimeError, match="no CPU path"):
            getattr(td4, f"forward_path{k}")(x)
    td2 = td2_psp50.td2_psp50(nclass=19, path_num=2, backbone="resnet18").eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        td2.forward_path2(x)
    with pytest.raises(AttributeError):
        td2.forward_path3(x)
    fa = td2_fa.td2_fa(nclass=19, backbone="resnet18", path_num=2).eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        fa.forward_path1([x, x])
    with pytest.raises(AttributeError):
        fa.forward_path3([x, x])


def test_foreign_norm_layer_is_refused_not_ignored():
    """BatchNorm2d (eval) is folded into the convolution epilogues; a user-supplied normalisation layer of another kind
    would silently compute something else, so the constructor raises (the reference's own default and torch's
    BatchNorm2d are accepted; td4_psp18.py:11-24,52-53)."""
    import torch.nn as nn
    from tdnet_b200.model import td2_psp50, td4_psp18
    td4_psp18.td4_psp18(nclass=19, path_num=4)                                   # the reference's default
    td4_psp18.td4_psp18(nclass=19, path_num=4, norm_layer=nn.BatchNorm2d)
    td4_psp18.td4_psp18(nclass=19, path_num=4, norm_layer=None)
    for cls, kw in ((td4_psp18.td4_psp18, dict(path_num=4)), (td2_psp50.td2_psp50, dict(path_num=2))):
        with pytest.raises(RuntimeError, match="norm_layer"):
            cls(nclass=19, norm_layer=nn.GroupNorm, **kw)
        with pytest.raises(RuntimeError, match="norm_layer"):
            cls(nclass=19, norm_layer=nn.InstanceNorm2d, **kw)
