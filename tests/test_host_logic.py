"""CPU: host-side logic of the engine -- BN folding, SPLIT16 weight packing, NHWC view arithmetic, and the
architecture tables -- checked against torch on CPU tensors (no kernel is launched)."""
import torch
import torch.nn.functional as F

from tdnet_b200.engine import PackedConv, View, split_rows_pow2
from tdnet_b200.model import arch as A
from tdnet_b200.synth import synth_state_dict


def _weights(arch="td2_psp50", backbone="resnet18", hw=(4, 4)):
    m = A.build_arch(arch, backbone, 19)
    tmpl = {k: torch.zeros(shape, dtype=torch.long if kind == "long_buffer" else torch.float32)
            for k, (shape, kind) in A.parameter_table(m, hw).items()}
    return m, synth_state_dict(tmpl, seed=1)


def test_bn_folding_equals_conv_bn_act():
    """PackedConv: BN(conv(x) + b) == conv(x) * scale + bias (td4_psp18.py:11-24, transformer.py:142-161)."""
    m, sd = _weights()
    spec = A.encoding_convs(m, 1)["w_qs"][0]            # conv1x1 + bias -> BN -> LeakyReLU
    pc = PackedConv(spec, sd, torch.device("cpu"))
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, spec.cin, 5, 7, generator=g)
    ref = F.conv2d(x, sd[spec.name + ".weight"], sd[spec.name + ".bias"])
    ref = F.leaky_relu(F.batch_norm(ref, sd[spec.bn + ".running_mean"], sd[spec.bn + ".running_var"],
                                    sd[spec.bn + ".weight"], sd[spec.bn + ".bias"], False, 0.0, 1e-5), 0.01)
    w = pc.weight.permute(0, 3, 1, 2)                   # packed [cout,kh,kw,cin] back to OIHW
    got = F.leaky_relu(F.conv2d(x, w) * pc.scale.view(1, -1, 1, 1) + pc.bias.view(1, -1, 1, 1), 0.01)
    assert (got - ref).abs().max() < 2e-5
    # the 3-channel stem is padded to 4 input channels with zeros
    stem = PackedConv(m.stems[1][0], sd, torch.device("cpu"))
    assert stem.weight.shape == (64, 7, 7, 4) and float(stem.weight[..., 3].abs().max()) == 0.0
    # PSP branch: only this path's output-channel slice is packed
    psp = PackedConv(A.psp_convs(m, 2)[0], sd, torch.device("cpu"), row_slice=(64, 128))
    assert psp.weight.shape[0] == 64 and torch.equal(psp.weight[:, 0, 0], sd["psp2.conv1.0.weight"][64:128, :, 0, 0])


def test_split16_weight_packing_is_fp32_faithful():
    g = torch.Generator().manual_seed(1)
    w = torch.randn(37, 576, generator=g) * torch.logspace(-6, 1, 37).view(-1, 1)   # rows of very different scale
    hi, lo, inv = split_rows_pow2(w)
    assert hi.dtype == lo.dtype == torch.float16
    rec = (hi.double() + lo.double()) * inv.double().view(-1, 1)
    rel = ((rec - w.double()).abs() / w.double().abs().amax(dim=1, keepdim=True)).max()
    assert rel < 2.0 ** -22, float(rel)
    assert torch.equal(torch.log2(inv), torch.log2(inv).round())       # exact powers of two
    assert torch.isfinite(hi.float()).all() and hi.float().abs().max() < 32768


def test_nhwc_views():
    base = torch.arange(2 * 6 * 10 * 8, dtype=torch.float32)
    v = View(base, 2, 6, 10, 8)
    t = v.torch()
    assert torch.equal(v.channels(2, 6).torch(), t[..., 2:6])
    assert torch.equal(v.subsample(4).torch(), t[:, ::4, ::4])          # MaxPool2d(kernel 1, stride 4)
    assert v.subsample(4).h == 2 and v.subsample(4).w == 3
    tok = v.tokens()
    assert (tok.n, tok.h, tok.w, tok.c, tok.sn) == (1, 1, 60, 8, 480)
    assert torch.equal(v.image(1).torch()[0], t[1])
    pooled = View(torch.arange(2 * 50 * 4, dtype=torch.float32), 2, 1, 50, 4)
    assert torch.equal(pooled.rows(5, 14, 3, 3).torch(), pooled.torch()[:, 0, 5:14].reshape(2, 3, 3, 4))
    c = v.ct()
    assert (c.n, c.h, c.w, c.c, c.stride_n, c.stride_h, c.stride_w, c.dtype) == (2, 6, 10, 8, 480, 80, 8, 0)
    s = View.alloc(1, 2, 3, 8, torch.device("cpu"), split=True)
    assert s.split and s.ct().dtype == 1 and s.ct().data_lo is not None


def test_arch_tables_match_reference_structure():
    m = A.build_arch("td4_psp18", "resnet18", 19)
    assert (m.paths, m.depth, m.c4, m.d_v, m.head_mid) == (4, 3, 512, 512, 128)
    assert m.hop_modules(1) == ["atn1_2", "atn1_3", "atn1_4"] and m.hop_modules(2) == ["atn2_3", "atn2_4", "atn2_1"]
    assert m.hop_modules(4) == ["atn4_1", "atn4_2", "atn4_3"]          # td4_psp18.py:145-147,...,204-206
    assert [m.psp_pid(p) for p in (1, 2, 3, 4)] == [0, 1, 0, 1]
    l4 = m.stages[1][-2:]                                               # layer4 multi-grid (resnet.py:181-197)
    assert [(b.convs[0].dilation, b.convs[1].dilation) for b in l4] == [(4, 4), (8, 4)]
    assert l4[0].downsample is not None and l4[0].downsample.stride == 1
    m2 = A.build_arch("td2_psp50", "resnet50", 19)
    assert (m2.paths, m2.depth, m2.c4, m2.d_v, m2.head_mid) == (2, 1, 2048, 512, 256)
    assert m2.hop_modules(2) == ["atn2"] and len(m2.stems[1]) == 3
    assert A.feature_hw(769, 1537) == (97, 193) and A.feature_hw(720, 960) == (90, 120)
    with_bn = [c for c in A.head_convs(m, 1)]
    assert with_bn[0].bn == "head1.conv5.1" and with_bn[1].bias and with_bn[1].cout == 19
