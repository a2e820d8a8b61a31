"""Architecture tables of the two temporally-distributed PSP models, of the single-path PSPNet
comparison model and of TD2-FANet (product side).

A declarative description of what the reference builds imperatively in
Testing/model/pspnet/td4_psp18.py:29-121, td2_psp50.py:29-96, pspnet.py:31-157 and resnet.py:114-202: which
convolutions exist, their geometry, and which state-dict entries hold their parameters.  The
engine (tdnet_b200/engine.py) turns these tables into C-ABI calls; `parameter_table` gives the flat
state-dict (name -> shape) that checkpoints of the reference are loaded against with strict=True.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

BLOCKS = {"resnet18": ("basic", (2, 2, 2, 2)), "resnet34": ("basic", (3, 4, 6, 3)),
          "resnet50": ("bottleneck", (3, 4, 6, 3)), "resnet101": ("bottleneck", (3, 4, 23, 3))}
STAGE_PLANES = (64, 128, 256, 512)
MULTI_GRID = (4, 8, 16)  # resnet.py:181 `multi_dilations`


@dataclass(frozen=True)
class Conv:
    """One convolution with optional BatchNorm (folded at prepare time) and activation."""
    name: str              # state-dict prefix of the conv ('....conv1')
    cin: int
    cout: int
    k: int = 1
    stride: int = 1
    dilation: int = 1
    bn: Optional[str] = None   # state-dict prefix of the BatchNorm that follows, if any
    bias: bool = False
    act: str = "none"          # 'none' | 'relu' | 'leaky_relu'

    @property
    def pad(self):
        return self.dilation * (self.k - 1) // 2


@dataclass(frozen=True)
class Block:
    """Residual block: `convs` in order, the last one takes the residual (identity or `downsample`)
    and the closing ReLU (resnet.py:43-59, 91-111)."""
    convs: Tuple[Conv, ...]
    downsample: Optional[Conv]


@dataclass
class ModelArch:
    arch: str
    backbone: str
    paths: int
    depth: int        # FIFO depth: 3 (td4_psp18.py:131) / 1 (td2_psp50.py:106)
    c4: int           # backbone output channels
    d_k: int
    d_v: int
    head_mid: int
    nclass: int
    stems: Dict[int, List[Conv]] = field(default_factory=dict)
    stages: Dict[int, List[Block]] = field(default_factory=dict)
    stage_ends: Tuple[int, ...] = ()   # td2_fa: index of the last block of layer1..4 (the FPN taps feat4..feat32)
    c_exp: int = 1                     # td2_fa: block expansion (4 for the Bottleneck backbone)

    def prefix(self, path: int) -> str:
        """State-dict prefix of the path's backbone: `pretrained{path}` for the TD models, plain `pretrained`
        for pspnet (pspnet.py:49-66)."""
        return "pretrained" if self.arch == "pspnet" else f"pretrained{path}"

    def hop_modules(self, path: int) -> List[str]:
        """Attention modules a path walks through, oldest key frame first
        (td4_psp18.py:145-147,166-168,185-187,204-206; td2_psp50.py:120,137)."""
        if self.paths == 2:
            return [f"atn{path}"]
        return [f"atn{path}_{(path + j) % 4 + 1}" for j in range(3)]

    def psp_pid(self, path: int) -> int:
        return (path - 1) % 2  # td4_psp18.py:80-83, td2_psp50.py:76-77


def _backbone(prefix: str, backbone: str):
    kind, counts = BLOCKS[backbone]
    exp = 4 if kind == "bottleneck" else 1
    if kind == "bottleneck":  # deep_base stem, resnet.py:122-131 (+ bn1/relu :135-136)
        stem = [Conv(f"{prefix}.conv1.0", 3, 64, 3, stride=2, bn=f"{prefix}.conv1.1", act="relu"),
                Conv(f"{prefix}.conv1.3", 64, 64, 3, bn=f"{prefix}.conv1.4", act="relu"),
                Conv(f"{prefix}.conv1.6", 64, 128, 3, bn=f"{prefix}.bn1", act="relu")]
        inplanes = 128
    else:  # resnet.py:133-136
        stem = [Conv(f"{prefix}.conv1", 3, 64, 7, stride=2, bn=f"{prefix}.bn1", act="relu")]
        inplanes = 64
    blocks: List[Block] = []
    for si, (planes, count) in enumerate(zip(STAGE_PLANES, counts)):
        stride = 2 if si == 1 else 1                      # layer3/4 keep stride 1 when dilated (:139-147)
        layer_dil = (1, 1, 2, 4)[si]
        for bi in range(count):
            p = f"{prefix}.layer{si + 1}.{bi}"
            if bi == 0:
                first = MULTI_GRID[0] if si == 3 else (1 if layer_dil in (1, 2) else 2)
                s = stride
                ds = None
                if s != 1 or inplanes != planes * exp:
                    ds = Conv(f"{p}.downsample.0", inplanes, planes * exp, 1, stride=s, bn=f"{p}.downsample.1")
            else:
                first = MULTI_GRID[bi] if si == 3 else layer_dil
                s, ds = 1, None
            if kind == "basic":
                convs = (Conv(f"{p}.conv1", inplanes, planes, 3, stride=s, dilation=first, bn=f"{p}.bn1", act="relu"),
                         Conv(f"{p}.conv2", planes, planes, 3, dilation=layer_dil, bn=f"{p}.bn2", act="relu"))
            else:
                convs = (Conv(f"{p}.conv1", inplanes, planes, 1, bn=f"{p}.bn1", act="relu"),
                         Conv(f"{p}.conv2", planes, planes, 3, stride=s, dilation=first, bn=f"{p}.bn2", act="relu"),
                         Conv(f"{p}.conv3", planes, planes * 4, 1, bn=f"{p}.bn3", act="relu"))
            blocks.append(Block(convs, ds))
            inplanes = planes * exp
    return stem, blocks, 512 * exp


FA_LEVELS = (32, 16, 8, 4)       # ffm_32 .. ffm_4 (td2_fa.py:56-63); level L reads the backbone tap feat<L>
FA_DK = 32                       # FAModule w_qs / w_ks output channels (td2_fa.py:339-341)
FA_OUT = 128                     # FAModule `smooth` output channels
FA_KEY_STRIDE = 3                # Encoding.maxpool_k/v of the td2_fanet tree: MaxPool2d(kernel 1, stride 3)


def _fa_backbone(prefix: str, backbone: str):
    """The FANet ResNet (Training/ptsemseg/models/td2_fanet/resnet.py:113-145): 7x7 s2 stem + BN + LeakyReLU, maxpool,
    four stages that ALL start with a stride-2 block (`[2, 2, 2, 2]`, :161-176), no dilation.  BasicBlock (:36-67):
    conv3x3(stride) BN LeakyReLU, conv3x3 BN, + shortcut, ReLU; Bottleneck (:69-108) likewise with 1x1/3x3/1x1."""
    kind, counts = BLOCKS[backbone]
    exp = 4 if kind == "bottleneck" else 1
    stem = [Conv(f"{prefix}.conv1", 3, 64, 7, stride=2, bn=f"{prefix}.bn1", act="leaky_relu")]
    blocks: List[Block] = []
    ends = []
    inplanes = 64
    for si, (planes, count) in enumerate(zip(STAGE_PLANES, counts)):
        for bi in range(count):
            p = f"{prefix}.layer{si + 1}.{bi}"
            s = 2 if bi == 0 else 1
            ds = None
            if s != 1 or inplanes != planes * exp:
                ds = Conv(f"{p}.downsample.0", inplanes, planes * exp, 1, stride=s, bn=f"{p}.downsample.1")
            if kind == "basic":
                convs = (Conv(f"{p}.conv1", inplanes, planes, 3, stride=s, bn=f"{p}.bn1", act="leaky_relu"),
                         Conv(f"{p}.conv2", planes, planes, 3, bn=f"{p}.bn2", act="relu"))
            else:
                convs = (Conv(f"{p}.conv1", inplanes, planes, 1, bn=f"{p}.bn1", act="leaky_relu"),
                         Conv(f"{p}.conv2", planes, planes, 3, stride=s, bn=f"{p}.bn2", act="leaky_relu"),
                         Conv(f"{p}.conv3", planes, planes * 4, 1, bn=f"{p}.bn3", act="relu"))
            blocks.append(Block(convs, ds))
            inplanes = planes * exp
        ends.append(len(blocks) - 1)
    return stem, blocks, tuple(ends), exp


def fa_module_convs(m: ModelArch, level: int, idx: int) -> Dict[str, Conv]:
    """FAModule (td2_fa.py:334-349) of pyramid level `level` in sub-network `idx`: ConvBNReLU = conv without bias +
    norm_layer(activation=...).  `up` is a 1x1 conv with padding 1 (:348): the engine handles the 1-pixel frame."""
    c = STAGE_PLANES[3 - FA_LEVELS.index(level)] * (m.c_exp)
    f = f"ffm_{level}_{idx}"

    def cbr(name, cout, k=1, act="leaky_relu"):
        return Conv(f"{f}.{name}.conv", c, cout, k, bn=f"{f}.{name}.bn", act=act)

    return dict(w_qs=cbr("w_qs", FA_DK, act="none"), w_ks=cbr("w_ks", FA_DK, act="none"), w_vs=cbr("w_vs", c),
                latlayer3=cbr("latlayer3", c), up=cbr("up", c // 2), smooth=cbr("smooth", FA_OUT, 3))


def fa_head_convs(m: ModelArch, name: str, cin: int, mid: int) -> List[Conv]:
    """FPNOutput (td2_fa.py:306-317): ConvBNReLU 3x3 (LeakyReLU) -> conv1x1 without bias."""
    return [Conv(f"{name}.conv.conv", cin, mid, 3, bn=f"{name}.conv.bn", act="leaky_relu"),
            Conv(f"{name}.conv_out", mid, m.nclass, 1)]


def build_arch(arch: str, backbone: str, nclass: int) -> ModelArch:
    if arch == "td2_fa":
        if backbone not in ("resnet18", "resnet34", "resnet50"):
            raise RuntimeError("unknown backbone: {}".format(backbone))      # td2_fa.py:48-49
        # two sub-networks, no FIFO (both run on every call); Encoding(256, 64, 256), Attention(256, 64) (:66-69)
        m = ModelArch(arch, backbone, 2, 0, 2 * FA_OUT, 64, 2 * FA_OUT, 2 * FA_OUT, nclass)
        for idx in (1, 2):
            m.stems[idx], m.stages[idx], m.stage_ends, exp = _fa_backbone(f"pretrained{idx}", backbone)
        m.c_exp = exp
        return m
    if arch == "pspnet":
        if backbone not in BLOCKS:
            raise RuntimeError("unknown backbone: {}".format(backbone))      # pspnet.py:67-68
        c4 = 512 * (4 if BLOCKS[backbone][0] == "bottleneck" else 1)
        # one path, no FIFO, no attention: d_k / d_v unused; head_mid = PSPHead inter_channels (pspnet.py:105)
        m = ModelArch(arch, backbone, 1, 0, c4, 0, c4, c4 // 4, nclass)
        m.stems[1], m.stages[1], _ = _backbone("pretrained", backbone)
        return m
    if backbone not in BLOCKS or backbone == "resnet101":
        raise RuntimeError("Four branch model only support ResNet18 amd ResNet34")  # td4_psp18.py:68
    paths = 4 if arch == "td4_psp18" else 2
    exp = 4 if BLOCKS[backbone][0] == "bottleneck" else 1
    c4 = 512 * exp
    d_v = c4 if arch == "td4_psp18" else c4 // 4         # td4_psp18.py:85 / td2_psp50.py:79
    head_mid = d_v // (4 if arch == "td4_psp18" else 2)  # FCNHead chn_down, :112 / td2 :88
    m = ModelArch(arch, backbone, paths, 3 if paths == 4 else 1, c4, 64, d_v, head_mid, nclass)
    for path in range(1, paths + 1):
        stem, blocks, _ = _backbone(f"pretrained{path}", backbone)
        m.stems[path], m.stages[path] = stem, blocks
    return m


def parameter_table(m: ModelArch, ln_shape=(97, 193)):
    """Ordered {state-dict key: (shape, kind)} with kind in {'param', 'buffer', 'long_buffer'}."""
    t: Dict[str, tuple] = {}

    def conv(c: Conv):
        t[c.name + ".weight"] = ((c.cout, c.cin, c.k, c.k), "param")
        if c.bias:
            t[c.name + ".bias"] = ((c.cout,), "param")
        if c.bn:
            bn(c.bn, c.cout)

    def bn(p, ch):
        t[p + ".weight"], t[p + ".bias"] = ((ch,), "param"), ((ch,), "param")
        t[p + ".running_mean"], t[p + ".running_var"] = ((ch,), "buffer"), ((ch,), "buffer")
        t[p + ".num_batches_tracked"] = ((), "long_buffer")

    if m.arch == "td2_fa":   # td2_fa.py:53-79 (the FANet ResNet has no fc layer)
        for idx in (1, 2):
            for c in m.stems[idx]:
                conv(c)
            for b in m.stages[idx]:
                for c in b.convs:
                    conv(c)
                if b.downsample:
                    conv(b.downsample)
            for level in FA_LEVELS:
                for c in fa_module_convs(m, level, idx).values():
                    conv(c)
            for cs in encoding_convs(m, idx).values():
                for c in cs:
                    conv(c)
            conv(fc_conv(m, f"atn{idx}"))
            t[f"layer_norm{idx}.ln.weight"] = (tuple(ln_shape), "param")
            t[f"layer_norm{idx}.ln.bias"] = (tuple(ln_shape), "param")
            for c in fa_head_convs(m, f"head{idx}", m.d_v, m.head_mid) + fa_head_convs(m, f"head_aux{idx}", FA_OUT, 64):
                conv(c)
        return t
    for path in range(1, m.paths + 1):
        for c in m.stems[path]:
            conv(c)
        for b in m.stages[path]:
            for c in b.convs:
                conv(c)
            if b.downsample:
                conv(b.downsample)
        t[f"{m.prefix(path)}.fc.weight"] = ((1000, m.c4), "param")  # resnet.py:160, never used by forward
        t[f"{m.prefix(path)}.fc.bias"] = ((1000,), "param")
    if m.arch == "pspnet":
        for c in psp_convs(m, 1) + head_convs(m, 1):
            conv(c)
        return t
    for path in range(1, m.paths + 1):
        for c in psp_convs(m, path, full=True):
            conv(c)
        for c in encoding_convs(m, path).values():
            for cc in c:
                conv(cc)
        for name in m.hop_modules(path):
            conv(fc_conv(m, name))
        t[f"layer_norm{path}.ln.weight"] = (tuple(ln_shape), "param")
        t[f"layer_norm{path}.ln.bias"] = (tuple(ln_shape), "param")
        for c in head_convs(m, path):
            conv(c)
    return t


def psp_convs(m: ModelArch, path: int, full=False) -> List[Conv]:
    """PyramidPooling conv1..4 (td4_psp18.py:255-266; pspnet.py:131-142, where the module is element 0 of
    PSPHead.conv5): 1x1, c4 -> c4/4, BN, ReLU."""
    p = "head.conv5.0" if m.arch == "pspnet" else f"psp{path}"
    return [Conv(f"{p}.conv{i}.0", m.c4, m.c4 // 4, 1, bn=f"{p}.conv{i}.1", act="relu") for i in range(1, 5)]


def encoding_convs(m: ModelArch, path: int) -> Dict[str, List[Conv]]:
    """Encoding (transformer.py:10-26): w_qs / w_ks = [1x1 c4->64 +b, BN, LeakyReLU ; 1x1 64->64 +b],
    w_vs = [1x1 c4->d_v +b]."""
    e = f"enc{path}"
    out = {}
    for w in ("w_qs", "w_ks"):
        out[w] = [Conv(f"{e}.{w}.0.conv", m.c4, m.d_k, 1, bn=f"{e}.{w}.0.bn", bias=True, act="leaky_relu"),
                  Conv(f"{e}.{w}.1.conv", m.d_k, m.d_k, 1, bias=True)]
    out["w_vs"] = [Conv(f"{e}.w_vs.0.conv", m.c4, m.d_v, 1, bias=True)]
    return out


def fc_conv(m: ModelArch, module: str) -> Conv:
    """Attention.fc (transformer.py:67): 1x1 d_v -> d_v with bias, no norm, no activation."""
    return Conv(f"{module}.fc.0.conv", m.d_v, m.d_v, 1, bias=True)


def head_convs(m: ModelArch, path: int) -> List[Conv]:
    """FCNHead.conv5 (td4_psp18.py:295-299); for pspnet the rest of PSPHead.conv5 behind the pyramid
    (pspnet.py:108-113): conv3x3 2*C4 -> C4/4, BN, ReLU, Dropout2d, conv1x1 -> nclass."""
    if m.arch == "pspnet":
        return [Conv("head.conv5.1", 2 * m.c4, m.head_mid, 3, bn="head.conv5.2", act="relu"),
                Conv("head.conv5.5", m.head_mid, m.nclass, 1, bias=True)]
    h = f"head{path}.conv5"
    return [Conv(f"{h}.0", m.d_v, m.head_mid, 3, bn=f"{h}.1", act="relu"),
            Conv(f"{h}.4", m.head_mid, m.nclass, 1, bias=True)]


def feature_hw(h: int, w: int) -> Tuple[int, int]:
    """Three stride-2 stages with out = floor((in-1)/2)+1."""
    for _ in range(3):
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    return h, w
