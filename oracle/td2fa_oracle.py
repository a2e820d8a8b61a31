"""CPU oracle for the TD2-FANet inference path.  TEST INFRASTRUCTURE ONLY.

Only tests/ (and tests/golden/make_golden_fanet.py for the key/shape check) may import this file; tdnet_b200/ never
does.  It restates, as stateless functions over a flat state dict, what the reference computes in eval mode in
/root/reference/Training/ptsemseg/models/td2_fanet/{td2_fa,resnet,transformer}.py, issuing the same fp32 torch CPU
primitives in the same order (every numeric primitive of the reference is a torch library call).

Pinning: the reference holds no tests or golden vectors for this model.  tests/golden/td2fa_*.npz are outputs of
the reference module itself (tests/golden/make_golden_fanet.py imports it unmodified from
/root/reference/Training; only import-time obstacles are stubbed there: the absent `encoding` package that provides
the norm layer, the `pdb.set_trace()` in the constructor and the ImageNet download), and tests/test_oracle.py
checks this file against them.  Parity status: pinned against outputs of the reference itself.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .tdnet_oracle import _LN_EPS, _bn, _conv, _template_helpers, _tokens, attention_hop

# resnet.py:161-186 (td2_fanet): block kind, blocks per stage; every stage has stride 2 (`[2, 2, 2, 2]`)
_BACKBONES = {
    "resnet18": ("basic", (2, 2, 2, 2)),
    "resnet34": ("basic", (3, 4, 6, 3)),
    "resnet50": ("bottleneck", (3, 4, 6, 3)),
}
_PLANES = (64, 128, 256, 512)
FA_LEVELS = (32, 16, 8, 4)      # ffm_32 .. ffm_4 (td2_fa.py:56-63)
FA_KEY_STRIDE = 3               # Encoding.maxpool_k / maxpool_v, transformer.py:26-27 (td2_fanet)


def _block_plan(backbone):
    """(in_chan, out_chan, stride, has_downsample) per block, create_layer resnet.py:128-133."""
    kind, counts = _BACKBONES[backbone]
    exp = 4 if kind == "bottleneck" else 1
    inpl = 64
    plan = []
    for planes, count in zip(_PLANES, counts):
        blocks = []
        for bi in range(count):
            stride = 2 if bi == 0 else 1
            ds = inpl != planes * exp or stride != 1      # resnet.py:46, 82
            blocks.append(dict(cin=inpl, planes=planes, stride=stride, downsample=ds))
            inpl = planes * exp
        plan.append(blocks)
    return kind, plan


def _shortcut(sd, p, x, b):
    if not b["downsample"]:
        return x
    return _bn(sd, p + ".downsample.1", _conv(sd, p + ".downsample.0", x, stride=b["stride"]))


def _basic_block(sd, p, x, b):
    """BasicBlock.forward resnet.py:53-67: conv3x3(stride) -> BN+LeakyReLU -> conv3x3 -> BN -> shortcut + out -> ReLU."""
    out = _bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x, stride=b["stride"], padding=1), "leaky_relu")
    out = _bn(sd, p + ".bn2", _conv(sd, p + ".conv2", out, padding=1))
    return F.relu(_shortcut(sd, p, x, b) + out)


def _bottleneck(sd, p, x, b):
    """Bottleneck.forward resnet.py:91-108: 1x1 -> 3x3(stride) -> 1x1, LeakyReLU after bn1/bn2, none after bn3."""
    out = _bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x), "leaky_relu")
    out = _bn(sd, p + ".bn2", _conv(sd, p + ".conv2", out, stride=b["stride"], padding=1), "leaky_relu")
    out = _bn(sd, p + ".bn3", _conv(sd, p + ".conv3", out))
    return F.relu(_shortcut(sd, p, x, b) + out)


def fa_resnet(sd, p, img, backbone):
    """ResNet.forward resnet.py:135-145 -> (feat4, feat8, feat16, feat32); with four stride-2 stages behind the
    stride-4 stem these are 1/8, 1/16, 1/32 and 1/64 of the input."""
    kind, plan = _block_plan(backbone)
    x = _bn(sd, p + ".bn1", _conv(sd, p + ".conv1", img, stride=2, padding=3), "leaky_relu")
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    block = _bottleneck if kind == "bottleneck" else _basic_block
    feats = []
    for si, blocks in enumerate(plan):
        for bi, b in enumerate(blocks):
            x = block(sd, f"{p}.layer{si + 1}.{bi}", x, b)
        feats.append(x)
    return feats


def _cbr(sd, p, x, act="leaky_relu", padding=0):
    """ConvBNReLU td2_fa.py:282-303: conv (no bias) -> norm_layer(out, activation=...)."""
    return _bn(sd, p + ".bn", _conv(sd, p + ".conv", x, padding=padding), act)


def fa_module(sd, p, feat, up_fea_in, up_flag, smf_flag):
    """FAModule.forward td2_fa.py:350-395.  Returns (up_feat or None, smooth_feat or None)."""
    query = _cbr(sd, p + ".w_qs", feat, "none")
    key = _cbr(sd, p + ".w_ks", feat, "none")
    value = _cbr(sd, p + ".w_vs", feat)
    n, c, h, w = feat.shape
    query = F.normalize(query.view(n, 32, -1).permute(0, 2, 1), p=2, dim=2, eps=1e-12)
    key = F.normalize(key.view(n, 32, -1), p=2, dim=1, eps=1e-12)
    value = value.view(n, c, -1).permute(0, 2, 1)
    f = torch.matmul(key, value)
    y = torch.matmul(query, f)
    y = y.permute(0, 2, 1).contiguous().view(n, c, h, w)
    p_feat = _cbr(sd, p + ".latlayer3", y) + feat
    if up_fea_in is not None:      # _upsample_add td2_fa.py:398-402
        p_feat = F.interpolate(up_fea_in, (h, w), mode="bilinear", align_corners=True) + p_feat
    up_feat = smooth = None
    if up_flag:                    # `up`: kernel 1 with padding 1 (td2_fa.py:348) -> the map grows by 2 pixels
        up_feat = _cbr(sd, p + ".up", p_feat, padding=1)
    if smf_flag and (up_fea_in is not None or not up_flag):
        smooth = _cbr(sd, p + ".smooth", p_feat, padding=1)
    return up_feat, smooth


def fa_subnet(sd, idx, img, backbone, taps=None):
    """One sub-network of forward_path1/2 (td2_fa.py:95-101): backbone -> four fast-attention modules top-down ->
    z = cat(upsampled smooth_16, smooth_4), 256 channels at the feat4 resolution (_upsample_cat :191-197)."""
    feat4, feat8, feat16, feat32 = fa_resnet(sd, f"pretrained{idx}", img, backbone)
    up32, _ = fa_module(sd, f"ffm_32_{idx}", feat32, None, True, True)
    up16, sm16 = fa_module(sd, f"ffm_16_{idx}", feat16, up32, True, True)
    up8, _ = fa_module(sd, f"ffm_8_{idx}", feat8, up16, True, False)
    _, sm4 = fa_module(sd, f"ffm_4_{idx}", feat4, up8, False, True)
    h, w = sm4.shape[2:]
    z = torch.cat([F.interpolate(sm16, (h, w), mode="bilinear", align_corners=True), sm4], dim=1)
    if taps is not None:
        taps.update({f"feat4_{idx}": feat4, f"feat32_{idx}": feat32, f"up32_{idx}": up32, f"up16_{idx}": up16,
                     f"sm16_{idx}": sm16, f"up8_{idx}": up8, f"sm4_{idx}": sm4, f"z_{idx}": z})
    return z


def _proj(sd, p, fea):
    """w_qs / w_ks of Encoding (transformer.py:18-22): conv1x1(+b) -> BN -> LeakyReLU -> conv1x1(+b)."""
    x = _bn(sd, p + ".0.bn", _conv(sd, p + ".0.conv", fea), "leaky_relu")
    return _conv(sd, p + ".1.conv", x)


def fa_encode_full(sd, p, z):
    """Encoding.forward(pre=False) transformer.py:47-53."""
    return _tokens(_proj(sd, p + ".w_qs", z)), _conv(sd, p + ".w_vs.0.conv", z)


def fa_encode_sub(sd, p, z):
    """Encoding.forward(pre=True) transformer.py:35-46: K and V on the full map, then MaxPool2d(kernel 1, stride 3)."""
    k = F.max_pool2d(_proj(sd, p + ".w_ks", z), kernel_size=1, stride=FA_KEY_STRIDE)
    v = F.max_pool2d(_conv(sd, p + ".w_vs.0.conv", z), kernel_size=1, stride=FA_KEY_STRIDE)
    return _tokens(k), _tokens(v)


def fpn_output(sd, p, x):
    """FPNOutput.forward td2_fa.py:314-317: ConvBNReLU 3x3 (LeakyReLU) -> conv1x1 without bias."""
    return _conv(sd, p + ".conv_out", _cbr(sd, p + ".conv", x, padding=1))


class TD2FAOracle:
    """Restates td2_fa.forward in eval mode (td2_fa.py:200-218, forward_path1 :87-131, forward_path2 :134-186).

    `frames` = [previous frame, current frame] (f_img[0], f_img[1]).  pos_id 0: sub-network 1 reads the current
    frame and supplies q / v, sub-network 2 reads the previous frame and supplies the keys / values; pos_id 1
    swaps the roles.  Stateless between calls (both sub-networks run on every call)."""
    paths = 2

    def __init__(self, state_dict, backbone="resnet18", nclass=19):
        assert backbone in _BACKBONES
        self.sd, self.backbone, self.nclass = state_dict, backbone, nclass
        self.taps = {}

    @torch.no_grad()
    def forward(self, frames, pos_id=0):
        if pos_id not in (0, 1):
            raise RuntimeError("Only Two Paths.")                       # td2_fa.py:207
        sd, t = self.sd, {}
        prev, cur = frames[0], frames[1]
        h, w = cur.shape[2:]
        if pos_id == 0:
            z1 = fa_subnet(sd, 1, cur, self.backbone, t)
            z2 = fa_subnet(sd, 2, prev, self.backbone, t)
            z_cur, z_prev, a, b = z1, z2, 1, 2
        else:
            z1 = fa_subnet(sd, 1, prev, self.backbone, t)
            z2 = fa_subnet(sd, 2, cur, self.backbone, t)
            z_cur, z_prev, a, b = z2, z1, 2, 1
        q, v = fa_encode_full(sd, f"enc{a}", z_cur)
        k_, v_ = fa_encode_sub(sd, f"enc{b}", z_prev)
        atn = attention_hop(sd, f"atn{a}", k_, v_, q, fea_size=z_cur.shape)
        lw = sd[f"layer_norm{a}.ln.weight"]
        normed = F.layer_norm(atn + v, tuple(lw.shape), lw, sd[f"layer_norm{a}.ln.bias"], _LN_EPS)
        low = fpn_output(sd, f"head{a}", normed)
        t.update(q=q, v=v, k_sub=k_, v_sub=v_, atn=atn, normed=normed, head=low)
        self.taps = t
        return F.interpolate(low, (h, w), mode="bilinear", align_corners=True)

    __call__ = forward


def td2fa_state_dict_template(backbone="resnet18", nclass=19, ln_shape=(96, 192)):
    """Key -> zero tensor for td2_fa.state_dict() (td2_fa.py:53-79); checked against the real module in
    tests/golden/make_golden_fanet.py.  head_aux{1,2} exist in the state dict but are not used by forward."""
    sd = {}
    conv, bn = _template_helpers(sd)
    kind, plan = _block_plan(backbone)
    exp = 4 if kind == "bottleneck" else 1

    def cbr(p, co, ci, k):
        conv(p + ".conv", co, ci, k), bn(p + ".bn", co)

    for idx in (1, 2):
        p = f"pretrained{idx}"
        conv(p + ".conv1", 64, 3, 7), bn(p + ".bn1", 64)
        for si, blocks in enumerate(plan):
            for bi, b in enumerate(blocks):
                q, pl, ci = f"{p}.layer{si + 1}.{bi}", b["planes"], b["cin"]
                if kind == "bottleneck":
                    conv(q + ".conv1", pl, ci, 1), bn(q + ".bn1", pl)
                    conv(q + ".conv2", pl, pl, 3), bn(q + ".bn2", pl)
                    conv(q + ".conv3", pl * 4, pl, 1), bn(q + ".bn3", pl * 4)
                else:
                    conv(q + ".conv1", pl, ci, 3), bn(q + ".bn1", pl)
                    conv(q + ".conv2", pl, pl, 3), bn(q + ".bn2", pl)
                if b["downsample"]:
                    conv(q + ".downsample.0", pl * exp, ci, 1), bn(q + ".downsample.1", pl * exp)
        for level, planes in zip(FA_LEVELS, (512, 256, 128, 64)):
            c, f = planes * exp, f"ffm_{level}_{idx}"
            cbr(f + ".w_qs", 32, c, 1), cbr(f + ".w_ks", 32, c, 1), cbr(f + ".w_vs", c, c, 1)
            cbr(f + ".latlayer3", c, c, 1), cbr(f + ".up", c // 2, c, 1), cbr(f + ".smooth", 128, c, 3)
        e = f"enc{idx}"
        for wn in ("w_qs", "w_ks"):
            conv(f"{e}.{wn}.0.conv", 64, 256, 1, True), bn(f"{e}.{wn}.0.bn", 64)
            conv(f"{e}.{wn}.1.conv", 64, 64, 1, True)
        conv(f"{e}.w_vs.0.conv", 256, 256, 1, True)
        conv(f"atn{idx}.fc.0.conv", 256, 256, 1, True)
        sd[f"layer_norm{idx}.ln.weight"] = torch.zeros(*ln_shape)
        sd[f"layer_norm{idx}.ln.bias"] = torch.zeros(*ln_shape)
        cbr(f"head{idx}.conv", 256, 256, 3), conv(f"head{idx}.conv_out", nclass, 256, 1)
        cbr(f"head_aux{idx}.conv", 64, 128, 3), conv(f"head_aux{idx}.conv_out", nclass, 64, 1)
    return sd


def fa_feature_hw(h, w, stages=3):
    """Map size after `stages` stride-2 steps with out = floor((in - 1) / 2) + 1: 3 -> feat4 (1/8), 6 -> feat32."""
    for _ in range(stages):
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    return h, w
