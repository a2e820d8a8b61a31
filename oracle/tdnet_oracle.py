"""CPU oracle for the TDNet per-frame inference hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, not the product: only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import it.  tdnet_b200/ never does.

It restates, as stateless functions over a flat state-dict, what the reference computes in
/root/reference/Testing/model/pspnet/{td4_psp18,td2_psp50,pspnet,resnet,transformer}.py.  Every numeric
primitive of the reference is a PyTorch library call ("third-party arithmetic"), so
the restatement issues the *same* fp32 torch CPU primitives in the same order (conv2d ->
batch_norm -> relu, bmm -> div -> softmax -> bmm, layer_norm, interpolate); nothing is folded or
re-associated here.  That keeps it bit-comparable with the reference on the same machine.

Pinning: the reference holds no golden vectors or tests.  The pin is
tests/golden/*.npz, produced by tests/golden/make_golden.py, which imports the unmodified
reference from /root/reference/Testing, loads the same synthetic weights (tdnet_b200/synth.py),
runs it on CPU and stores its outputs; tests/test_oracle.py checks this file against those
outputs.  Parity status: pinned against outputs of the reference itself run in the build container.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

_BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, td4_psp18.py:11-24 subclasses it unchanged
_LN_EPS = 1e-5  # torch.nn.LayerNorm default, td4_psp18.py:306-312

# resnet.py:219-256  -> (block kind, blocks per stage)
_BACKBONES = {
    "resnet18": ("basic", (2, 2, 2, 2)),
    "resnet34": ("basic", (3, 4, 6, 3)),
    "resnet50": ("bottleneck", (3, 4, 6, 3)),
    "resnet101": ("bottleneck", (3, 4, 23, 3)),
}


def _bn(sd, p, x, act="none"):
    """Eval-mode BatchNorm2d wrapper, td4_psp18.py:11-24 ('leaky_relu' -> nn.LeakyReLU(), slope 0.01)."""
    y = F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                     sd[p + ".bias"], training=False, eps=_BN_EPS)
    if act == "relu":
        return F.relu(y)
    if act == "leaky_relu":
        return F.leaky_relu(y, 0.01)
    return y


def _conv(sd, p, x, stride=1, padding=0, dilation=1):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding,
                    dilation=dilation)


def stage_plan(backbone):
    """Per-block (stride, conv1/3x3 dilation, conv2 dilation, has_downsample) for the four stages of
    the dilated, multi-grid ResNet: resnet.py:139-147 (dilated=True, multi_grid=True) and
    _make_layer resnet.py:170-202."""
    kind, counts = _BACKBONES[backbone]
    exp = 4 if kind == "bottleneck" else 1
    inplanes = 128 if kind == "bottleneck" else 64  # deep_base stem ends at 128 ch, resnet.py:117
    plan = []
    for si, (planes, nblk) in enumerate(zip((64, 128, 256, 512), counts)):
        stride = 2 if si == 1 else 1
        layer_dil = (1, 1, 2, 4)[si]
        multi_grid = si == 3
        blocks = []
        for bi in range(nblk):
            if bi == 0:
                ds = stride != 1 or inplanes != planes * exp
                if multi_grid:
                    d1 = 4
                elif layer_dil in (1, 2):
                    d1 = 1
                else:
                    d1 = 2
                blocks.append(dict(stride=stride, d1=d1, d2=layer_dil, downsample=ds))
                inplanes = planes * exp
            else:
                d1 = (4, 8, 16)[bi] if multi_grid else layer_dil
                blocks.append(dict(stride=1, d1=d1, d2=layer_dil, downsample=False))
        plan.append(blocks)
    return kind, plan


def _basic_block(sd, p, x, b):
    """BasicBlock.forward resnet.py:43-59; conv1 uses `dilation`, conv2 `previous_dilation` (:29-36)."""
    out = _conv(sd, p + ".conv1", x, stride=b["stride"], padding=b["d1"], dilation=b["d1"])
    out = _bn(sd, p + ".bn1", out, "relu")
    out = _conv(sd, p + ".conv2", out, padding=b["d2"], dilation=b["d2"])
    out = _bn(sd, p + ".bn2", out)
    res = x
    if b["downsample"]:
        res = _bn(sd, p + ".downsample.1", _conv(sd, p + ".downsample.0", x, stride=b["stride"]))
    return F.relu(out + res)


def _bottleneck(sd, p, x, b):
    """Bottleneck.forward resnet.py:91-111; only the 3x3 is strided/dilated (:71-73)."""
    out = _bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x), "relu")
    out = _conv(sd, p + ".conv2", out, stride=b["stride"], padding=b["d1"], dilation=b["d1"])
    out = _bn(sd, p + ".bn2", out, "relu")
    out = _bn(sd, p + ".bn3", _conv(sd, p + ".conv3", out))
    res = x
    if b["downsample"]:
        res = _bn(sd, p + ".downsample.1", _conv(sd, p + ".downsample.0", x, stride=b["stride"]))
    return F.relu(out + res)


def resnet_c4(sd, p, img, backbone, taps=None):
    """ResNet.forward resnet.py:204-215 -> c4 [n, 512*exp, H/8, W/8]."""
    kind, plan = stage_plan(backbone)
    if kind == "bottleneck":  # deep_base stem, resnet.py:122-131
        x = _bn(sd, p + ".conv1.1", _conv(sd, p + ".conv1.0", img, stride=2, padding=1), "relu")
        x = _bn(sd, p + ".conv1.4", _conv(sd, p + ".conv1.3", x, padding=1), "relu")
        x = _conv(sd, p + ".conv1.6", x, padding=1)
    else:  # resnet.py:133-134
        x = _conv(sd, p + ".conv1", img, stride=2, padding=3)
    x = _bn(sd, p + ".bn1", x, "relu")
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    if taps is not None:
        taps["stem"] = x
    block = _bottleneck if kind == "bottleneck" else _basic_block
    for si, blocks in enumerate(plan):
        for bi, b in enumerate(blocks):
            x = block(sd, f"{p}.layer{si + 1}.{bi}", x, b)
        if taps is not None:
            taps[f"layer{si + 1}"] = x
    return x


def pyramid_slice(sd, p, c4, pid, path_num=2):
    """PyramidPooling.forward td4_psp18.py:271-284: four pooled 1x1-conv branches, bilinear
    (align_corners) back to H8xW8, then this path's channel slices, concatenated."""
    n, c, h, w = c4.shape
    feats = []
    for i, bins in enumerate((1, 2, 3, 6)):
        f = F.adaptive_avg_pool2d(c4, bins)
        f = _bn(sd, f"{p}.conv{i + 1}.1", _conv(sd, f"{p}.conv{i + 1}.0", f), "relu")
        f = F.interpolate(f, (h, w), mode="bilinear", align_corners=True)
        lo, hi = pid * c // (path_num * 4), (pid + 1) * c // (path_num * 4)
        feats.append(f[:, lo:hi])
    x = c4[:, pid * c // path_num:(pid + 1) * c // path_num]
    return torch.cat([x] + feats, 1)


def _proj_qk(sd, p, fea):
    """w_qs / w_ks: conv1x1(+b) -> BN -> LeakyReLU -> conv1x1(+b); transformer.py:18-22,142-161."""
    x = _bn(sd, p + ".0.bn", _conv(sd, p + ".0.conv", fea), "leaky_relu")
    return _conv(sd, p + ".1.conv", x)


def _tokens(x):
    n, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).contiguous().view(n, h * w, c)


def encode_full(sd, p, z):
    """Encoding.forward(pre=False) transformer.py:52-56 -> q [n,P,64], v [n,d_v,H8,W8]."""
    v = _conv(sd, p + ".w_vs.0.conv", z)
    q = _tokens(_proj_qk(sd, p + ".w_qs", z))
    return q, v


def encode_sub(sd, p, z):
    """Encoding.forward(pre=True) transformer.py:34-50: MaxPool2d(kernel 1, stride 4) == z[:,:,::4,::4]
    (transformer.py:26), then K, V, Q projections, all as [n, P', c] token matrices."""
    zs = F.max_pool2d(z, kernel_size=1, stride=4, padding=0)
    k = _tokens(_proj_qk(sd, p + ".w_ks", zs))
    v = _tokens(_conv(sd, p + ".w_vs.0.conv", zs))
    q = _tokens(_proj_qk(sd, p + ".w_qs", zs))
    return q, k, v


def attention_hop(sd, p, k_src, v_src, q_tgt, fea_size=None):
    """Attention.forward transformer.py:71-92 with ScaledDotProductAttention :126-139
    (temperature = sqrt(d_k) = 8.0, softmax over keys, dropout = identity in eval) and the fc 1x1
    conv applied per token (:84-86)."""
    d_k = q_tgt.shape[-1]
    attn = torch.bmm(q_tgt, k_src.transpose(1, 2))
    attn = attn / float(d_k) ** 0.5
    attn = torch.softmax(attn, dim=2)
    out = torch.bmm(attn, v_src)
    n, pq, c = out.shape
    out = _conv(sd, p + ".fc.0.conv", out.view(n * pq, c, 1, 1)).view(n, pq, c)
    if fea_size is not None:
        _, _, h, w = fea_size
        out = out.permute(0, 2, 1).contiguous().view(n, -1, h, w)
    return out


def layer_norm_hw(sd, p, x):
    """Layer_Norm td4_psp18.py:306-312: nn.LayerNorm([H8, W8]) -> per (n, c) statistics over the
    map, affine of shape [H8, W8] shared by all channels."""
    w = sd[p + ".ln.weight"]
    return F.layer_norm(x, tuple(w.shape), w, sd[p + ".ln.bias"], _LN_EPS)


def fcn_head(sd, p, x):
    """FCNHead td4_psp18.py:287-302: conv3x3 (no bias) -> BN -> ReLU -> Dropout2d(id) -> conv1x1(+b)."""
    y = _bn(sd, p + ".conv5.1", _conv(sd, p + ".conv5.0", x, padding=1), "relu")
    return _conv(sd, p + ".conv5.4", y)


class TDOracle:
    """Restates td4_psp18.forward (td4_psp18.py:216-229, forward_path1..4 :137-212, buffer_contral
    :123-134) and td2_psp50.forward (td2_psp50.py:146-155, :112-143, :98-109).

    arch: 'td4_psp18' (4 paths, FIFO depth 3, three attention hops) or 'td2_psp50' (2 paths, FIFO
    depth 1, one hop).  `taps` holds the intermediates of the last forward for per-stage tests.
    """

    def __init__(self, arch, state_dict, backbone=None, nclass=19):
        assert arch in ("td4_psp18", "td2_psp50")
        self.arch = arch
        self.paths = 4 if arch == "td4_psp18" else 2
        self.depth = 3 if arch == "td4_psp18" else 1
        self.backbone = backbone or ("resnet18" if arch == "td4_psp18" else "resnet50")
        self.sd = state_dict
        self.nclass = nclass
        self.Q_queue, self.K_queue, self.V_queue = [], [], []
        self.taps = {}

    def reset(self):
        self.Q_queue, self.K_queue, self.V_queue = [], [], []

    def _hop_names(self, path):
        if self.arch == "td2_psp50":
            return [f"atn{path}"]
        return [f"atn{path}_{(path + j) % 4 + 1}" for j in range(3)]

    @torch.no_grad()
    def forward(self, img, pos_id=0):
        sd, t = self.sd, {}
        path = pos_id + 1
        h, w = img.shape[2:]
        c4 = resnet_c4(sd, f"pretrained{path}", img, self.backbone, t)
        z = pyramid_slice(sd, f"psp{path}", c4, pid=(path - 1) % 2, path_num=2)
        q_cur, v_cur = encode_full(sd, f"enc{path}", z)
        t.update(c4=c4, z=z, q_cur=q_cur, v_cur=v_cur)
        if len(self.Q_queue) < self.depth:
            fused = v_cur
        else:
            names = self._hop_names(path)
            if self.arch == "td2_psp50":
                v_last = attention_hop(sd, names[0], self.K_queue[0], self.V_queue[0], q_cur, z.shape)
            else:
                v2 = attention_hop(sd, names[0], self.K_queue[0], self.V_queue[0], self.Q_queue[1])
                v3 = attention_hop(sd, names[1], self.K_queue[1], v2 + self.V_queue[1], self.Q_queue[2])
                v_last = attention_hop(sd, names[2], self.K_queue[2], v3 + self.V_queue[2], q_cur, z.shape)
                t.update(v2=v2, v3=v3)
            t["v_prop"] = v_last
            fused = v_last + v_cur
        normed = layer_norm_hw(sd, f"layer_norm{path}", fused)
        low = fcn_head(sd, f"head{path}", normed)
        t.update(normed=normed, head=low)
        q, k, v = encode_sub(sd, f"enc{path}", z)
        self.Q_queue.append(q), self.K_queue.append(k), self.V_queue.append(v)
        if len(self.Q_queue) > self.depth:
            self.Q_queue.pop(0), self.K_queue.pop(0), self.V_queue.pop(0)
        t.update(q_sub=q, k_sub=k, v_sub=v)
        self.taps = t
        return F.interpolate(low, (h, w), mode="bilinear", align_corners=True)

    __call__ = forward


def _template_helpers(sd):
    def conv(p, co, ci, k, bias=False):
        sd[p + ".weight"] = torch.zeros(co, ci, k, k)
        if bias:
            sd[p + ".bias"] = torch.zeros(co)

    def bn(p, c):
        sd[p + ".weight"], sd[p + ".bias"] = torch.zeros(c), torch.zeros(c)
        sd[p + ".running_mean"], sd[p + ".running_var"] = torch.zeros(c), torch.zeros(c)
        sd[p + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    return conv, bn


def _backbone_template(sd, p, backbone):
    """State-dict entries of one ResNet (resnet.py:116-168) under prefix `p`; returns c4 channels."""
    conv, bn = _template_helpers(sd)
    kind, plan = stage_plan(backbone)
    exp = 4 if kind == "bottleneck" else 1
    if kind == "bottleneck":
        conv(p + ".conv1.0", 64, 3, 3), bn(p + ".conv1.1", 64)
        conv(p + ".conv1.3", 64, 64, 3), bn(p + ".conv1.4", 64)
        conv(p + ".conv1.6", 128, 64, 3), bn(p + ".bn1", 128)
        inpl = 128
    else:
        conv(p + ".conv1", 64, 3, 7), bn(p + ".bn1", 64)
        inpl = 64
    for si, blocks in enumerate(plan):
        planes = (64, 128, 256, 512)[si]
        for bi, b in enumerate(blocks):
            q = f"{p}.layer{si + 1}.{bi}"
            if kind == "bottleneck":
                conv(q + ".conv1", planes, inpl, 1), bn(q + ".bn1", planes)
                conv(q + ".conv2", planes, planes, 3), bn(q + ".bn2", planes)
                conv(q + ".conv3", planes * 4, planes, 1), bn(q + ".bn3", planes * 4)
            else:
                conv(q + ".conv1", planes, inpl, 3), bn(q + ".bn1", planes)
                conv(q + ".conv2", planes, planes, 3), bn(q + ".bn2", planes)
            if b["downsample"]:
                conv(q + ".downsample.0", planes * exp, inpl, 1), bn(q + ".downsample.1", planes * exp)
            inpl = planes * exp
    c4 = 512 * exp
    sd[p + ".fc.weight"], sd[p + ".fc.bias"] = torch.zeros(1000, c4), torch.zeros(1000)
    return c4


def state_dict_template(arch, backbone=None, nclass=19, ln_shape=(97, 193)):
    """Key -> zero tensor of the right shape for the reference model's state_dict (strict=True load,
    td4_psp18.py:236-237), derived from the architecture alone so that the GPU box (which has no
    /root/reference) can synthesise weights.  Checked against the real reference in make_golden.py."""
    if arch == "pspnet":
        return pspnet_state_dict_template(backbone or "resnet101", nclass)
    paths = 4 if arch == "td4_psp18" else 2
    backbone = backbone or ("resnet18" if arch == "td4_psp18" else "resnet50")
    sd = {}
    conv, bn = _template_helpers(sd)
    for path in range(1, paths + 1):
        c4 = _backbone_template(sd, f"pretrained{path}", backbone)
    d_v = c4 if arch == "td4_psp18" else c4 // 4
    inter = d_v // (4 if arch == "td4_psp18" else 2)
    for path in range(1, paths + 1):
        for i in range(1, 5):
            conv(f"psp{path}.conv{i}.0", c4 // 4, c4, 1), bn(f"psp{path}.conv{i}.1", c4 // 4)
    for path in range(1, paths + 1):
        e = f"enc{path}"
        for w in ("w_qs", "w_ks"):
            conv(f"{e}.{w}.0.conv", 64, c4, 1, True), bn(f"{e}.{w}.0.bn", 64)
            conv(f"{e}.{w}.1.conv", 64, 64, 1, True)
        conv(f"{e}.w_vs.0.conv", d_v, c4, 1, True)
    probe = TDOracle(arch, {}, backbone)
    for path in range(1, paths + 1):
        for name in probe._hop_names(path):
            conv(f"{name}.fc.0.conv", d_v, d_v, 1, True)
    for path in range(1, paths + 1):
        sd[f"layer_norm{path}.ln.weight"] = torch.zeros(*ln_shape)
        sd[f"layer_norm{path}.ln.bias"] = torch.zeros(*ln_shape)
    for path in range(1, paths + 1):
        conv(f"head{path}.conv5.0", inter, d_v, 3), bn(f"head{path}.conv5.1", inter)
        conv(f"head{path}.conv5.4", nclass, inter, 1, True)
    return sd


# ------------------------------------------------------------------------------------------------
# Single-path PSPNet comparison model: Testing/model/pspnet/pspnet.py
# ------------------------------------------------------------------------------------------------
def psp_head(sd, p, c4):
    """PSPHead.forward pspnet.py:102-115 = conv5 Sequential: PyramidPooling (:118-157, all channels of
    the four pooled branches, concatenated behind c4) -> conv3x3 2*C4 -> C4/4 (no bias) -> BN -> ReLU ->
    Dropout2d (identity in eval) -> conv1x1 -> nclass (+bias)."""
    n, c, h, w = c4.shape
    feats = [c4]
    for i, bins in enumerate((1, 2, 3, 6)):
        f = F.adaptive_avg_pool2d(c4, bins)
        f = _bn(sd, f"{p}.conv5.0.conv{i + 1}.1", _conv(sd, f"{p}.conv5.0.conv{i + 1}.0", f), "relu")
        feats.append(F.interpolate(f, (h, w), mode="bilinear", align_corners=True))
    z = torch.cat(feats, 1)
    y = _bn(sd, p + ".conv5.2", _conv(sd, p + ".conv5.1", z, padding=1), "relu")
    return z, _conv(sd, p + ".conv5.5", y)


class PSPNetOracle:
    """Restates pspnet.forward (pspnet.py:73-89): only the LAST image of the batch is segmented
    (`x = x[-1:]`), `pos_id` is accepted and ignored; backbone -> PSPHead -> bilinear upsample
    (align_corners=True) to the input size.  Stateless between frames."""
    paths, depth = 1, 0

    def __init__(self, state_dict, backbone="resnet101", nclass=19):
        assert backbone in _BACKBONES
        self.sd, self.backbone, self.nclass = state_dict, backbone, nclass
        self.taps = {}

    @torch.no_grad()
    def forward(self, img, pos_id=None):
        x = img[-1:]
        h, w = x.shape[2:]
        t = {}
        c4 = resnet_c4(self.sd, "pretrained", x, self.backbone, t)
        z, low = psp_head(self.sd, "head", c4)
        t.update(c4=c4, z=z, head=low)
        self.taps = t
        return F.interpolate(low, (h, w), mode="bilinear", align_corners=True)

    __call__ = forward


def pspnet_state_dict_template(backbone="resnet101", nclass=19):
    """State-dict layout of pspnet.pspnet (pspnet.py:31-71): `pretrained.*`, `head.conv5.0.conv{1..4}.{0,1}`
    (PyramidPooling), `head.conv5.1` (3x3), `head.conv5.2` (BN), `head.conv5.5` (classifier)."""
    sd = {}
    conv, bn = _template_helpers(sd)
    c4 = _backbone_template(sd, "pretrained", backbone)
    for i in range(1, 5):
        conv(f"head.conv5.0.conv{i}.0", c4 // 4, c4, 1), bn(f"head.conv5.0.conv{i}.1", c4 // 4)
    conv("head.conv5.1", c4 // 4, 2 * c4, 3), bn("head.conv5.2", c4 // 4)
    conv("head.conv5.5", nclass, c4 // 4, 1, True)
    return sd
