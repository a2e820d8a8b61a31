"""GPU: every C-ABI operator against the torch CPU primitive the reference calls at that site."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from common import max_abs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from tdnet_b200 import _cabi
    from tdnet_b200.engine import View
    lib = _cabi.load()
    return lib, _cabi, View, torch.device("cuda:0")


def nhwc(x):  # NCHW cpu -> NHWC cuda contiguous
    return x.permute(0, 2, 3, 1).contiguous().cuda()


def run_conv(env, x, w, scale=None, bias=None, residual=None, stride=1, dilation=1, act=0):
    lib, cabi, View, dev = env
    n, cin, h, wd = x.shape
    cout, _, k, _ = w.shape
    pad = dilation * (k - 1) // 2
    cin4 = (cin + 3) // 4 * 4
    xin = F.pad(x, (0, 0, 0, 0, 0, cin4 - cin))
    wk = F.pad(w.permute(0, 2, 3, 1), (0, cin4 - cin)).contiguous().cuda()
    xd = nhwc(xin)
    ref = F.conv2d(x, w, None, stride, pad, dilation)
    oh, ow = ref.shape[2:]
    out = torch.empty(n, oh, ow, cout, device=dev)
    d = cabi.Conv2dDesc()
    d.in_ = View(xd.view(-1), n, h, wd, cin4).ct()
    d.out = View(out.view(-1), n, oh, ow, cout).ct()
    keep = [xd, wk, out]
    if residual is not None:
        rd = nhwc(residual)
        keep.append(rd)
        d.residual = View(rd.view(-1), n, oh, ow, cout).ct()
        ref = None
    d.weight = wk.data_ptr()
    if scale is not None:
        sd_, bd_ = scale.cuda(), bias.cuda()
        keep += [sd_, bd_]
        d.scale, d.bias = sd_.data_ptr(), bd_.data_ptr()
    d.cout, d.kh, d.kw, d.stride, d.pad, d.dilation = cout, k, k, stride, pad, dilation
    d.act, d.leaky_slope, d.batch = act, 0.01, 1
    cabi.check(lib.tdn_conv2d(C.byref(d), None), "conv2d")
    torch.cuda.synchronize()
    return out.permute(0, 3, 1, 2).cpu()


CONV_CASES = [
    # cin, cout, k, stride, dil, h, w
    (3, 64, 7, 2, 1, 37, 53),     # stem (resnet.py:133)
    (64, 64, 3, 1, 1, 19, 27),
    (64, 128, 3, 2, 1, 19, 27),   # layer2.0.conv1
    (64, 128, 1, 2, 1, 19, 27),   # downsample
    (128, 256, 3, 1, 2, 13, 21),  # dilation 2
    (256, 96, 3, 1, 4, 13, 21),   # dilation 4 (wider than the map: mostly padding)
    (96, 40, 3, 1, 8, 13, 21),
    (128, 19, 1, 1, 1, 13, 21),   # classifier: cout not a multiple of 4
]


@pytest.mark.parametrize("cin,cout,k,stride,dil,h,w", CONV_CASES)
def test_conv2d_geometries(env, cin, cout, k, stride, dil, h, w):
    g = torch.Generator().manual_seed(cin * 1000 + cout + k)
    x = torch.randn(2, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    got = run_conv(env, x, wt, stride=stride, dilation=dil)
    ref = F.conv2d(x, wt, None, stride, dil * (k - 1) // 2, dil)
    assert got.shape == ref.shape
    assert max_abs(got, ref) < 2e-5


def test_conv2d_bn_residual_relu_epilogue(env):
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 64, 17, 23, generator=g)
    wt = torch.randn(64, 64, 3, 3, generator=g) / 24
    s, b = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g)
    r = torch.randn(1, 64, 17, 23, generator=g)
    got = run_conv(env, x, wt, scale=s, bias=b, residual=r, act=1)
    ref = F.relu(F.conv2d(x, wt, None, 1, 1) * s.view(1, -1, 1, 1) + b.view(1, -1, 1, 1) + r)
    assert max_abs(got, ref) < 2e-5
    got = run_conv(env, x, wt, scale=s, bias=b, act=2)
    ref = F.leaky_relu(F.conv2d(x, wt, None, 1, 1) * s.view(1, -1, 1, 1) + b.view(1, -1, 1, 1), 0.01)
    assert max_abs(got, ref) < 2e-5


def test_batched_gemm_both_weight_layouts(env):
    """bmm(q, k^T) and bmm(attn, v) (transformer.py:128,137) through tdn_conv2d with batch=2."""
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(3)
    q, k = torch.randn(2, 70, 64, generator=g), torch.randn(2, 30, 64, generator=g)
    qd, kd = q.cuda(), k.cuda()
    s = torch.zeros(2, 70, 32, device=dev)  # row pitch 32 (30 padded to a multiple of 4)
    d = cabi.Conv2dDesc()
    d.in_ = View(qd.view(-1), 1, 1, 70, 64).ct()
    d.out = View(s.view(-1), 1, 1, 70, 30, 70 * 32, 70 * 32, 32).ct()
    d.weight, d.cout, d.kh, d.kw, d.stride, d.dilation, d.batch = kd.data_ptr(), 30, 1, 1, 1, 1, 2
    d.in_batch_stride, d.out_batch_stride, d.weight_batch_stride = 70 * 64, 70 * 32, 30 * 64
    cabi.check(lib.tdn_conv2d(C.byref(d), None))
    torch.cuda.synchronize()
    assert max_abs(s[:, :, :30].cpu(), torch.bmm(q, k.transpose(1, 2))) < 2e-5
    assert float(s[:, :, 30:].abs().max()) == 0.0
    cabi.check(lib.tdn_softmax_rows(s.data_ptr(), 140, 30, 32, C.c_float(0.125), None))
    torch.cuda.synchronize()
    attn = torch.softmax(torch.bmm(q, k.transpose(1, 2)) / 8.0, dim=2)
    assert max_abs(s[:, :, :30].cpu(), attn) < 1e-6
    v = torch.randn(2, 30, 48, generator=g)
    vd = torch.zeros(2, 32, 48, device=dev)
    vd[:, :30] = v.cuda()
    res = torch.randn(2, 70, 48, generator=g)
    rd = res.cuda()
    o = torch.empty(2, 70, 48, device=dev)
    d = cabi.Conv2dDesc()
    d.in_ = View(s.view(-1), 1, 1, 70, 32).ct()
    d.out = View(o.view(-1), 1, 1, 70, 48).ct()
    d.residual = View(rd.view(-1), 1, 1, 70, 48).ct()
    d.weight, d.cout, d.kh, d.kw, d.stride, d.dilation, d.batch = vd.data_ptr(), 48, 1, 1, 1, 1, 2
    d.weight_kn = 1
    d.in_batch_stride, d.out_batch_stride, d.weight_batch_stride = 70 * 32, 70 * 48, 32 * 48
    d.residual_batch_stride = 70 * 48
    cabi.check(lib.tdn_conv2d(C.byref(d), None))
    torch.cuda.synchronize()
    assert max_abs(o.cpu(), torch.bmm(attn, v) + res) < 2e-5


def test_image_layout_and_maxpool(env):
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(5)
    img = torch.randn(2, 3, 21, 34, generator=g)
    imgd = img.cuda()
    o = torch.empty(2, 21, 34, 4, device=dev)
    t = View(o.view(-1), 2, 21, 34, 4).ct()
    cabi.check(lib.tdn_image_to_nhwc(imgd.data_ptr(), 2, 3, 21, 34, C.byref(t), None))
    torch.cuda.synchronize()
    assert torch.equal(o[..., :3].cpu(), img.permute(0, 2, 3, 1)) and float(o[..., 3].abs().max()) == 0
    x = torch.randn(2, 64, 21, 34, generator=g)
    xd = nhwc(x)
    y = torch.empty(2, 11, 17, 64, device=dev)
    ti, to = View(xd.view(-1), 2, 21, 34, 64).ct(), View(y.view(-1), 2, 11, 17, 64).ct()
    cabi.check(lib.tdn_maxpool3x3s2(C.byref(ti), C.byref(to), None))
    torch.cuda.synchronize()
    assert torch.equal(y.permute(0, 3, 1, 2).cpu(), F.max_pool2d(x, 3, 2, 1))


@pytest.mark.parametrize("h,w", [(13, 21), (16, 32), (6, 6), (97, 193), (4, 4), (7, 5), (23, 30)])
def test_psp_pool_matches_adaptive_avg_pool(env, h, w):
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(h * w)
    x = torch.randn(2, 128, h, w, generator=g)
    xd = nhwc(x)
    out = torch.empty(2, 50, 128, device=dev)
    nbytes = int(lib.tdn_psp_pool_workspace_bytes(2, h, 128))
    ws = torch.empty(nbytes // 4, device=dev)
    ti, to = View(xd.view(-1), 2, h, w, 128).ct(), View(out.view(-1), 2, 1, 50, 128).ct()
    cabi.check(lib.tdn_psp_pool(C.byref(ti), C.byref(to), ws.data_ptr(), nbytes, None))
    assert lib.tdn_psp_pool(C.byref(ti), C.byref(to), ws.data_ptr(), nbytes - 4, None) == -5
    torch.cuda.synchronize()
    off = 0
    for bins in (1, 2, 3, 6):
        ref = F.adaptive_avg_pool2d(x, bins).permute(0, 2, 3, 1).reshape(2, bins * bins, 128)
        assert max_abs(out[:, off:off + bins * bins].cpu(), ref) < 2e-6
        off += bins * bins


@pytest.mark.parametrize("h,w,c", [(128, 256, 512), (97, 193, 512), (16, 32, 256), (13, 21, 64), (33, 40, 72)])
def test_psp_pool_split16_channel_slice_view(env, h, w, c):
    """The vectorised row pass (8 channels per thread, 16 x-segments) on a SPLIT16 map that is a channel range of a wider
    buffer -- the c4 view of the pyramid fold -- against adaptive_avg_pool2d of the merged values."""
    lib, cabi, View, dev = env
    g = torch.Generator(device="cuda").manual_seed(h * w + c)
    big = View.alloc(1, h, w, c + 128, dev, split=True)
    x = torch.randn(1, h, w, c + 128, generator=g, device="cuda") * 2 + 0.5
    big.base.copy_(x.reshape(-1).half()); big.lo.copy_((x.reshape(-1) - big.base.float()).half())
    v = big.channels(64, 64 + c)
    out = torch.full((1, 50, c), float("nan"), device=dev)
    nbytes = int(lib.tdn_psp_pool_workspace_bytes(1, h, c))
    ws = torch.empty(nbytes // 4, device=dev)
    ti, to = v.ct(), View(out.view(-1), 1, 1, 50, c).ct()
    cabi.check(lib.tdn_psp_pool(C.byref(ti), C.byref(to), ws.data_ptr(), nbytes, None))
    torch.cuda.synchronize()
    ref_in = v.torch().permute(0, 3, 1, 2).double().cpu()
    off = 0
    for bins in (1, 2, 3, 6):
        ref = F.adaptive_avg_pool2d(ref_in, bins).permute(0, 2, 3, 1).reshape(1, bins * bins, c)
        assert max_abs(out[:, off:off + bins * bins].cpu(), ref) < 3e-6
        off += bins * bins


@pytest.mark.parametrize("hs,ws", [(1, 1), (2, 2), (3, 3), (6, 6)])
def test_bilinear_align_corners_into_channel_slice(env, hs, ws):
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(hs)
    x = torch.randn(2, 16, hs, ws, generator=g)
    xd = nhwc(x)
    big = torch.full((2, 13, 21, 40), -7.0, device=dev)
    ti = View(xd.view(-1), 2, hs, ws, 16).ct()
    to = View(big.view(-1), 2, 13, 21, 40).channels(8, 24).ct()
    cabi.check(lib.tdn_bilinear_nhwc(C.byref(ti), C.byref(to), None))
    torch.cuda.synchronize()
    ref = F.interpolate(x, (13, 21), mode="bilinear", align_corners=True)
    assert max_abs(big[..., 8:24].permute(0, 3, 1, 2).cpu(), ref) < 2e-6
    assert float((big[..., :8] + 7).abs().max()) == 0 and float((big[..., 24:] + 7).abs().max()) == 0


def test_layernorm_hw(env):
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 64, 13, 21, generator=g) * 3 + 1.5
    gamma, beta = torch.rand(13, 21, generator=g) + 0.5, torch.randn(13, 21, generator=g)
    xd = nhwc(x)
    mean, rstd = torch.empty(128, device=dev), torch.empty(128, device=dev)
    nbytes = int(lib.tdn_layernorm_hw_workspace_bytes(2, 13, 21, 64))
    ws = torch.empty(nbytes // 4, device=dev)
    tx = View(xd.view(-1), 2, 13, 21, 64).ct()
    cabi.check(lib.tdn_layernorm_hw_stats(C.byref(tx), mean.data_ptr(), rstd.data_ptr(), C.c_float(1e-5),
                                          ws.data_ptr(), nbytes, None))
    y = torch.empty(2, 13, 21, 64, device=dev)
    gd, bd = gamma.cuda().view(-1), beta.cuda().view(-1)
    ty = View(y.view(-1), 2, 13, 21, 64).ct()
    cabi.check(lib.tdn_layernorm_hw_apply(C.byref(tx), mean.data_ptr(), rstd.data_ptr(), gd.data_ptr(),
                                          bd.data_ptr(), C.byref(ty), None))
    torch.cuda.synchronize()
    ref = F.layer_norm(x, (13, 21), gamma, beta, 1e-5)
    assert max_abs(y.permute(0, 3, 1, 2).cpu(), ref) < 5e-6


def test_layernorm_hw_split16_wide(env):
    """SPLIT16 map with 512 channels (the 8-channel-per-thread partial-sum kernel) on a ragged number of pixels."""
    lib, cabi, View, dev = env
    g = torch.Generator(device="cuda").manual_seed(5)
    n, h, w, c = 2, 37, 53, 512
    x = torch.randn(n, h, w, c, generator=g, device="cuda") * 3 + 1.5
    xv = View.alloc(n, h, w, c, dev, split=True)
    xv.base.copy_(x.reshape(-1).half()); xv.lo.copy_((x.reshape(-1) - xv.base.float()).half())
    xm = xv.torch()
    mean, rstd = torch.empty(n * c, device=dev), torch.empty(n * c, device=dev)
    nbytes = int(lib.tdn_layernorm_hw_workspace_bytes(n, h, w, c))
    ws = torch.empty(nbytes // 4, device=dev)
    tx = xv.ct()
    cabi.check(lib.tdn_layernorm_hw_stats(C.byref(tx), mean.data_ptr(), rstd.data_ptr(), C.c_float(1e-5),
                                          ws.data_ptr(), nbytes, None))
    torch.cuda.synchronize()
    xd = xm.double()
    mu, var = xd.mean(dim=(1, 2)), xd.var(dim=(1, 2), unbiased=False)
    assert max_abs(mean.cpu(), mu.reshape(-1).cpu()) < 1e-6
    assert max_abs(rstd.cpu(), (1.0 / torch.sqrt(var + 1e-5)).reshape(-1).cpu()) < 1e-6


@pytest.mark.parametrize("h,w,H,W", [(13, 21, 97, 161), (16, 32, 128, 256), (8, 8, 64, 64), (128, 256, 1024, 2048),
                                     (64, 128, 512, 1024), (45, 60, 360, 480)])
def test_upsample_logits_nchw(env, h, w, H, W):
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(H)
    x = torch.randn(2, 19, h, w, generator=g)
    xd = nhwc(x)
    out = torch.empty(2, 19, H, W, device=dev)
    t = View(xd.view(-1), 2, h, w, 19).ct()
    cabi.check(lib.tdn_upsample_logits(C.byref(t), out.data_ptr(), H, W, None))
    torch.cuda.synchronize()
    ref = F.interpolate(x, (H, W), mode="bilinear", align_corners=True)
    assert max_abs(out.cpu(), ref) < 2e-6


@pytest.mark.parametrize("h,w,H,W", [(13, 21, 97, 161), (128, 256, 1024, 2048), (97, 193, 769, 1537), (24, 40, 192, 320)])
def test_upsample_argmax_equals_argmax_of_upsampled_logits(env, h, w, H, W):
    """tdn_upsample_argmax (Testing/test.py:61 fused into the final interpolation) == arg-max over the classes of what
    tdn_upsample_logits writes, bit for bit: both kernel pairs (per-thread loads / shared-memory rows) share one bilerp."""
    lib, cabi, View, dev = env
    g = torch.Generator(device="cuda").manual_seed(H + w)
    pad = torch.randn(1, h, w, 24, generator=g, device="cuda")          # 19 classes at a 24-float pixel pitch
    t = View(pad.view(-1), 1, h, w, 24).narrow_c(19).ct()
    out = torch.empty(1, 19, H, W, device=dev)
    lab = torch.full((1, H, W), 255, dtype=torch.uint8, device=dev)
    cabi.check(lib.tdn_upsample_logits(C.byref(t), out.data_ptr(), H, W, None))
    cabi.check(lib.tdn_upsample_argmax(C.byref(t), lab.data_ptr(), H, W, None))
    torch.cuda.synchronize()
    ref = F.interpolate(pad[..., :19].permute(0, 3, 1, 2).cpu(), (H, W), mode="bilinear", align_corners=True)
    assert max_abs(out.cpu(), ref) < 2e-6
    assert torch.equal(lab.long(), out.max(1)[1])


def test_errors_are_reported_not_swallowed(env):
    lib, cabi, View, dev = env
    x = torch.zeros(1, 5, 5, 6, device=dev)
    d = cabi.Conv2dDesc()
    d.in_ = View(x.view(-1), 1, 5, 5, 6).ct()          # cin = 6: not a multiple of 4
    d.out = View(x.view(-1), 1, 5, 5, 6).ct()
    d.weight, d.cout, d.kh, d.kw, d.stride, d.dilation, d.batch = x.data_ptr(), 6, 1, 1, 1, 1, 1
    assert lib.tdn_conv2d(C.byref(d), None) == -2
    with pytest.raises(RuntimeError, match="multiple of 4"):
        cabi.check(lib.tdn_conv2d(C.byref(d), None), "conv2d")


# ------------------------------------------------------------------------------------------------
# SPLIT16 (hi/lo fp16 planes) views through the same operators
# ------------------------------------------------------------------------------------------------
def split_planes(x):
    hi = x.half()
    return hi.contiguous(), (x - hi.float()).half().contiguous()


def test_split16_roundtrip_precision(env):
    """copy F32 -> SPLIT16 -> F32 keeps >= 22 significant bits (and is exact for small magnitudes)."""
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(21)
    x = (torch.randn(2, 9, 11, 64, generator=g) * 10).cuda()
    s = View.alloc(2, 9, 11, 64, dev, split=True)
    y = torch.empty_like(x)
    tx, ty, ts = View(x.view(-1), 2, 9, 11, 64).ct(), View(y.view(-1), 2, 9, 11, 64).ct(), s.ct()
    cabi.check(lib.tdn_split16(C.byref(tx), C.byref(ts), None))
    cabi.check(lib.tdn_merge16(C.byref(ts), C.byref(ty), None))
    torch.cuda.synchronize()
    rel = ((y - x).abs() / x.abs().clamp_min(0.25)).max().item()
    assert rel < 2.0 ** -21, rel
    hi, lo = split_planes(x)
    assert torch.equal(s.base.view_as(x), hi) and torch.equal(s.lo.view_as(x), lo)


def test_simt_conv_with_split16_views(env):
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(22)
    x = torch.randn(1, 64, 15, 22, generator=g)
    wt = torch.randn(128, 64, 3, 3, generator=g) / 24
    r = torch.randn(1, 128, 8, 11, generator=g)
    xs = View.alloc(1, 15, 22, 64, dev, split=True)
    hi, lo = split_planes(nhwc(x))
    xs.base.copy_(hi.view(-1)); xs.lo.copy_(lo.view(-1))
    rs = View.alloc(1, 8, 11, 128, dev, split=True)
    hi, lo = split_planes(nhwc(r))
    rs.base.copy_(hi.view(-1)); rs.lo.copy_(lo.view(-1))
    out = View.alloc(1, 8, 11, 128, dev, split=True)
    wk = wt.permute(0, 2, 3, 1).contiguous().cuda()
    d = cabi.Conv2dDesc()
    d.in_, d.out, d.residual = xs.ct(), out.ct(), rs.ct()
    d.weight, d.cout, d.kh, d.kw, d.stride, d.pad, d.dilation, d.act, d.batch = wk.data_ptr(), 128, 3, 3, 2, 1, 1, 1, 1
    cabi.check(lib.tdn_conv2d(C.byref(d), None))
    torch.cuda.synchronize()
    xr, rr = xs.torch().permute(0, 3, 1, 2).cpu(), rs.torch().permute(0, 3, 1, 2).cpu()
    ref = F.relu(F.conv2d(xr, wt, None, 2, 1) + rr)
    assert max_abs(out.torch().permute(0, 3, 1, 2).cpu(), ref) < 2e-5


def test_softmax_rows_split16(env):
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(23)
    s = torch.randn(37, 128, generator=g).cuda() * 20
    p_hi = torch.full((37, 128), 7.0, dtype=torch.half, device=dev)
    p_lo = torch.full((37, 128), 7.0, dtype=torch.half, device=dev)
    cabi.check(lib.tdn_softmax_rows_split16(s.data_ptr(), 37, 100, 128, C.c_float(0.125), p_hi.data_ptr(),
                                            p_lo.data_ptr(), 128, C.c_float(1024.0), None))
    torch.cuda.synchronize()
    ref = torch.softmax(s[:, :100].cpu() / 8.0, dim=1)
    got = (p_hi.float() + p_lo.float()).cpu() / 1024.0
    assert max_abs(got[:, :100], ref) < 2e-7
    assert float(got[:, 100:].abs().max()) == 0.0


TC_CASES = [
    # n, h, w, cin, cout, k, dil
    (1, 1, 256, 64, 128, 1, 1),      # plain GEMM, one K block
    (1, 1, 1000, 512, 256, 1, 1),    # ragged M, 8 K blocks (ring wraps), two N tiles
    (1, 1, 384, 128, 64, 1, 1),      # BLOCK_N = 64 variant
    (2, 24, 40, 64, 128, 3, 2),      # 3x3 dilated, 8x16 pixel tiles, batch 2
    (1, 97, 193, 128, 96, 3, 8),     # ragged map (reference-native 97x193), cout not a multiple of 32... 96
    (1, 23, 30, 64, 40, 3, 4),       # cout = 40: partial 32-channel chunk in the epilogue
]


@pytest.mark.parametrize("n,h,w,cin,cout,k,dil", TC_CASES)
def test_tc_conv_exact_mode(env, n, h, w, cin, cout, k, dil):
    """tcgen05 SPLIT16 conv against an fp64 convolution of the same fp32 inputs: fp32-level error."""
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(n * h + w + cin + cout)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    xs = View.alloc(n, h, w, cin, dev, split=True)
    hi, lo = split_planes(nhwc(x))
    xs.base.copy_(hi.view(-1)); xs.lo.copy_(lo.view(-1))
    wh, wl = split_planes(wt.permute(0, 2, 3, 1).reshape(cout, -1).cuda())
    out = View.alloc(n, h, w, cout, dev)
    d = cabi.TcConvDesc()
    d.in_, d.out = xs.ct(), out.ct()
    d.weight_hi, d.weight_lo, d.weight_ld = wh.data_ptr(), wl.data_ptr(), k * k * cin
    d.cout, d.kh, d.kw, d.dilation = cout, k, k, dil
    cabi.check(lib.tdn_conv2d_tc(C.byref(d), None), "conv2d_tc")
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), wt.double(), None, 1, dil * (k - 1) // 2, dil)
    got = out.torch().permute(0, 3, 1, 2).cpu()
    err = max_abs(got, ref)
    assert err < 2e-6 * max(1.0, float(ref.abs().max())), err


PAIR_CASES = [
    # n, h, w, cin, cout, k, dil, stride      (cout % 128 == 0: N = 256 pair tiles when cout % 256 == 0, else 128)
    (1, 1, 512, 64, 256, 1, 1, 1),       # plain GEMM: 4 M tiles = 2 pair tiles, one K block
    (1, 1, 1000, 512, 256, 1, 1, 1),     # ragged M (8 M tiles), 8 K blocks: the stage ring and the chunk ring wrap
    (1, 1, 640, 128, 512, 1, 1, 1),      # 5 M tiles (odd: the last pair has an empty half), two N tiles
    (2, 24, 40, 64, 256, 3, 2, 1),       # 3x3 dilated, batch 2, zero padding through TMA
    (1, 97, 193, 128, 256, 3, 4, 1),     # reference-native ragged map, 18 K blocks
    (1, 33, 47, 64, 128, 3, 1, 1),       # N = 128 pair tiles
    (1, 40, 56, 64, 256, 3, 1, 2),       # stride 2 (TMA element strides)
    (3, 16, 8, 192, 384, 1, 1, 1),       # cout = 384: N = 128 tiles x 3, 3 K blocks, batch 3 (3 M tiles)
]


@pytest.mark.parametrize("n,h,w,cin,cout,k,dil,stride", PAIR_CASES)
def test_tc_conv_pair_variant_is_bit_identical(env, n, h, w, cin, cout, k, dil, stride):
    """tc_conv_pair_kernel (2-CTA clusters, tcgen05 cta_group::2) against the single-CTA kernel: the same products
    in the same order, so the SPLIT16 planes must be bit-identical -- with BN / residual / ReLU epilogue -- and both
    within fp32 rounding of an fp64 convolution."""
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(n * h + w + cin + cout + k)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    sc, bi = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.2
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    res = torch.randn(n, cout, ho, wo, generator=g)
    xs = View.alloc(n, h, w, cin, dev, split=True)
    hi, lo = split_planes(nhwc(x))
    xs.base.copy_(hi.view(-1)); xs.lo.copy_(lo.view(-1))
    rs = View.alloc(n, ho, wo, cout, dev, split=True)
    hi, lo = split_planes(nhwc(res))
    rs.base.copy_(hi.view(-1)); rs.lo.copy_(lo.view(-1))
    wh, wl = split_planes(wt.permute(0, 2, 3, 1).reshape(cout, -1).cuda())
    scd, bid = sc.cuda(), bi.cuda()
    outs = {}
    for variant in (cabi.TC_BASE, cabi.TC_PAIR, cabi.TC_BASE_TS):
        out = View.alloc(n, ho, wo, cout, dev, split=True)
        out.base.fill_(float("nan")); out.lo.fill_(float("nan"))
        d = cabi.TcConvDesc()
        d.in_, d.out, d.residual = xs.ct(), out.ct(), rs.ct()
        d.weight_hi, d.weight_lo, d.weight_ld = wh.data_ptr(), wl.data_ptr(), k * k * cin
        d.scale, d.bias = scd.data_ptr(), bid.data_ptr()
        d.cout, d.kh, d.kw, d.dilation, d.stride, d.act = cout, k, k, dil, stride, 1
        d.variant = variant
        cabi.check(lib.tdn_conv2d_tc(C.byref(d), None), "conv2d_tc")
        torch.cuda.synchronize()
        outs[variant] = (out.base.clone(), out.lo.clone(), out.torch().permute(0, 3, 1, 2).cpu())
    base, pair = outs[cabi.TC_BASE], outs[cabi.TC_PAIR]
    xr = (rs.base.float() + rs.lo.float()).view(n, ho, wo, cout).permute(0, 3, 1, 2).cpu().double()
    ref = F.relu(F.conv2d(x.double(), wt.double(), None, stride, dil * (k - 1) // 2, dil)
                 * sc.double().view(1, -1, 1, 1) + bi.double().view(1, -1, 1, 1) + xr)
    assert max_abs(pair[2], ref) < 3e-6 * max(1.0, float(ref.abs().max()))
    assert torch.equal(base[0].view(torch.int16), pair[0].view(torch.int16))
    assert torch.equal(base[1].view(torch.int16), pair[1].view(torch.int16))
    ts = outs[cabi.TC_BASE_TS]       # the single-CTA kernel with the A tile copied to tensor memory per K block (tcgen05.cp)
    assert torch.equal(base[0].view(torch.int16), ts[0].view(torch.int16))
    assert torch.equal(base[1].view(torch.int16), ts[1].view(torch.int16))


def test_tc_conv_auto_pair_plus_tail_split_is_bit_identical(env):
    """Layer-4-like tiling (256 M tiles x 512 channels = 256 pair tiles on 74 clusters): the single-CTA kernel, the library's
    choice, the CTA-pair kernel, its tail variant (full rounds as N = 256 pair tiles + the ragged last round as N = 128 pair
    tiles in a second launch) and its quad variant (clusters of two pairs sharing the weight tile by TMA multicast) must
    agree bit for bit."""
    lib, cabi, View, dev = env
    n, h, w, cin, cout, k, dil = 1, 128, 256, 128, 512, 3, 4
    g = torch.Generator(device="cuda").manual_seed(5)
    xs = View.alloc(n, h, w, cin, dev, split=True)
    x = torch.randn(n * h * w * cin, generator=g, device="cuda")
    xs.base.copy_(x.half()); xs.lo.copy_((x - x.half().float()).half())
    wt = torch.randn(cout, k * k * cin, generator=g, device="cuda") / (cin * k * k) ** 0.5
    wh, wl = split_planes(wt)
    outs = []
    for variant in (cabi.TC_BASE, cabi.TC_AUTO, cabi.TC_PAIR, cabi.TC_PAIR_TAIL, cabi.TC_PAIR_QUAD):
        out = View.alloc(n, h, w, cout, dev, split=True)
        out.base.fill_(float("nan")); out.lo.fill_(float("nan"))
        d = cabi.TcConvDesc()
        d.in_, d.out = xs.ct(), out.ct()
        d.weight_hi, d.weight_lo, d.weight_ld = wh.data_ptr(), wl.data_ptr(), k * k * cin
        d.cout, d.kh, d.kw, d.dilation, d.variant = cout, k, k, dil, variant
        cabi.check(lib.tdn_conv2d_tc(C.byref(d), None), "conv2d_tc")
        torch.cuda.synchronize()
        outs.append((out.base.clone(), out.lo.clone()))
    assert not torch.isnan(outs[1][0].float()).any()
    for other in outs[1:]:
        assert torch.equal(outs[0][0].view(torch.int16), other[0].view(torch.int16))
        assert torch.equal(outs[0][1].view(torch.int16), other[1].view(torch.int16))


@pytest.mark.parametrize("n,h,w,cin,cout,dil", [(1, 64, 96, 64, 64, 1), (2, 45, 77, 128, 128, 2), (1, 33, 40, 64, 128, 1)])
def test_tc_conv_swizzled_halo_variant_is_bit_identical_to_halo(env, n, h, w, cin, cout, dil):
    """tc_conv_halo_sw.cu (the halo region in the 128-byte-swizzled layout, a filter tap = a row offset into it, descriptor
    base offset 0) against tc_conv_halo.cu (no-swizzle region): same products in the same order; ragged map sizes exercise
    the zero fill of the 16-pixel-wide region boxes; and both against the fp64 convolution."""
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(h * w + cin)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    xs = View.alloc(n, h, w, cin, dev, split=True)
    hi, lo = split_planes(nhwc(x))
    xs.base.copy_(hi.view(-1)); xs.lo.copy_(lo.view(-1))
    wh, wl = split_planes(wt.permute(0, 2, 3, 1).reshape(cout, -1).cuda())
    outs = {}
    for variant in (cabi.TC_HALO, cabi.TC_HALO_SW):
        out = View.alloc(n, h, w, cout, dev, split=True)
        out.base.fill_(float("nan")); out.lo.fill_(float("nan"))
        d = cabi.TcConvDesc()
        d.in_, d.out = xs.ct(), out.ct()
        d.weight_hi, d.weight_lo, d.weight_ld = wh.data_ptr(), wl.data_ptr(), 9 * cin
        d.cout, d.kh, d.kw, d.dilation, d.variant = cout, 3, 3, dil, variant
        cabi.check(lib.tdn_conv2d_tc(C.byref(d), None), "conv2d_tc")
        torch.cuda.synchronize()
        outs[variant] = (out.base.clone(), out.lo.clone(), out.torch().permute(0, 3, 1, 2).cpu())
    a, b = outs[cabi.TC_HALO], outs[cabi.TC_HALO_SW]
    assert torch.equal(a[0].view(torch.int16), b[0].view(torch.int16))
    assert torch.equal(a[1].view(torch.int16), b[1].view(torch.int16))
    ref = F.conv2d(x.double(), wt.double(), None, 1, dil, dil)
    assert max_abs(b[2], ref) < 2e-6 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("n,h,w,cin,cout,dil", [(1, 64, 96, 64, 256, 2), (2, 45, 77, 128, 256, 1), (1, 40, 56, 64, 512, 4),
                                                (1, 33, 24, 128, 256, 3)])
def test_tc_conv_pair_band_variant(env, n, h, w, cin, cout, dil):
    """tc_conv_pair_band.cu (CTA pairs, one activation band per channel block and filter row, taps = row offsets into the
    swizzled band): bit-identical to tc_conv_halo.cu where that applies (same (channel block, tap) order), within fp32
    rounding of the tap-major pair kernel and of the fp64 convolution everywhere (dilation 3 / 4, ragged maps, BN + residual +
    ReLU epilogue)."""
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(h * w + cin + dil)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    sc, bi = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.2
    res = torch.randn(n, cout, h, w, generator=g)
    xs = View.alloc(n, h, w, cin, dev, split=True)
    hi, lo = split_planes(nhwc(x))
    xs.base.copy_(hi.view(-1)); xs.lo.copy_(lo.view(-1))
    rs = View.alloc(n, h, w, cout, dev, split=True)
    hi, lo = split_planes(nhwc(res))
    rs.base.copy_(hi.view(-1)); rs.lo.copy_(lo.view(-1))
    wh, wl = split_planes(wt.permute(0, 2, 3, 1).reshape(cout, -1).cuda())
    scd, bid = sc.cuda(), bi.cuda()
    outs = {}
    for variant in (cabi.TC_PAIR, cabi.TC_PAIR_BAND) + ((cabi.TC_HALO,) if dil <= 2 else ()):
        out = View.alloc(n, h, w, cout, dev, split=True)
        out.base.fill_(float("nan")); out.lo.fill_(float("nan"))
        d = cabi.TcConvDesc()
        d.in_, d.out, d.residual = xs.ct(), out.ct(), rs.ct()
        d.weight_hi, d.weight_lo, d.weight_ld = wh.data_ptr(), wl.data_ptr(), 9 * cin
        d.scale, d.bias = scd.data_ptr(), bid.data_ptr()
        d.cout, d.kh, d.kw, d.dilation, d.act, d.variant = cout, 3, 3, dil, 1, variant
        cabi.check(lib.tdn_conv2d_tc(C.byref(d), None), "conv2d_tc")
        torch.cuda.synchronize()
        outs[variant] = (out.base.clone(), out.lo.clone(), out.torch().permute(0, 3, 1, 2).cpu())
    band = outs[cabi.TC_PAIR_BAND]
    assert not torch.isnan(band[2]).any()
    if dil <= 2:
        halo = outs[cabi.TC_HALO]
        assert torch.equal(halo[0].view(torch.int16), band[0].view(torch.int16))
        assert torch.equal(halo[1].view(torch.int16), band[1].view(torch.int16))
    xr = (rs.base.float() + rs.lo.float()).view(n, h, w, cout).permute(0, 3, 1, 2).cpu().double()
    ref = F.relu(F.conv2d(x.double(), wt.double(), None, 1, dil, dil) * sc.double().view(1, -1, 1, 1)
                 + bi.double().view(1, -1, 1, 1) + xr)
    assert max_abs(band[2], ref) < 3e-6 * max(1.0, float(ref.abs().max()))
    assert max_abs(band[2], outs[cabi.TC_PAIR][2]) < 3e-6 * max(1.0, float(ref.abs().max()))


def test_sm_clock_probe(env):
    """tdn_sm_clock_probe: %clock64 cycles per %globaltimer nanosecond on a few SMs = a plausible SM clock, and the spin lasts
    at least the requested time (the diagnostic behind bench.py's clocks.sm_mhz_on_sm)."""
    lib, cabi, View, dev = env
    blocks = 8
    out = torch.zeros(3 * blocks, dtype=torch.int64, device=dev)
    cabi.check(lib.tdn_sm_clock_probe(out.data_ptr(), blocks, 2_000_000, None), "sm_clock_probe")
    torch.cuda.synchronize()
    o = out.view(blocks, 3).cpu()
    assert bool((o[:, 1] >= 2_000_000).all())
    ghz = o[:, 0].double() / o[:, 1].double()
    assert bool((ghz > 0.3).all()) and bool((ghz < 3.0).all()), ghz
    assert lib.tdn_sm_clock_probe(None, blocks, 2_000_000, None) < 0      # null output: TDN_ERR_INVALID, nothing launched


def test_tc_conv_variant_errors(env):
    lib, cabi, View, dev = env
    xs = View.alloc(1, 8, 16, 64, dev, split=True)
    out = View.alloc(1, 8, 16, 40, dev, split=True)
    w = torch.zeros(40 * 64, dtype=torch.half, device=dev)
    d = cabi.TcConvDesc()
    d.in_, d.out = xs.ct(), out.ct()
    d.weight_hi, d.weight_lo, d.weight_ld = w.data_ptr(), w.data_ptr(), 64
    d.cout, d.kh, d.kw, d.dilation = 40, 1, 1, 1
    d.variant = cabi.TC_PAIR                    # cout % 128 != 0
    assert lib.tdn_conv2d_tc(C.byref(d), None) == -2
    d.variant = cabi.TC_HALO                    # not a 3x3 convolution
    assert lib.tdn_conv2d_tc(C.byref(d), None) == -2
    d.variant = 9
    assert lib.tdn_conv2d_tc(C.byref(d), None) == -1


def test_tc_conv_epilogue_and_batched_weights(env):
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(31)
    x = torch.randn(1, 64, 16, 32, generator=g)
    wt = torch.randn(128, 64, 3, 3, generator=g) / 24
    sc, bi = torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g)
    r = torch.randn(1, 128, 16, 32, generator=g)
    xs, rs = View.alloc(1, 16, 32, 64, dev, split=True), View.alloc(1, 16, 32, 128, dev, split=True)
    for v, t in ((xs, x), (rs, r)):
        hi, lo = split_planes(nhwc(t))
        v.base.copy_(hi.view(-1)); v.lo.copy_(lo.view(-1))
    wh, wl = split_planes(wt.permute(0, 2, 3, 1).reshape(128, -1).cuda())
    out = View.alloc(1, 16, 32, 128, dev, split=True)
    scd, bid = sc.cuda(), bi.cuda()
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    d = cabi.TcConvDesc()
    d.in_, d.out, d.residual = xs.ct(), out.ct(), rs.ct()
    d.weight_hi, d.weight_lo, d.weight_ld = wh.data_ptr(), wl.data_ptr(), 576
    d.scale, d.bias, d.cout, d.kh, d.kw, d.dilation, d.act = scd.data_ptr(), bid.data_ptr(), 128, 3, 3, 1, 1
    d.range_flag = flag.data_ptr()
    cabi.check(lib.tdn_conv2d_tc(C.byref(d), None), "conv2d_tc")
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(x.double(), wt.double(), None, 1, 1) * sc.double().view(1, -1, 1, 1)
                 + bi.double().view(1, -1, 1, 1) + rs.torch().permute(0, 3, 1, 2).cpu().double())
    assert max_abs(out.torch().permute(0, 3, 1, 2).cpu(), ref) < 2e-6 * float(ref.abs().max())
    assert int(flag.item()) == 0
    # q k^T with one key matrix per image (transformer.py:128)
    q, k = torch.randn(2, 300, 64, generator=g), torch.randn(2, 100, 64, generator=g)
    qs = View.alloc(2, 1, 300, 64, dev, split=True)
    hi, lo = split_planes(q.cuda())
    qs.base.copy_(hi.view(-1)); qs.lo.copy_(lo.view(-1))
    kh, kl = split_planes(k.cuda())
    s = View.alloc(2, 1, 300, 128, dev)
    d = cabi.TcConvDesc()
    d.in_, d.out = qs.ct(), s.narrow_c(100).ct()
    d.weight_hi, d.weight_lo, d.weight_ld, d.weight_batched, d.weight_batch_stride = kh.data_ptr(), kl.data_ptr(), 64, 1, 6400
    d.cout, d.kh, d.kw, d.dilation = 100, 1, 1, 1
    cabi.check(lib.tdn_conv2d_tc(C.byref(d), None), "conv2d_tc")
    torch.cuda.synchronize()
    ref = torch.bmm(q.double(), k.double().transpose(1, 2))
    assert max_abs(s.torch()[:, 0, :, :100].cpu(), ref) < 2e-6 * float(ref.abs().max())


@pytest.mark.parametrize("n,pq,pk,dv", [(2, 300, 100, 128), (1, 2048, 2048, 512), (1, 1000, 690, 256),
                                        # 200 / 188 items of 256 channels on 148 SMs: the ragged last round runs as a
                                        # second launch of 128-channel items over the remaining query tiles
                                        (1, 12777, 200, 512), (2, 6000, 130, 512)])
def test_fused_attention_tc(env, n, pq, pk, dv):
    """tdn_attention_tc vs fp64 softmax(q k^T / 8) v + residual (transformer.py:126-139)."""
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(pq + pk)
    q, k = torch.randn(n, pq, 64, generator=g) * 1.3, torch.randn(n, pk, 64, generator=g) * 1.4
    v, r = torch.randn(n, pk, dv, generator=g) * 3, torch.randn(n, pq, dv, generator=g)
    pkp = (pk + 63) // 64 * 64
    vt = torch.zeros(n, dv, pkp)
    vt[:, :, :pk] = v.transpose(1, 2)
    planes = {}
    for name, t in (("q", q), ("k", k), ("vt", vt), ("r", r)):
        planes[name] = split_planes(t.cuda())
    out = torch.empty(n, pq, dv, device=dev)
    d = cabi.AttentionDesc()
    d.q_hi, d.q_lo, d.q_ld, d.q_batch_stride = planes["q"][0].data_ptr(), planes["q"][1].data_ptr(), 64, pq * 64
    d.k_hi, d.k_lo, d.k_ld, d.k_batch_stride = planes["k"][0].data_ptr(), planes["k"][1].data_ptr(), 64, pk * 64
    d.vt_hi, d.vt_lo, d.vt_ld, d.vt_batch_stride = planes["vt"][0].data_ptr(), planes["vt"][1].data_ptr(), pkp, dv * pkp
    d.out = cabi.Tensor(out.data_ptr(), None, 0, n, 1, pq, dv, pq * dv, pq * dv, dv)
    d.residual = cabi.Tensor(planes["r"][0].data_ptr(), planes["r"][1].data_ptr(), 1, n, 1, pq, dv, pq * dv, pq * dv, dv)
    d.n, d.pq, d.pk, d.d_k, d.d_v = n, pq, pk, 64, dv
    cabi.check(lib.tdn_attention_tc(C.byref(d), None), "attention_tc")
    torch.cuda.synchronize()
    a = torch.softmax(torch.bmm(q.double(), k.double().transpose(1, 2)) / 8.0, dim=2)
    ref = torch.bmm(a, v.double()) + r.double()
    # the O accumulator chains up to 12 * ceil(pk / 64) tensor-core adds (truncating), hence the looser bound
    assert max_abs(out.cpu(), ref) < 1.5e-5 * float(ref.abs().max())
    assert float((out.cpu().double() - ref).norm() / ref.norm()) < 5e-6


def _attention_case(cabi, lib, dev, n, pq, pk, dv, out_fmt="f32", res_fmt="split", seed=None):
    """One tdn_attention_tc call on seeded operands; returns (out fp32 [n,pq,dv] on the device, q, k, v, r on the host)."""
    g = torch.Generator().manual_seed(pq + pk if seed is None else seed)
    q, k = torch.randn(n, pq, 64, generator=g) * 1.3, torch.randn(n, pk, 64, generator=g) * 1.4
    v, r = torch.randn(n, pk, dv, generator=g) * 3, torch.randn(n, pq, dv, generator=g)
    pkp = (pk + 63) // 64 * 64
    vt = torch.zeros(n, dv, pkp)
    vt[:, :, :pk] = v.transpose(1, 2)
    pl = {name: split_planes(t.cuda()) for name, t in (("q", q), ("k", k), ("vt", vt), ("r", r))}
    rf = r.cuda().contiguous()
    of = torch.full((n, pq, dv), float("nan"), device=dev)
    oh = torch.full((n, pq, dv), float("nan"), device=dev, dtype=torch.half)
    ol = torch.full((n, pq, dv), float("nan"), device=dev, dtype=torch.half)
    d = cabi.AttentionDesc()
    d.q_hi, d.q_lo, d.q_ld, d.q_batch_stride = pl["q"][0].data_ptr(), pl["q"][1].data_ptr(), 64, pq * 64
    d.k_hi, d.k_lo, d.k_ld, d.k_batch_stride = pl["k"][0].data_ptr(), pl["k"][1].data_ptr(), 64, pk * 64
    d.vt_hi, d.vt_lo, d.vt_ld, d.vt_batch_stride = pl["vt"][0].data_ptr(), pl["vt"][1].data_ptr(), pkp, dv * pkp
    if out_fmt == "f32":
        d.out = cabi.Tensor(of.data_ptr(), None, 0, n, 1, pq, dv, pq * dv, pq * dv, dv)
    else:
        d.out = cabi.Tensor(oh.data_ptr(), ol.data_ptr(), 1, n, 1, pq, dv, pq * dv, pq * dv, dv)
    if res_fmt == "split":
        d.residual = cabi.Tensor(pl["r"][0].data_ptr(), pl["r"][1].data_ptr(), 1, n, 1, pq, dv, pq * dv, pq * dv, dv)
    elif res_fmt == "f32":
        d.residual = cabi.Tensor(rf.data_ptr(), None, 0, n, 1, pq, dv, pq * dv, pq * dv, dv)
    d.n, d.pq, d.pk, d.d_k, d.d_v = n, pq, pk, 64, dv
    cabi.check(lib.tdn_attention_tc(C.byref(d), None), "attention_tc")
    torch.cuda.synchronize()
    return (of if out_fmt == "f32" else oh.float() + ol.float()), q, k, v, r


@pytest.fixture
def attn_family():
    """Restores TDNET_ATTN_TS (the library reads it on every call) after a test that flips kernel families."""
    import os
    old = os.environ.get("TDNET_ATTN_TS")
    yield lambda ts: os.environ.__setitem__("TDNET_ATTN_TS", str(int(ts)))   # 0 SS, 1 TS, 3 TS with Q in tensor memory, 4 TS with 128-key S tiles
    if old is None:
        os.environ.pop("TDNET_ATTN_TS", None)
    else:
        os.environ["TDNET_ATTN_TS"] = old


@pytest.mark.parametrize("out_fmt", ["f32", "split"])
@pytest.mark.parametrize("res_fmt", ["split", "f32", "none"])
@pytest.mark.parametrize("n,pq,pk,dv", [(2, 1000, 690, 512), (1, 300, 100, 256), (1, 20000, 200, 512), (1, 37888, 128, 512)])
def test_attention_kernel_families_bit_identical(env, attn_family, n, pq, pk, dv, out_fmt, res_fmt):
    """The tensor-memory-operand kernels (tc_attn_ts.cu, default) and the shared-memory-operand kernels (tc_attn.cu)
    issue the same products in the same order per output element: equal bit for bit in every out / residual format,
    including the split into 256- and 128-channel launches (20000 queries: 314 items on 148 SMs) and ragged tiles.  The last
    shape (four items per CTA of only two key tiles each) lets the S issuer run a whole item ahead of P.V': its o_full
    parity waits must not skip a phase (a bug of the first tc_attn_s128.cu, latent in tc_attn_ts.cu)."""
    lib, cabi, View, dev = env
    got = {}
    for ts in (0, 1, 3, 4):
        attn_family(ts)
        got[ts] = _attention_case(cabi, lib, dev, n, pq, pk, dv, out_fmt, res_fmt)[0]
    assert not torch.isnan(got[1]).any()
    assert torch.equal(got[0], got[1])
    assert torch.equal(got[0], got[3])         # Q as a tensor-memory operand: same products, same order
    assert torch.equal(got[0], got[4])         # 128-key S MMAs (tc_attn_s128.cu): same products, same order


@pytest.mark.parametrize("n,pq,pk,dv", [(1, 32768, 2048, 512), (1, 32768, 1225, 512), (1, 4096, 2048, 1024)])
def test_fused_attention_tc_big_hop(env, attn_family, n, pq, pk, dv):
    """The big hop of td4-psp18 at 1024x2048 (32768 queries x 2048 keys, d_v 512), the 769x1537 key count (P' = 1225)
    and a ResNet-50 d_v: both kernel families bit-identical, and rows sampled across the map against the fp64
    softmax(q k^T / 8) v + residual of transformer.py:126-139 (the full fp64 matrix would not fit the test budget)."""
    lib, cabi, View, dev = env
    got = {}
    for ts in (0, 1, 3, 4):
        attn_family(ts)
        got[ts], q, k, v, r = _attention_case(cabi, lib, dev, n, pq, pk, dv, "split", "split")
    assert torch.equal(got[0], got[1])
    assert torch.equal(got[0], got[3])
    assert torch.equal(got[0], got[4])
    rows = torch.arange(0, pq, 61)
    a = torch.softmax(q[:, rows].double() @ k.double().transpose(1, 2) / 8.0, dim=2)
    ref = a @ v.double() + r[:, rows].double()
    out = got[1][:, rows.to(dev)].cpu().double()
    assert max_abs(out, ref) < 1.5e-5 * float(ref.abs().max())
    assert float((out - ref).norm() / ref.norm()) < 5e-6


@pytest.mark.parametrize("n,h,w,cin,cout,k", [(1, 64, 96, 64, 128, 3), (2, 25, 41, 64, 128, 3), (1, 193, 385, 64, 64, 3),
                                               (1, 25, 41, 128, 256, 1)])
def test_tc_conv_stride2(env, n, h, w, cin, cout, k):
    """Stride-2 'same'-padded conv on the tensor cores (layer2.0.conv1 / downsample, resnet.py:170-178):
    TMA element strides pick every second pixel; odd map sizes exercise the zero fill."""
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(h * w + cout)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    xs = View.alloc(n, h, w, cin, dev, split=True)
    hi, lo = split_planes(nhwc(x))
    xs.base.copy_(hi.view(-1)); xs.lo.copy_(lo.view(-1))
    wh, wl = split_planes(wt.permute(0, 2, 3, 1).reshape(cout, -1).cuda())
    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    out = View.alloc(n, oh, ow, cout, dev)
    d = cabi.TcConvDesc()
    d.in_, d.out = (xs.subsample(2).ct() if k == 1 else xs.ct()), out.ct()
    d.weight_hi, d.weight_lo, d.weight_ld = wh.data_ptr(), wl.data_ptr(), k * k * cin
    d.cout, d.kh, d.kw, d.dilation, d.stride = cout, k, k, 1, (0 if k == 1 else 2)
    cabi.check(lib.tdn_conv2d_tc(C.byref(d), None), "conv2d_tc")
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), wt.double(), None, 2, (k - 1) // 2, 1)
    got = out.torch().permute(0, 3, 1, 2).cpu()
    assert got.shape == ref.shape
    assert max_abs(got, ref) < 2e-6 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("h,w,split", [(64, 96, False), (97, 161, True), (130, 70, True)])
def test_fused_stem_conv_bn_relu_maxpool(env, h, w, split):
    """tdn_stem_conv_pool vs conv2d(7,2,3) -> BN(eval) -> ReLU -> max_pool2d(3,2,1) (resnet.py:205-208)."""
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(h + w)
    img = torch.randn(2, 3, h, w, generator=g)
    wt = torch.randn(64, 3, 7, 7, generator=g) / 12
    sc, bi = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.3
    ref = F.max_pool2d(F.relu(F.conv2d(img, wt, None, 2, 3) * sc.view(1, -1, 1, 1) + bi.view(1, -1, 1, 1)), 3, 2, 1)
    hp, wp = ref.shape[2:]
    out = View.alloc(2, hp, wp, 64, dev, split=split)
    imgd, wk = img.cuda(), wt.permute(1, 2, 3, 0).reshape(147, 64).contiguous().cuda()
    scd, bid = sc.cuda(), bi.cuda()
    t = out.ct()
    cabi.check(lib.tdn_stem_conv_pool(imgd.data_ptr(), 2, h, w, wk.data_ptr(), scd.data_ptr(), bid.data_ptr(),
                                      C.byref(t), None), "stem")
    torch.cuda.synchronize()
    assert max_abs(out.torch().permute(0, 3, 1, 2).cpu(), ref) < 1e-5


@pytest.mark.parametrize("h,w,split,u8", [(64, 96, True, False), (97, 161, False, False), (130, 70, True, True),
                                          (257, 530, True, False), (9, 9, True, False)])
def test_tc_stem_conv_bn_relu_maxpool(env, h, w, split, u8):
    """tdn_stem_conv_pool_tc (tcgen05, exact mode) vs conv2d(7,2,3) -> BN(eval) -> ReLU -> max_pool2d(3,2,1) in fp64
    (resnet.py:205-208); several strips / bands, ragged edges, and the uint8 HWC ingest (dataloader.py:66-71)."""
    from tdnet_b200.engine import pack_stem_tc
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(h * 7 + w)
    wt = torch.randn(64, 3, 7, 7, generator=g) / 12
    sc, bi = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.3
    if u8:
        frame = torch.randint(0, 256, (2, h, w, 3), generator=g, dtype=torch.uint8)
        mean = torch.tensor([.485, .456, .406], dtype=torch.float64)
        std = torch.tensor([.229, .224, .225], dtype=torch.float64)
        img = ((frame.double() / 255.0 - mean) / std).float().permute(0, 3, 1, 2).contiguous()
        lut = ((torch.arange(256, dtype=torch.float64)[None] / 255.0 - mean[:, None]) / std[:, None]).float().cuda()
    else:
        img = torch.randn(2, 3, h, w, generator=g)
    ref = F.max_pool2d(F.relu(F.conv2d(img.double(), wt.double(), None, 2, 3) * sc.double().view(1, -1, 1, 1)
                              + bi.double().view(1, -1, 1, 1)), 3, 2, 1)
    hp, wp = ref.shape[2:]
    out = View.alloc(2, hp, wp, 64, dev, split=split)
    wk, inv = pack_stem_tc(wt.cuda())
    scd, bid = (sc.cuda() * inv).contiguous(), bi.cuda()
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    t = out.ct()
    if u8:
        fd = frame.cuda()
        rc = lib.tdn_stem_conv_pool_tc(None, fd.data_ptr(), lut.data_ptr(), 2, h, w, wk.data_ptr(), scd.data_ptr(),
                                       bid.data_ptr(), C.byref(t), flag.data_ptr(), None)
    else:
        imgd = img.cuda()
        rc = lib.tdn_stem_conv_pool_tc(imgd.data_ptr(), None, None, 2, h, w, wk.data_ptr(), scd.data_ptr(),
                                       bid.data_ptr(), C.byref(t), flag.data_ptr(), None)
    cabi.check(rc, "stem_tc")
    torch.cuda.synchronize()
    got = out.torch().permute(0, 3, 1, 2).cpu()
    assert got.shape == ref.shape
    assert max_abs(got, ref) < 3e-6 * max(1.0, float(ref.abs().max()))
    assert int(flag.item()) == 0


def test_tc_stem_rejects_bad_arguments(env):
    lib, cabi, View, dev = env
    out = View.alloc(1, 16, 24, 64, dev, split=True)
    t = out.ct()
    x = torch.zeros(16, device=dev)
    assert lib.tdn_stem_conv_pool_tc(None, None, None, 1, 64, 96, x.data_ptr(), x.data_ptr(), x.data_ptr(),
                                     C.byref(t), None, None) == -1
    assert lib.tdn_stem_conv_pool_tc(x.data_ptr(), None, None, 1, 64, 100, x.data_ptr(), x.data_ptr(), x.data_ptr(),
                                     C.byref(t), None, None) == -1


@pytest.mark.parametrize("h,w,split", [(13, 21, False), (16, 32, True)])
def test_psp_concat_matches_interpolate_and_cat(env, h, w, split):
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(h)
    x = torch.randn(2, 32, h, w, generator=g)
    feats = [torch.randn(2, 8, b, b, generator=g) for b in (1, 2, 3, 6)]
    xd = View.alloc(2, h, w, 32, dev, split=split)
    src = nhwc(x)
    if split:
        hi, lo = split_planes(src)
        xd.base.copy_(hi.view(-1)); xd.lo.copy_(lo.view(-1))
    else:
        xd.base.copy_(src.view(-1))
    smalls = [nhwc(f) for f in feats]
    ptrs = (C.c_void_p * 4)(*[t.data_ptr() for t in smalls])
    z = View.alloc(2, h, w, 64, dev, split=split)
    tx, tz = xd.ct(), z.ct()
    cabi.check(lib.tdn_psp_concat(C.byref(tx), ptrs, 8, C.byref(tz), None), "psp_concat")
    torch.cuda.synchronize()
    ref = torch.cat([xd.torch().permute(0, 3, 1, 2).cpu()] +
                    [F.interpolate(f, (h, w), mode="bilinear", align_corners=True) for f in feats], 1)
    assert max_abs(z.torch().permute(0, 3, 1, 2).cpu(), ref) < 2e-6


def test_psp_branch_convs_one_launch(env):
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(77)
    c4, eighth, n = 512, 64, 2
    pooled = torch.randn(n, 50, c4, generator=g)
    ws = [torch.randn(eighth, c4, generator=g) / 22 for _ in range(4)]
    scs = [torch.rand(eighth, generator=g) + 0.5 for _ in range(4)]
    bis = [torch.randn(eighth, generator=g) * 0.2 for _ in range(4)]
    pd = pooled.cuda()
    wd, sd_, bd = [w.cuda() for w in ws], [s.cuda() for s in scs], [b.cuda() for b in bis]
    outs = [torch.empty(n, b * b, eighth, device=dev) for b in (1, 2, 3, 6)]
    arr = lambda xs: (C.c_void_p * 4)(*[t.data_ptr() for t in xs])  # noqa: E731
    t = View(pd.view(-1), n, 1, 50, c4).ct()
    cabi.check(lib.tdn_psp_branch_convs(C.byref(t), arr(wd), arr(sd_), arr(bd), eighth, arr(outs), None), "psp_branch")
    torch.cuda.synchronize()
    off = 0
    for i, b in enumerate((1, 2, 3, 6)):
        ref = F.relu(pooled[:, off:off + b * b] @ ws[i].t() * scs[i] + bis[i])
        assert max_abs(outs[i].cpu(), ref) < 5e-6
        off += b * b


@pytest.mark.parametrize("n,c4,couts", [(1, 512, (512, 64, 64)), (2, 2048, (256, 64))])
def test_psp_branch_project_writes_projected_features(env, n, c4, couts):
    """tdn_psp_branch_project: the branch maps of tdn_psp_branch_convs plus, per projection, dst[i][o][bin] =
    sum_c w[lv * eighth + c][o] * b_lv[i][bin][c] as SPLIT16 in the dynamic columns of a wider weight matrix whose other
    columns must stay untouched."""
    lib, cabi, View, dev = env
    g = torch.Generator().manual_seed(c4 + n)
    eighth = c4 // 8
    pooled = torch.randn(n, 50, c4, generator=g)
    ws = [torch.randn(eighth, c4, generator=g) / c4 ** 0.5 for _ in range(4)]
    scs = [torch.rand(eighth, generator=g) + 0.5 for _ in range(4)]
    bis = [torch.randn(eighth, generator=g) * 0.2 for _ in range(4)]
    pd = pooled.cuda()
    wd, sd_, bd = [w.cuda() for w in ws], [s.cuda() for s in scs], [b.cuda() for b in bis]
    outs = [torch.empty(n, b * b, eighth, device=dev) for b in (1, 2, 3, 6)]
    arr = lambda xs: (C.c_void_p * 4)(*[t.data_ptr() for t in xs])  # noqa: E731
    K, dyn = 64 + 128, 128
    projs, pw, dst = (cabi.PspProjection * len(couts))(), [], []
    for q, cout in enumerate(couts):
        w = (torch.randn(4 * eighth, cout, generator=g) * 3).cuda()
        hi = torch.full((n, cout, K), 7.0, dtype=torch.float16, device=dev)
        lo = torch.full((n, cout, K), -3.0, dtype=torch.float16, device=dev)
        pw.append(w); dst.append((hi, lo))
        projs[q].w, projs[q].dst_hi, projs[q].dst_lo = w.data_ptr(), hi.data_ptr() + 2 * dyn, lo.data_ptr() + 2 * dyn
        projs[q].ld, projs[q].batch_stride, projs[q].cout = K, cout * K, cout
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    t = View(pd.view(-1), n, 1, 50, c4).ct()
    cabi.check(lib.tdn_psp_branch_project(C.byref(t), arr(wd), arr(sd_), arr(bd), eighth, arr(outs), projs, len(couts),
                                          flag.data_ptr(), None), "psp_branch_project")
    torch.cuda.synchronize()
    off, feats = 0, []
    for i, b in enumerate((1, 2, 3, 6)):
        ref = F.relu(pooled[:, off:off + b * b] @ ws[i].t() * scs[i] + bis[i])
        assert max_abs(outs[i].cpu(), ref) < 5e-6
        feats.append(outs[i].double().cpu())
        off += b * b
    for q, cout in enumerate(couts):
        hi, lo = dst[q]
        got = hi.float().cpu() + lo.float().cpu()
        w = pw[q].double().cpu().view(4, eighth, cout)
        want = torch.cat([torch.einsum("co,nbc->nob", w[lv], feats[lv]) for lv in range(4)], 2)        # [n, cout, 50]
        scale = float(want.abs().max())
        assert max_abs(got[:, :, dyn:dyn + 50], want) <= 2e-6 * scale
        assert float((hi[:, :, :dyn] - 7).abs().max()) == 0 and float((hi[:, :, dyn + 50:] - 7).abs().max()) == 0
        assert float((lo[:, :, :dyn] + 3).abs().max()) == 0 and float((lo[:, :, dyn + 50:] + 3).abs().max()) == 0
    assert int(flag.item()) == 0
    assert lib.tdn_psp_branch_project(C.byref(t), arr(wd), arr(sd_), arr(bd), eighth, arr(outs), projs, 0, None, None) == -1


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("n,h,w,cin,cout,affine", [(1, 13, 21, 128, 19, "bias"), (2, 16, 32, 256, 19, "none"),
                                                   (1, 7, 9, 512, 21, "both"), (1, 128, 256, 128, 19, "bias"),
                                                   (1, 5, 5, 64, 3, "bias")])
def test_pointwise_linear_classifier_against_torch(env, n, h, w, cin, cout, affine, split):
    """tdn_pointwise_linear == the nclass 1x1 classifier conv (td4_psp18.py:299, pspnet.py:113, td2_fa.py:316)."""
    lib, cabi, View, dev = env
    g = torch.Generator(device="cuda").manual_seed(cin + cout)
    x = torch.randn(n, h, w, cin, generator=g, device="cuda") * 3
    wt = torch.randn(cout, cin, generator=g, device="cuda") / cin ** 0.5
    scale = torch.rand(cout, generator=g, device="cuda") + 0.5 if affine == "both" else None
    bias = torch.randn(cout, generator=g, device="cuda") if affine in ("bias", "both") else None
    xv = View.alloc(n, h, w, cin, dev, split=split)
    if split:
        xv.base.copy_(x.reshape(-1).half()); xv.lo.copy_((x.reshape(-1) - xv.base.float()).half())
        x = xv.torch()
    else:
        xv.base.copy_(x.reshape(-1))
    ov = View.alloc(n, h, w, cout, dev)
    ov.base.fill_(float("nan"))
    xt, ot = xv.ct(), ov.ct()
    cabi.check(lib.tdn_pointwise_linear(C.byref(xt), wt.data_ptr(), scale.data_ptr() if scale is not None else None,
                                        bias.data_ptr() if bias is not None else None, C.byref(ot), None), "linear")
    torch.cuda.synchronize()
    ref = x.double() @ wt.double().t()
    if scale is not None:
        ref = ref * scale.double()
    if bias is not None:
        ref = ref + bias.double()
    assert max_abs(ov.torch().cpu(), ref.cpu()) <= 2e-6 * float(ref.abs().max()) + 1e-6
    big = View.alloc(n, h, w, 64, dev).ct()
    assert lib.tdn_pointwise_linear(C.byref(xt), wt.data_ptr(), None, None, C.byref(big), None) == -2   # cout > 32
