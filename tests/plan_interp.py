"""CPU interpreter for engine frame plans.  TEST INFRASTRUCTURE ONLY (never imported by tdnet_b200/).

An `Engine` built on `torch.device("cpu")` allocates its buffers in host memory and records the same flat list of
C-ABI calls (function, ctypes descriptors) it would enqueue on the GPU.  This module executes such a plan with plain
torch CPU ops, reading and writing the very buffers the descriptors point at (through their raw addresses), each op
implemented from the contract stated in include/tdnet_b200.h -- not from the CUDA sources.  It checks everything the
HOST side decides: op order, views / strides / offsets, weight packing and BatchNorm folding, padded scratch, the
algebraic rewrites (fc folded into the values, 1x1 convs on sub-sampled views).  It says nothing about the kernels;
those are compared with the oracle on the GPU (tests/test_*_gpu.py).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from tdnet_b200 import _cabi


def _flat(addr: int, numel: int, dtype: torch.dtype) -> torch.Tensor:
    size = numel * torch.empty((), dtype=dtype).element_size()
    return torch.frombuffer((C.c_char * size).from_address(addr), dtype=dtype, count=numel)


def _strided(addr, dtype, t, elem_offset=0):
    extent = (t.n - 1) * t.stride_n + (t.h - 1) * t.stride_h + (t.w - 1) * t.stride_w + t.c
    isz = 4 if dtype == torch.float32 else 2
    flat = _flat(addr + elem_offset * isz, extent, dtype)
    return torch.as_strided(flat, (t.n, t.h, t.w, t.c), (t.stride_n, t.stride_h, t.stride_w, 1))


def read(t: _cabi.Tensor, elem_offset=0) -> torch.Tensor:
    """tdn_tensor -> dense fp32 [n,h,w,c] copy."""
    if t.dtype == _cabi.TDN_F32:
        return _strided(t.data, torch.float32, t, elem_offset).clone()
    return (_strided(t.data, torch.float16, t, elem_offset).float()
            + _strided(t.data_lo, torch.float16, t, elem_offset).float())


def write(t: _cabi.Tensor, val: torch.Tensor, elem_offset=0):
    val = val.float()
    assert tuple(val.shape) == (t.n, t.h, t.w, t.c), (tuple(val.shape), (t.n, t.h, t.w, t.c))
    if t.dtype == _cabi.TDN_F32:
        _strided(t.data, torch.float32, t, elem_offset).copy_(val)
        return
    hi = val.half()
    _strided(t.data, torch.float16, t, elem_offset).copy_(hi)
    _strided(t.data_lo, torch.float16, t, elem_offset).copy_((val - hi.float()).half())


def _vec(addr, n):
    return None if not addr else _flat(addr, n, torch.float32).clone()


def _act(x, act, slope):
    if act == _cabi.ACT_RELU:
        return F.relu(x)
    if act == _cabi.ACT_LEAKY:
        return F.leaky_relu(x, slope)
    return x


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


def _bilinear(x_nhwc, h, w):
    return _nhwc(F.interpolate(_nchw(x_nhwc), (h, w), mode="bilinear", align_corners=True))


# ------------------------------------------------------------------------------------------------ ops
def op_image_to_nhwc(nchw, n, c, h, w, out, stream):
    img = _flat(nchw, n * c * h * w, torch.float32).view(n, c, h, w)
    o = torch.zeros(n, h, w, out._obj.c)
    o[..., :c] = _nhwc(img)
    write(out._obj, o)


def op_stem_conv_pool_tc_act(nchw, hwc_u8, lut, n, h, w, weight_tc, scale, bias, out, act, slope, flag, stream):
    """include/tdnet_b200.h: weight_tc = fp16 [2 planes][7 ky][4 kx pairs][64 cout][2 kx][4 c] (zero for kx = 7, c = 3)."""
    assert nchw and not hwc_u8
    img = _flat(nchw, n * 3 * h * w, torch.float32).view(n, 3, h, w)
    planes = _flat(weight_tc, 2 * 7 * 4 * 64 * 2 * 4, torch.float16).view(2, 7, 4, 64, 2, 4).float()
    wt = (planes[0] + planes[1]).permute(2, 4, 0, 1, 3).reshape(64, 4, 7, 8)        # [cout][c][ky][kx]
    assert float(wt[:, 3].abs().max()) == 0.0 and float(wt[:, :, :, 7].abs().max()) == 0.0
    y = F.conv2d(img.double(), wt[:, :3, :, :7].double(), None, 2, 3).float()
    y = y * _vec(scale, 64).view(1, 64, 1, 1) + _vec(bias, 64).view(1, 64, 1, 1)
    y = _act(y, act, slope.value if hasattr(slope, "value") else slope)
    write(out._obj, _nhwc(F.max_pool2d(y, 3, 2, 1)))


def op_stem_conv_pool_tc(nchw, hwc_u8, lut, n, h, w, weight_tc, scale, bias, out, flag, stream):
    op_stem_conv_pool_tc_act(nchw, hwc_u8, lut, n, h, w, weight_tc, scale, bias, out, _cabi.ACT_RELU, 0.0, flag, stream)


def op_stem_conv_pool(nchw, n, h, w, weight, scale, bias, out, stream):
    """fp32 CUDA-core stem: weight fp32 [147][64] with k = (c*7 + ky)*7 + kx; conv7x7 s2 p3 + BN + ReLU + maxpool."""
    img = _flat(nchw, n * 3 * h * w, torch.float32).view(n, 3, h, w)
    wt = _flat(weight, 147 * 64, torch.float32).view(3, 7, 7, 64).permute(3, 0, 1, 2)
    y = F.conv2d(img, wt, None, 2, 3) * _vec(scale, 64).view(1, 64, 1, 1) + _vec(bias, 64).view(1, 64, 1, 1)
    write(out._obj, _nhwc(F.max_pool2d(F.relu(y), 3, 2, 1)))


PSP_BINS = (1, 2, 3, 6)


def op_psp_pool(x, out, ws, ws_bytes, stream):
    v = _nchw(read(x._obj))
    n, c, h, _ = v.shape
    assert ws_bytes >= n * h * 12 * c * 4
    rows = [F.adaptive_avg_pool2d(v, b).permute(0, 2, 3, 1).reshape(n, b * b, c) for b in PSP_BINS]
    write(out._obj, torch.cat(rows, 1).reshape(n, 1, 50, c))


def op_psp_branch_convs(pooled, w, scale, bias, eighth, out, stream):
    p = read(pooled._obj)                                  # [n,1,50,c4]
    n, c4 = p.shape[0], p.shape[3]
    off = 0
    for i, b in enumerate(PSP_BINS):
        wt = _flat(w[i], eighth * c4, torch.float32).view(eighth, c4)
        y = p[:, 0, off:off + b * b] @ wt.t() * _vec(scale[i], eighth) + _vec(bias[i], eighth)
        _flat(out[i], n * b * b * eighth, torch.float32).view(n, b * b, eighth).copy_(F.relu(y))
        off += b * b


def op_psp_branch_project(pooled, w, scale, bias, eighth, out, proj, n_proj, flag, stream):
    """include/tdnet_b200.h: the branch convs, then dst[i][o][bin] = sum_c w[lv(bin) * eighth + c][o] * b_lv[i][bin][c]
    stored as SPLIT16 into column `bin` of each projection's destination."""
    op_psp_branch_convs(pooled, w, scale, bias, eighth, out, stream)
    n = pooled._obj.n
    for q in range(n_proj):
        pr = proj[q]
        wq = _flat(pr.w, pr.cout * 4 * eighth, torch.float32).view(4, eighth, pr.cout).permute(2, 0, 1)
        off = 0
        for lv, b in enumerate(PSP_BINS):
            feat = _flat(out[lv], n * b * b * eighth, torch.float32).view(n, b * b, eighth)
            t = torch.einsum("oc,nbc->nob", wq[:, lv].double(), feat.double()).float()      # [n, cout, bins^2]
            for i in range(n):
                ext = (pr.cout - 1) * pr.ld + b * b
                base = i * pr.batch_stride + off
                hi = torch.as_strided(_flat(pr.dst_hi + 2 * base, ext, torch.float16), (pr.cout, b * b), (pr.ld, 1))
                lo = torch.as_strided(_flat(pr.dst_lo + 2 * base, ext, torch.float16), (pr.cout, b * b), (pr.ld, 1))
                h = t[i].half()
                hi.copy_(h)
                lo.copy_((t[i] - h.float()).half())
            off += b * b


def op_psp_concat(x, small, eighth, z, stream):
    xv = read(x._obj)
    n, h, w, cx = xv.shape
    parts = [xv]
    for i, b in enumerate(PSP_BINS):
        sm = _flat(small[i], n * b * b * eighth, torch.float32).view(n, b, b, eighth)
        parts.append(_bilinear(sm, h, w))
    write(z._obj, torch.cat(parts, 3))


def op_maxpool3x3s2(x, out, stream):
    write(out._obj, _nhwc(F.max_pool2d(_nchw(read(x._obj)), 3, 2, 1)))


def op_conv2d(dref, stream):
    d = dref._obj
    for b in range(d.batch):
        x = read(d.in_, b * d.in_batch_stride)
        cin, K = d.in_.c, d.kh * d.kw * d.in_.c
        wflat = _flat(d.weight + 4 * b * d.weight_batch_stride, K * d.cout, torch.float32)
        w = wflat.view(K, d.cout).t() if d.weight_kn else wflat.view(d.cout, K)
        w = w.reshape(d.cout, d.kh, d.kw, cin).permute(0, 3, 1, 2)
        y = _nhwc(F.conv2d(_nchw(x), w, None, d.stride, d.pad, d.dilation))
        scale, bias = _vec(d.scale, d.cout), _vec(d.bias, d.cout)
        if scale is not None:
            y = y * scale
        if bias is not None:
            y = y + bias
        if d.residual.data:
            y = y + read(d.residual, b * d.residual_batch_stride)
        write(d.out, _act(y, d.act, d.leaky_slope), b * d.out_batch_stride)


def op_conv2d_tc(dref, stream):
    d = dref._obj
    assert d.in_.dtype == _cabi.TDN_SPLIT16 and d.in_.c % 64 == 0
    x = read(d.in_)
    n, cin = d.in_.n, d.in_.c
    K = d.kh * d.kw * cin
    stride = max(d.stride, 1)
    pad = d.dilation * (d.kh - 1) // 2
    assert d.weight_ld >= K
    outs = []
    for i in range(n):
        off = i * d.weight_batch_stride if d.weight_batched else 0
        rows = (d.cout - 1) * d.weight_ld + K
        wh = torch.as_strided(_flat(d.weight_hi + 2 * off, rows, torch.float16), (d.cout, K), (d.weight_ld, 1)).float()
        wl = torch.as_strided(_flat(d.weight_lo + 2 * off, rows, torch.float16), (d.cout, K), (d.weight_ld, 1)).float()
        w = (wh + wl).reshape(d.cout, d.kh, d.kw, cin).permute(0, 3, 1, 2)
        outs.append(F.conv2d(_nchw(x[i:i + 1]).double(), w.double(), None, stride, pad, d.dilation).float())
    y = _nhwc(torch.cat(outs))
    assert tuple(y.shape) == (d.out.n, d.out.h, d.out.w, d.out.c), (tuple(y.shape), (d.out.n, d.out.h, d.out.w, d.out.c))
    scale = _vec(d.scale, d.cout)
    if scale is not None:
        y = y * scale
    if d.bias:
        if d.bias_along_m:
            m = d.out.h * d.out.w
            y = y + _vec(d.bias, m).view(1, d.out.h, d.out.w, 1)
        else:
            y = y + _vec(d.bias, d.cout)
    if d.residual.data:
        y = y + read(d.residual)
    write(d.out, _act(y, d.act, d.leaky_slope))


def op_attention_tc(dref, stream):
    d = dref._obj
    assert d.d_k == 64 and d.d_v % 128 == 0 and d.vt_ld % 64 == 0 and d.vt_ld >= d.pk

    def mat(hi, lo, rows, cols, ld, off):
        ext = (rows - 1) * ld + cols
        return (torch.as_strided(_flat(hi + 2 * off, ext, torch.float16), (rows, cols), (ld, 1)).float()
                + torch.as_strided(_flat(lo + 2 * off, ext, torch.float16), (rows, cols), (ld, 1)).float())

    res = read(d.residual) if d.residual.data else None
    out = torch.empty(d.n, 1, d.pq, d.d_v)
    for i in range(d.n):
        q = mat(d.q_hi, d.q_lo, d.pq, 64, d.q_ld, i * d.q_batch_stride)
        k = mat(d.k_hi, d.k_lo, d.pk, 64, d.k_ld, i * d.k_batch_stride)
        vt = mat(d.vt_hi, d.vt_lo, d.d_v, d.vt_ld, d.vt_ld, i * d.vt_batch_stride)
        assert torch.isfinite(vt).all()                 # pad columns must be zero / finite
        p = torch.softmax((q @ k.t()) / 8.0, dim=1)
        out[i, 0] = p @ vt[:, :d.pk].t()
    if res is not None:
        out = out + res.reshape(out.shape)
    write(d.out, out.reshape(d.out.n, d.out.h, d.out.w, d.out.c))


def op_pointwise_linear(x, weight, scale, bias, out, stream):
    v = read(x._obj)
    cout, cin = out._obj.c, x._obj.c
    w = _flat(weight, cout * cin, torch.float32).view(cout, cin)
    y = v @ w.t()
    if scale:
        y = y * _vec(scale, cout)
    if bias:
        y = y + _vec(bias, cout)
    write(out._obj, y)


def op_softmax_rows(s, rows, cols, ld, scale, stream):
    flat = _flat(s, (rows - 1) * ld + cols, torch.float32)
    m = torch.as_strided(flat, (rows, cols), (ld, 1))
    m.copy_(torch.softmax(m * scale.value if hasattr(scale, "value") else m * scale, dim=1))


def op_copy_nhwc(x, out, stream):
    write(out._obj, read(x._obj))


def op_bilinear_nhwc(x, out, stream):
    write(out._obj, _bilinear(read(x._obj), out._obj.h, out._obj.w))


def op_fa_context(key, value, f, ws, ws_bytes, stream):
    k, v = read(key._obj), read(value._obj)
    n, h, w, c = v.shape
    assert ws_bytes >= n * ((h * w + 255) // 256) * 32 * c * 4
    kn = F.normalize(k.reshape(n, h * w, 32), p=2, dim=2, eps=1e-12)
    _flat(f, n * 32 * c, torch.float32).view(n, 32, c).copy_(kn.transpose(1, 2) @ v.reshape(n, h * w, c))


def op_fa_apply(query, f, out, out_scale, flag, stream):
    q = read(query._obj)
    n, h, w, _ = q.shape
    c = out._obj.c
    fm = _flat(f, n * 32 * c, torch.float32).view(n, 32, c)
    qn = F.normalize(q.reshape(n, h * w, 32), p=2, dim=2, eps=1e-12)
    write(out._obj, (qn @ fm).reshape(n, h, w, c) * out_scale.value)


def op_add_upsampled(a, b, up, out, stream):
    s = read(a._obj) + read(b._obj)
    if up is not None:
        s = _bilinear(read(up._obj), out._obj.h, out._obj.w) + s
    write(out._obj, s)


def op_layernorm_hw_stats(x, mean, rstd, eps, ws, ws_bytes, stream):
    v = read(x._obj).double()
    n, h, w, c = v.shape
    mu = v.mean(dim=(1, 2))
    var = v.var(dim=(1, 2), unbiased=False)
    e = eps.value if hasattr(eps, "value") else eps
    _flat(mean, n * c, torch.float32).copy_(mu.reshape(-1).float())
    _flat(rstd, n * c, torch.float32).copy_((1.0 / torch.sqrt(var + e)).reshape(-1).float())


def op_layernorm_hw_apply(x, mean, rstd, gamma, beta, out, stream):
    v = read(x._obj)
    n, h, w, c = v.shape
    mu, rs = _flat(mean, n * c, torch.float32).view(n, 1, 1, c), _flat(rstd, n * c, torch.float32).view(n, 1, 1, c)
    g, b = _flat(gamma, h * w, torch.float32).view(1, h, w, 1), _flat(beta, h * w, torch.float32).view(1, h, w, 1)
    write(out._obj, (v - mu) * rs * g + b)


def op_upsample_logits(x, out, H, W, stream):
    v = read(x._obj)
    n, _, _, c = v.shape
    _flat(out, n * c * H * W, torch.float32).view(n, c, H, W).copy_(
        F.interpolate(_nchw(v), (H, W), mode="bilinear", align_corners=True))


def _full_res_logits(x, H, W):
    return F.interpolate(_nchw(read(x._obj)), (H, W), mode="bilinear", align_corners=True)


def op_upsample_argmax(x, out, H, W, stream):
    lab = _full_res_logits(x, H, W).argmax(1).to(torch.uint8)
    _flat(out, lab.numel(), torch.uint8).view(lab.shape).copy_(lab)


def op_upsample_argmax_sampled(x, out, H, W, ys, xs, Ho, Wo, stream):
    yi = _flat(ys, Ho, torch.int32).long()
    xi = _flat(xs, Wo, torch.int32).long()
    lab = _full_res_logits(x, H, W).argmax(1)[:, yi][:, :, xi].to(torch.uint8)
    _flat(out, lab.numel(), torch.uint8).view(lab.shape).copy_(lab)


OPS = {
    "tdn_image_to_nhwc": op_image_to_nhwc, "tdn_stem_conv_pool_tc": op_stem_conv_pool_tc,
    "tdn_stem_conv_pool": op_stem_conv_pool, "tdn_psp_pool": op_psp_pool, "tdn_psp_branch_convs": op_psp_branch_convs,
    "tdn_psp_concat": op_psp_concat, "tdn_psp_branch_project": op_psp_branch_project, "tdn_stem_conv_pool_tc_act": op_stem_conv_pool_tc_act, "tdn_maxpool3x3s2": op_maxpool3x3s2, "tdn_conv2d": op_conv2d,
    "tdn_conv2d_tc": op_conv2d_tc, "tdn_attention_tc": op_attention_tc, "tdn_softmax_rows": op_softmax_rows,
    "tdn_copy_nhwc": op_copy_nhwc, "tdn_pointwise_linear": op_pointwise_linear, "tdn_bilinear_nhwc": op_bilinear_nhwc, "tdn_fa_context": op_fa_context,
    "tdn_fa_apply": op_fa_apply, "tdn_add_upsampled": op_add_upsampled,
    "tdn_layernorm_hw_stats": op_layernorm_hw_stats, "tdn_layernorm_hw_apply": op_layernorm_hw_apply,
    "tdn_upsample_logits": op_upsample_logits, "tdn_upsample_argmax": op_upsample_argmax,
    "tdn_upsample_argmax_sampled": op_upsample_argmax_sampled,
}


def run_plan(plan, subst, last_op=None):
    """Execute every op of `plan` in order; `subst` maps the symbolic arguments ("img", "img2", "out") to host
    addresses.  Stream arguments and fork / join marks are ignored (sequential execution is one valid schedule).
    `last_op` replaces the plan's last op (the labels / preview variants of the output stage)."""
    ops = plan.ops if last_op is None else list(plan.ops[:-1]) + [last_op]
    for fn, args in ops:
        if fn in ("fork", "join"):
            continue
        impl = OPS.get(fn.__name__)
        if impl is None:
            raise NotImplementedError(f"plan_interp: no CPU model of {fn.__name__}")
        impl(*[subst.get(a) if isinstance(a, str) else a for a in args])
