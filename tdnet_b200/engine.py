"""Host-side frame engine: turns a ModelArch + packed weights into a static list of C-ABI calls.

One `FramePlan` per (path, warm-up | steady) for a fixed input shape.  All buffers are allocated
once (torch is used for device memory and streams only); a plan is a flat list of ctypes calls on
the current CUDA stream, with no host synchronisation, so it can be replayed or graph-captured.

What a plan does, with the reference call sites it replaces (files under
/root/reference/Testing/model/pspnet/):
  stem + maxpool + residual stages ........ resnet.py:204-215
  pyramid pooling slice + concat .......... td4_psp18.py:271-284
  Encoding(pre=False) and (pre=True) ...... transformer.py:28-56
  attention hops .......................... transformer.py:71-92, 126-139
  residual + LayerNorm + FCN head ......... td4_psp18.py:151, 295-299, 306-312
  FIFO push ............................... td4_psp18.py:123-134
  final bilinear upsample ................. td4_psp18.py:227

Exact algebra used (each verified against the oracle in tests/):
  * fc after attention is applied to the values first: softmax rows sum to 1, so
    (A @ V) @ W^T + b == A @ (V @ W^T + b)            (17.2 -> 1.1 GFLOP on the big hop)
  * the queued V and Q of a frame are the stride-4 gather of its full-resolution V and Q (a 1x1
    conv commutes with sub-sampling; MaxPool2d(kernel 1, stride 4) is pure sub-sampling)
  * only this path's slice of each PSP branch conv is computed (64 of 128 output channels)
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, List, Optional

import torch

from . import _cabi
from ._cabi import ACT_LEAKY, ACT_NONE, ACT_RELU, Conv2dDesc, Tensor
from .model import arch as A

_ACT = {"none": ACT_NONE, "relu": ACT_RELU, "leaky_relu": ACT_LEAKY}
_BN_EPS = 1e-5
_LN_EPS = 1e-5
PSP_BINS = (1, 2, 3, 6)
PSP_OFFSETS = (0, 1, 5, 14)


class View:
    """NHWC fp32 view over a torch buffer (element strides), convertible to a C `tdn_tensor`."""

    def __init__(self, base: torch.Tensor, n, h, w, c, sn=None, sh=None, sw=None, offset=0):
        self.base, self.n, self.h, self.w, self.c = base, n, h, w, c
        self.sw = c if sw is None else sw
        self.sh = w * self.sw if sh is None else sh
        self.sn = h * self.sh if sn is None else sn
        self.offset = offset

    @staticmethod
    def alloc(n, h, w, c, device, zero=False):
        fn = torch.zeros if zero else torch.empty
        return View(fn(n * h * w * c, dtype=torch.float32, device=device), n, h, w, c)

    @property
    def ptr(self):
        return self.base.data_ptr() + 4 * self.offset

    def ct(self) -> Tensor:
        return Tensor(self.ptr, None, _cabi.TDN_F32, self.n, self.h, self.w, self.c, self.sn, self.sh, self.sw)

    def channels(self, lo, hi):
        return View(self.base, self.n, self.h, self.w, hi - lo, self.sn, self.sh, self.sw, self.offset + lo)

    def subsample(self, s):
        return View(self.base, self.n, (self.h - 1) // s + 1, (self.w - 1) // s + 1, self.c, self.sn,
                    self.sh * s, self.sw * s, self.offset)

    def rows(self, lo, hi, h, w):
        """Rows [lo, hi) of a [n,1,R,c] matrix view, reshaped to an h x w grid (PSP bins)."""
        assert self.h == 1 and (hi - lo) == h * w
        return View(self.base, self.n, h, w, self.c, self.sn, w * self.sw, self.sw, self.offset + lo * self.sw)

    def tokens(self):
        """[n,h,w,c] dense map -> per-image token matrix [1,1,h*w,c] (+ batch stride)."""
        assert self.sh == self.w * self.sw
        return View(self.base, 1, 1, self.h * self.w, self.c, self.sn, self.h * self.w * self.sw, self.sw, self.offset)

    def torch(self):
        """Dense torch view [n,h,w,c] (tests / FIFO introspection)."""
        return torch.as_strided(self.base, (self.n, self.h, self.w, self.c), (self.sn, self.sh, self.sw, 1), self.offset)


class PackedConv:
    """Device-resident parameters of one convolution: K-major weight [cout][kh][kw][cin4] and the
    per-channel (scale, bias) that folds the conv bias and the eval-mode BatchNorm:
        BN(conv(x) + b) = conv(x) * s + ((b - mean) * s + beta),  s = gamma / sqrt(var + eps)."""

    def __init__(self, spec: A.Conv, sd: Dict[str, torch.Tensor], device, row_slice=None):
        w = sd[spec.name + ".weight"].detach().to(torch.float32)
        cout = w.shape[0]
        bias = sd[spec.name + ".bias"].detach().float() if spec.bias else torch.zeros(cout)
        if spec.bn:
            g, b = sd[spec.bn + ".weight"].float(), sd[spec.bn + ".bias"].float()
            mu, var = sd[spec.bn + ".running_mean"].float(), sd[spec.bn + ".running_var"].float()
            s = g / torch.sqrt(var + _BN_EPS)
            scale, shift = s, (bias - mu) * s + b
        else:
            scale, shift = None, (bias if spec.bias else None)
        w = w.permute(0, 2, 3, 1).contiguous()  # [cout, kh, kw, cin]
        if w.shape[3] % 4:
            pad = 4 - w.shape[3] % 4
            w = torch.nn.functional.pad(w, (0, pad))
        if row_slice is not None:
            lo, hi = row_slice
            w = w[lo:hi].contiguous()
            scale = None if scale is None else scale[lo:hi].contiguous()
            shift = None if shift is None else shift[lo:hi].contiguous()
        self.spec = spec
        self.cout, self.cin = w.shape[0], w.shape[3]
        self.weight = w.to(device)
        self.scale = None if scale is None else scale.contiguous().to(device)
        self.bias = None if shift is None else shift.contiguous().to(device)


class FramePlan:
    def __init__(self):
        self.ops: List[Callable] = []
        self.names: List[str] = []   # per op: state-dict prefix of the conv it runs ('' for the rest)
        self.keep = []           # ctypes objects / tensors referenced by raw pointer
        self.kernel_launches = 0

    def add(self, fn, *args, launches=1, name=""):
        self.keep.append(args)
        self.ops.append((fn, args))
        self.names.append(name)
        self.kernel_launches += launches


class Engine:
    """Static buffers + plans for one (batch, H, W).  `weights` is the model's state dict."""

    def __init__(self, arch: A.ModelArch, state_dict, n, H, W, device, ln_shape):
        self.lib = _cabi.load()
        self.m, self.n, self.H, self.W, self.device = arch, n, H, W, device
        self.h8, self.w8 = A.feature_hw(H, W)
        if tuple(ln_shape) != (self.h8, self.w8):
            # same failure the reference has at any input but 769x1537 (td4_psp18.py:107-110)
            raise RuntimeError(f"Given normalized_shape={list(ln_shape)}, expected input with shape "
                               f"[*, {ln_shape[0]}, {ln_shape[1]}], but got feature map "
                               f"[{n}, {arch.d_v}, {self.h8}, {self.w8}]")
        self.hs, self.ws = (self.h8 - 1) // 4 + 1, (self.w8 - 1) // 4 + 1
        self.pk = self.hs * self.ws                       # keys per frame (P')
        self.pk_pad = (self.pk + 3) // 4 * 4
        self.sd = state_dict
        self._packed: Dict[str, PackedConv] = {}
        self._plans: Dict[tuple, FramePlan] = {}
        self._pool: Dict[tuple, List[torch.Tensor]] = {}
        self._cursor: Dict[tuple, int] = {}
        m = arch
        dev = device
        # FIFO slots: token matrices [n, P'(padded rows for V'), c]
        self.q_slots = [View.alloc(n, 1, self.pk, m.d_k, dev, zero=True) for _ in range(m.depth)]
        self.k_slots = [View.alloc(n, 1, self.pk, m.d_k, dev, zero=True) for _ in range(m.depth)]
        self.v_slots = [View.alloc(n, 1, self.pk, m.d_v, dev, zero=True) for _ in range(m.depth)]
        self.ln_gamma = {p: state_dict[f"layer_norm{p}.ln.weight"].detach().float().reshape(-1).contiguous().to(dev)
                         for p in range(1, m.paths + 1)}
        self.ln_beta = {p: state_dict[f"layer_norm{p}.ln.bias"].detach().float().reshape(-1).contiguous().to(dev)
                        for p in range(1, m.paths + 1)}

    # ------------------------------------------------------------------ helpers
    def packed(self, spec: A.Conv, row_slice=None) -> PackedConv:
        key = spec.name if row_slice is None else f"{spec.name}[{row_slice[0]}:{row_slice[1]}]"
        if key not in self._packed:
            self._packed[key] = PackedConv(spec, self.sd, self.device, row_slice)
        return self._packed[key]

    def buf(self, n, h, w, c, zero=False) -> View:
        """Scratch buffer for the plan being built.  Plans never run concurrently and nothing but the
        FIFO slots survives a frame, so all plans of an engine share one pool: the i-th request of a
        given size in every plan maps to the same storage.  `zero` buffers (padded S / V' matrices whose
        pad columns must stay 0) live in their own pool and are only ever reused at identical shape."""
        key = (n * h * w * c, (n, h, w, c) if zero else None)
        idx = self._cursor.get(key, 0)
        self._cursor[key] = idx + 1
        pool = self._pool.setdefault(key, [])
        if idx == len(pool):
            fn = torch.zeros if zero else torch.empty
            pool.append(fn(n * h * w * c, dtype=torch.float32, device=self.device))
        return View(pool[idx], n, h, w, c)

    def _conv(self, plan: FramePlan, pc: PackedConv, x: View, out: View, residual: Optional[View] = None,
              act=None, stride=None, batch=1, weight_ptr=None, weight_kn=0, in_bs=0, out_bs=0, res_bs=0, w_bs=0,
              k=None, dilation=None, cout=None, scale_ptr="auto", bias_ptr="auto"):
        spec = pc.spec if pc is not None else None
        d = Conv2dDesc()
        d.in_, d.out = x.ct(), out.ct()
        if residual is not None:
            d.residual = residual.ct()
        d.weight = weight_ptr if weight_ptr is not None else pc.weight.data_ptr()
        if scale_ptr == "auto":
            scale_ptr = pc.scale.data_ptr() if (pc is not None and pc.scale is not None) else None
        if bias_ptr == "auto":
            bias_ptr = pc.bias.data_ptr() if (pc is not None and pc.bias is not None) else None
        d.scale, d.bias = scale_ptr, bias_ptr
        d.cout = cout if cout is not None else pc.cout
        kk = k if k is not None else spec.k
        d.kh = d.kw = kk
        d.stride = stride if stride is not None else (spec.stride if spec else 1)
        d.dilation = dilation if dilation is not None else (spec.dilation if spec else 1)
        d.pad = d.dilation * (kk - 1) // 2
        d.act = _ACT[act if act is not None else (spec.act if spec else "none")]
        d.leaky_slope = 0.01
        d.weight_kn, d.batch = weight_kn, batch
        d.in_batch_stride, d.out_batch_stride = in_bs, out_bs
        d.residual_batch_stride, d.weight_batch_stride = res_bs, w_bs
        plan.add(self.lib.tdn_conv2d, C.byref(d), "stream", name=spec.name if spec else "")
        plan.keep.append((d, pc, x, out, residual))

    def _out_hw(self, h, w, c: A.Conv):
        pad = c.pad
        return ((h + 2 * pad - c.dilation * (c.k - 1) - 1) // c.stride + 1,
                (w + 2 * pad - c.dilation * (c.k - 1) - 1) // c.stride + 1)

    # ------------------------------------------------------------------ plan construction
    def plan(self, path: int, steady: bool) -> FramePlan:
        key = (path, steady)
        if key not in self._plans:
            self._plans[key] = self._build(path, steady)
        return self._plans[key]

    def _build(self, path: int, steady: bool) -> FramePlan:
        m, n, lib = self.m, self.n, self.lib
        plan = FramePlan()
        self._cursor = {}
        H, W, h8, w8 = self.H, self.W, self.h8, self.w8

        # --- stem: NCHW image -> NHWC(4) -> conv(s) -> maxpool
        img = self.buf(n, H, W, 4)
        plan.add(lib.tdn_image_to_nhwc, "img", n, 3, H, W, C.byref(self._ct(plan, img)), "stream")
        x = img
        for c in m.stems[path]:
            oh, ow = self._out_hw(x.h, x.w, c)
            y = self.buf(n, oh, ow, c.cout)
            self._conv(plan, self.packed(c), x, y)
            x = y
        y = self.buf(n, (x.h - 1) // 2 + 1, (x.w - 1) // 2 + 1, x.c)
        plan.add(lib.tdn_maxpool3x3s2, C.byref(self._ct(plan, x)), C.byref(self._ct(plan, y)), "stream")
        x = y

        # --- residual stages
        for blk in m.stages[path]:
            identity = x
            if blk.downsample is not None:
                oh, ow = self._out_hw(x.h, x.w, blk.downsample)
                identity = self.buf(n, oh, ow, blk.downsample.cout)
                self._conv(plan, self.packed(blk.downsample), x, identity)
            t = x
            for i, c in enumerate(blk.convs):
                oh, ow = self._out_hw(t.h, t.w, c)
                y = self.buf(n, oh, ow, c.cout)
                last = i == len(blk.convs) - 1
                self._conv(plan, self.packed(c), t, y, residual=identity if last else None)
                t = y
            x = t
        c4 = x
        assert (c4.h, c4.w, c4.c) == (h8, w8, m.c4), (c4.h, c4.w, c4.c)

        # --- pyramid pooling slice -> z  (channels: [c4 slice | 4 x upsampled branch slice])
        pid = m.psp_pid(path)
        half, eighth = m.c4 // 2, m.c4 // 8
        z = self.buf(n, h8, w8, m.c4)
        plan.add(lib.tdn_copy_nhwc, C.byref(self._ct(plan, c4.channels(pid * half, (pid + 1) * half))),
                 C.byref(self._ct(plan, z.channels(0, half))), "stream")
        pooled = self.buf(n, 1, 50, m.c4)
        ws_bytes = int(lib.tdn_psp_pool_workspace_bytes(n, h8, m.c4))
        ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=self.device)
        plan.add(lib.tdn_psp_pool, C.byref(self._ct(plan, c4)), C.byref(self._ct(plan, pooled)), ws.data_ptr(),
                 ws_bytes, "stream", launches=2)
        plan.keep.append(ws)
        for i, (bins, off, c) in enumerate(zip(PSP_BINS, PSP_OFFSETS, A.psp_convs(m, path))):
            pc = self.packed(c, row_slice=(pid * eighth, (pid + 1) * eighth))
            small = self.buf(n, bins, bins, eighth)
            self._conv(plan, pc, pooled.rows(off, off + bins * bins, bins, bins), small)
            plan.add(lib.tdn_bilinear_nhwc, C.byref(self._ct(plan, small)),
                     C.byref(self._ct(plan, z.channels(half + i * eighth, half + (i + 1) * eighth))), "stream")

        # --- Encoding(pre=False): full-resolution V and Q
        enc = A.encoding_convs(m, path)
        v_cur = self.buf(n, h8, w8, m.d_v)
        self._conv(plan, self.packed(enc["w_vs"][0]), z, v_cur)
        q_mid = self.buf(n, h8, w8, m.d_k)
        self._conv(plan, self.packed(enc["w_qs"][0]), z, q_mid)
        q_cur = self.buf(n, h8, w8, m.d_k)
        self._conv(plan, self.packed(enc["w_qs"][1]), q_mid, q_cur)

        # --- attention propagation over the FIFO
        if steady:
            fused = self._attention_chain(plan, path, q_cur, v_cur)
        else:
            fused = v_cur  # td4_psp18.py:142-143: head(layer_norm(v_cur)) while the FIFO fills

        # --- LayerNorm over (H8, W8) + FCN head
        mean = torch.empty(n * m.d_v, dtype=torch.float32, device=self.device)
        rstd = torch.empty_like(mean)
        lws_bytes = int(lib.tdn_layernorm_hw_workspace_bytes(n, h8, w8, m.d_v))
        lws = torch.empty(lws_bytes // 4 + 4, dtype=torch.float32, device=self.device)
        plan.add(lib.tdn_layernorm_hw_stats, C.byref(self._ct(plan, fused)), mean.data_ptr(), rstd.data_ptr(),
                 C.c_float(_LN_EPS), lws.data_ptr(), lws_bytes, "stream", launches=2)
        normed = self.buf(n, h8, w8, m.d_v)
        plan.add(lib.tdn_layernorm_hw_apply, C.byref(self._ct(plan, fused)), mean.data_ptr(), rstd.data_ptr(),
                 self.ln_gamma[path].data_ptr(), self.ln_beta[path].data_ptr(), C.byref(self._ct(plan, normed)),
                 "stream")
        plan.keep.append((mean, rstd, lws))
        hc = A.head_convs(m, path)
        mid = self.buf(n, h8, w8, m.head_mid)
        self._conv(plan, self.packed(hc[0]), normed, mid)
        low = self.buf(n, h8, w8, m.nclass)
        self._conv(plan, self.packed(hc[1]), mid, low)
        plan.add(lib.tdn_upsample_logits, C.byref(self._ct(plan, low)), "out", H, W, "stream")

        # --- Encoding(pre=True) on the stride-4 grid and FIFO push (oldest slot is overwritten by shifting)
        zs = z.subsample(4)
        k_mid = self.buf(n, self.hs, self.ws, m.d_k)
        self._conv(plan, self.packed(enc["w_ks"][0]), zs, k_mid)
        for j in range(m.depth - 1):  # shift: slot j <- slot j+1
            for slots in (self.q_slots, self.k_slots, self.v_slots):
                plan.add(lib.tdn_copy_nhwc, C.byref(self._ct(plan, slots[j + 1])), C.byref(self._ct(plan, slots[j])),
                         "stream")
        last = m.depth - 1
        k_new = self._grid_view(self.k_slots[last])
        self._conv(plan, self.packed(enc["w_ks"][1]), k_mid, k_new)
        plan.add(lib.tdn_copy_nhwc, C.byref(self._ct(plan, v_cur.subsample(4))),
                 C.byref(self._ct(plan, self._grid_view(self.v_slots[last]))), "stream")
        plan.add(lib.tdn_copy_nhwc, C.byref(self._ct(plan, q_cur.subsample(4))),
                 C.byref(self._ct(plan, self._grid_view(self.q_slots[last]))), "stream")
        plan.taps = dict(c4=c4, z=z, q_cur=q_cur, v_cur=v_cur, fused=fused, normed=normed, head=low)
        return plan

    def _grid_view(self, slot: View) -> View:
        """FIFO slot [n,1,P',c] seen as the [n,hs,ws,c] grid it was sampled from."""
        return View(slot.base, slot.n, self.hs, self.ws, slot.c, slot.sn, self.ws * slot.sw, slot.sw, slot.offset)

    def _ct(self, plan: FramePlan, v: View) -> Tensor:
        t = v.ct()
        plan.keep.append((t, v))
        return t

    def _attention_chain(self, plan: FramePlan, path: int, q_cur: View, v_cur: View) -> View:
        """v_2_, v_3_, v_4_ of td4_psp18.py:145-147 (or the single hop of td2_psp50.py:120) with the fc
        folded into the values.  Returns the map `v_last + v_cur`."""
        m, n, lib = self.m, self.n, self.lib
        hops = m.hop_modules(path)
        pk, pkp, pq_full = self.pk, self.pk_pad, self.h8 * self.w8
        carry = None  # v_{j} + V_queue[j], the value source of the next hop
        for j, name in enumerate(hops):
            last = j == len(hops) - 1
            v_src = self.v_slots[j] if carry is None else carry
            # V' = fc(v_src): [n, P', d_v] (rows padded to a multiple of 4 with zeros for the next GEMM's K)
            vp = self.buf(n, 1, pkp, m.d_v, zero=True)
            vp_rows = View(vp.base, n, 1, pk, m.d_v, vp.sn, vp.sh, vp.sw, vp.offset)
            self._conv(plan, self.packed(A.fc_conv(m, name)), v_src, vp_rows)
            # S = q k^T / 8 -> softmax
            if last:
                q, pq = q_cur.tokens(), pq_full
                q_bs = q_cur.sn
            else:
                q, pq = View(self.q_slots[j + 1].base, 1, 1, pk, m.d_k), pk
                q_bs = self.q_slots[j + 1].sn
            s = self.buf(n, 1, pq, pkp, zero=True)
            s_one = View(s.base, 1, 1, pq, pk, s.sn, s.sh, s.sw)
            self._conv(plan, None, q, s_one, batch=n, weight_ptr=self.k_slots[j].ptr, k=1, cout=pk,
                       in_bs=q_bs, out_bs=s.sn, w_bs=self.k_slots[j].sn, scale_ptr=None, bias_ptr=None)
            plan.add(lib.tdn_softmax_rows, s.ptr, n * pq, pk, pkp, C.c_float(1.0 / float(m.d_k) ** 0.5), "stream")
            plan.keep.append(s)
            # out = S @ V' + residual
            if last:
                out = self.buf(n, self.h8, self.w8, m.d_v)
                o_one, res_one, res_bs = out.tokens(), v_cur.tokens(), v_cur.sn
            else:
                out = self.buf(n, 1, pk, m.d_v)
                o_one = View(out.base, 1, 1, pk, m.d_v)
                res_one = View(self.v_slots[j + 1].base, 1, 1, pk, m.d_v)
                res_bs = self.v_slots[j + 1].sn
            s_in = View(s.base, 1, 1, pq, pkp, s.sn, s.sh, s.sw)
            self._conv(plan, None, s_in, o_one, residual=res_one, batch=n, weight_ptr=vp.ptr, weight_kn=1, k=1,
                       cout=m.d_v, in_bs=s.sn, out_bs=out.sn, res_bs=res_bs, w_bs=vp.sn, scale_ptr=None,
                       bias_ptr=None)
            carry = out
        return carry

    # ------------------------------------------------------------------ execution
    def run(self, plan: FramePlan, img_ptr: int, out_ptr: int, stream: int, probe=None):
        """Enqueue the frame.  probe = (op_name, event_before, event_after) brackets one op with CUDA
        events (bench.py times the dominant kernel live this way)."""
        subst = {"img": img_ptr, "out": out_ptr, "stream": stream}
        for i, (fn, args) in enumerate(plan.ops):
            hit = probe is not None and plan.names[i] == probe[0]
            if hit:
                probe[1].record()
            rc = fn(*[subst[a] if isinstance(a, str) else a for a in args])
            if hit:
                probe[2].record()
            if rc != 0:
                _cabi.check(rc, fn.__name__)
