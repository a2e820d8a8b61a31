"""Shared implementation of the two TD models: parameter containers with the reference's state-dict
layout, the Q/K/V FIFO, and `forward(img, pos_id)` dispatching to the CUDA frame engine.

API kept from the reference (Testing/model/pspnet/td4_psp18.py:29-240, td2_psp50.py:29-166):
constructor kwargs, `eval()/to()`, `state_dict()` keys and shapes (checkpoints load with
strict=True), `forward(img, pos_id) -> fp32 [n, nclass, H, W]` on the input's device, the observable
`Q_queue / K_queue / V_queue` lists, `pretrained_mp_load()`.
"""
from __future__ import annotations

import os
from typing import Dict

import torch
import torch.nn as nn

from . import arch as A


class _Group(nn.Module):
    """Pure parameter container; never called.  Loading a state dict into a sub-module (the reference's own idiom in
    td2_fa.pretrained_init, or `model.pretrained1.load_state_dict(...)`) invalidates the packed device weights of the
    model it belongs to."""

    def forward(self, *a, **k):
        # The sub-networks are callable like the reference's (`model.pretrained1(img)` -> c4, td4_psp18.py:70-77,
        # resnet.py:204-215); every other group is a pure parameter container.
        root = getattr(self, "_td_root", None)
        root = root() if root is not None else None
        name = getattr(self, "_td_name", "")
        if root is not None and name.startswith("pretrained") and len(a) == 1 and not k:
            return root._subnetwork(name, a[0])
        raise RuntimeError("parameter container; the computation runs in the CUDA engine")

    def _load_from_state_dict(self, *a, **k):
        out = super()._load_from_state_dict(*a, **k)
        root = getattr(self, "_td_root", None)
        if root is not None and root() is not None:
            root().invalidate()
        return out


def _attach(root: nn.Module, key: str, shape, kind: str):
    import weakref
    *parents, leaf = key.split(".")
    mod = root
    for p in parents:
        if not hasattr(mod, p):
            g = _Group()
            object.__setattr__(g, "_td_root", weakref.ref(root))     # plain attribute: not a sub-module, no cycle
            object.__setattr__(g, "_td_name", p if mod is root else "")
            mod.add_module(p, g)
        mod = getattr(mod, p)
    if kind == "param":
        t = torch.zeros(shape)
        if leaf == "weight" and len(shape) <= 2 and not key.endswith("fc.weight"):
            t.fill_(1.0)  # norm gammas default to 1 as in the reference constructors
        mod.register_parameter(leaf, nn.Parameter(t, requires_grad=False))
    elif kind == "buffer":
        mod.register_buffer(leaf, torch.ones(shape) if leaf == "running_var" else torch.zeros(shape))
    else:
        mod.register_buffer(leaf, torch.zeros(shape, dtype=torch.long))


def _default_init_(module: nn.Module, seed=0):
    """Random init for the no-checkpoint case (the reference also runs with random weights when the
    file is missing, td4_psp18.py:239-240).  He-normal on conv kernels (resnet.py:162-165)."""
    g = torch.Generator().manual_seed(seed)
    for name, p in module.named_parameters():
        if p.dim() == 4:
            fan = p.shape[0] * p.shape[2] * p.shape[3]
            p.data.copy_(torch.randn(p.shape, generator=g) * (2.0 / fan) ** 0.5)
        elif p.dim() == 2 and name.endswith("fc.weight"):
            p.data.copy_(torch.randn(p.shape, generator=g) * 0.01)


class TDModel(nn.Module):
    ARCH = None          # 'td4_psp18' | 'td2_psp50' | 'pspnet'
    PATHS = None
    MAX_ENGINES = 2      # engines (packed weights + plans + scratch) kept alive, keyed by (n, H, W, device, mode)
    BACKBONES = ("resnet50", "resnet34", "resnet18")

    def __init__(self, nclass=21, norm_layer=None, backbone=None, dilated=True, aux=True, multi_grid=True,
                 path_num=None, model_path=None, ln_shape=(97, 193)):
        super().__init__()
        if self.ARCH == "pspnet":
            if backbone not in self.BACKBONES:
                raise RuntimeError("unknown backbone: {}".format(backbone))   # pspnet.py:67-68
        else:
            assert backbone in self.BACKBONES                                # td4_psp18.py:52 / td2_psp50.py:52
        assert path_num == self.PATHS
        if not (dilated and multi_grid):
            raise RuntimeError("tdnet_b200 implements the dilated, multi-grid backbone the reference tests ship")
        # BatchNorm is folded into the convolution epilogues (eval mode, running statistics + affine, with the
        # activation of the reference's BatchNorm2d wrapper): any other normalisation layer would silently compute
        # something else, so only the reference's own default (or None) is accepted.
        if norm_layer is not None and getattr(norm_layer, "__name__", "") != "BatchNorm2d":
            raise RuntimeError("tdnet_b200 folds eval-mode BatchNorm2d into its kernels; norm_layer={!r} is not "
                               "supported".format(norm_layer))
        self.psp_path = model_path
        self.path_num = path_num
        self.nclass = nclass
        self.backbone = backbone
        self.norm_layer = norm_layer
        self.arch = A.build_arch(self.ARCH, backbone, nclass)
        self.expansion = self.arch.c4 // 512
        self.ln_shape = tuple(ln_shape)
        for key, (shape, kind) in A.parameter_table(self.arch, self.ln_shape).items():
            _attach(self, key, shape, kind)
        _default_init_(self)
        self._engines: Dict[tuple, object] = {}
        self.pretrained_mp_load()
        self._fifo_fill = 0            # frames in the FIFO (the tensors live in the engine's device slots)
        self._fifo_manual = None       # lists handed in through buffer_contral() / assignment (reference API)
        # 'tc': tcgen05 exact-mode kernels (product path on B200); 'simt': fp32 CUDA-core kernels only;
        # 'tc_fast': opt-in single-product tensor-core mode (NOT fp32-faithful, never the parity gate).
        self.engine_mode = os.environ.get("TDNET_B200_ENGINE", "tc")
        # Replay each static frame plan through a CUDA graph (removes ~60 launch gaps).  All plans of a model are
        # built, run once on a scratch frame and captured when the engine is created (first forward of a new input
        # shape, or prepare()), never inside a later frame.
        self.use_cuda_graph = os.environ.get("TDNET_B200_CUDA_GRAPH", "1") != "0"
        # SPLIT16 range guard (|x| > 6e4 overflows an fp16 plane): the kernels raise a flag in host-mapped pinned memory
        # that forward() examines after every frame without a copy or a synchronisation (_poll_range_flag).

    # ---- reference API ---------------------------------------------------------------------
    def pretrained_mp_load(self):
        if self.psp_path is not None:
            if os.path.isfile(self.psp_path):
                print("Loading pretrained model from '{}'".format(self.psp_path))
                self.load_state_dict(torch.load(self.psp_path, map_location="cpu"), strict=True)
            else:
                print("No pretrained found at '{}'".format(self.psp_path))

    def invalidate(self):
        """Drop the packed device weights, frame plans and CUDA graphs; they are rebuilt from the current parameters at
        the next forward.  Called automatically by load_state_dict() (of the model or any sub-module), .to() / .half() /
        .float(), set_ln_shape(); call it yourself after editing parameters in place (`p.data.copy_(...)`)."""
        self._engines.clear()

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        if getattr(self, "_engines", None):
            self.invalidate()
        return out

    def set_ln_shape(self, h8, w8):
        """Re-create the LayerNorm affine for another feature-map size (the reference hard-codes
        [97,193], td4_psp18.py:107-110, i.e. 769x1537 inputs; tests patch it the same way).  A no-op when the shape
        already matches (loaded affine weights are kept)."""
        if tuple(self.ln_shape) == (h8, w8):
            return
        self.ln_shape = (h8, w8)
        dev = next(self.parameters()).device
        for p in range(1, self.PATHS + 1):
            ln = getattr(self, f"layer_norm{p}").ln
            ln.weight = nn.Parameter(torch.ones(h8, w8, device=dev), requires_grad=False)
            ln.bias = nn.Parameter(torch.zeros(h8, w8, device=dev), requires_grad=False)
        self.invalidate()

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self.invalidate()  # packed device weights are rebuilt lazily
        return out

    def reset(self):
        """Start a new clip (the reference needs a new model instance for that)."""
        self._fifo_fill = 0
        self._fifo_manual = None

    # The observable FIFO of the reference (`Q_queue / K_queue / V_queue`, td4_psp18.py:118-134).  The entries live in the
    # engine's device slots (SPLIT16 planes); the lists are materialised ON ACCESS as new fp32 tensors [n, P', d] -- oldest
    # first, like the reference's -- so a caller holding them is not affected by later frames, and the frame loop itself
    # launches nothing for them.
    def _fifo_list(self, which):
        if self._fifo_manual is not None:
            return self._fifo_manual[which]
        eng = self._engines.get(getattr(self, "_active_key", None))
        if eng is None or self._fifo_fill == 0:
            return []
        depth, fill = self.arch.depth, self._fifo_fill
        slots = (eng.q_slots, eng.k_slots, eng.v_slots)[which]
        d = self.arch.d_v if which == 2 else self.arch.d_k
        return [slots[depth - fill + j].torch().reshape(eng.n, -1, d).clone() for j in range(fill)]

    def _fifo_assign(self, which, value):
        if self._fifo_manual is None:
            self._fifo_manual = [self._fifo_list(0), self._fifo_list(1), self._fifo_list(2)]
        self._fifo_manual[which] = list(value)
        if not any(self._fifo_manual):
            self._fifo_manual, self._fifo_fill = None, 0

    Q_queue = property(lambda self: self._fifo_list(0), lambda self, v: self._fifo_assign(0, v))
    K_queue = property(lambda self: self._fifo_list(1), lambda self, v: self._fifo_assign(1, v))
    V_queue = property(lambda self: self._fifo_list(2), lambda self, v: self._fifo_assign(2, v))

    def buffer_contral(self, q, k, v):
        """td4_psp18.py:123-134, kept for API parity: appends to the observable lists.  The engine pushes its own
        FIFO inside forward(); lists edited by hand are only reported back, they do not feed the kernels."""
        qs, ks, vs = self.Q_queue, self.K_queue, self.V_queue
        assert len(qs) == len(vs)
        assert len(qs) == len(ks)
        qs.append(q), vs.append(v), ks.append(k)
        if len(qs) > self.arch.depth:
            qs.pop(0), vs.pop(0), ks.pop(0)
        self._fifo_manual = [qs, ks, vs]

    # ---- engine ----------------------------------------------------------------------------
    def _engine(self, img: torch.Tensor, shape):
        from ..engine import Engine
        n, c, h, w = shape
        key = (n, h, w, img.device.index, self.engine_mode)
        eng = self._engines.get(key)
        if eng is None:
            sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
            eng = Engine(self.arch, sd, n, h, w, img.device, self.ln_shape, mode=self.engine_mode)
            while len(self._engines) >= self.MAX_ENGINES:          # small LRU: alternating two input shapes does not
                self._engines.pop(next(iter(self._engines)))       # re-pack every weight on each switch
            self._engines[key] = eng
            if self.use_cuda_graph:
                eng.prepare_graphs()
            self.reset()       # a different input shape starts a new clip: the FIFO lives in the engine
        elif key != getattr(self, "_active_key", key):
            self._engines[key] = self._engines.pop(key)            # most recently used last
            self.reset()
        self._active_key = key
        return eng

    def prepare(self, n, h, w, device=None):
        """Build the engine for [n,3,h,w] inputs now: weight packing, all frame plans, their CUDA graphs.  forward()
        does the same on the first frame of a new shape; call this to keep it out of the first frame's latency.
        Starts a new clip."""
        device = torch.device(device) if device is not None else next(self.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("tdnet_b200 runs on a CUDA (sm_100) device only; move the model with .to('cuda') first")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(device):
            self._engine(torch.empty(0, device=device), (n, 3, h, w))
        self.reset()
        return self

    def _poll_range_flag(self, eng):
        """Non-blocking range guard: the kernels set eng.range_flag (pinned, device-mapped host memory) when a SPLIT16
        output overflows; the host looks at it after every frame without a copy or a synchronisation, so an overflow
        raises as soon as a frame that saw it has finished -- at the latest on the call after -- never silently."""
        if int(eng.range_flag[0]) != 0:
            torch.cuda.synchronize(eng.device)     # let the offending frame finish before the flag is re-armed
            eng.range_flag.zero_()
            raise RuntimeError(self._RANGE_MSG)

    _RANGE_MSG = ("tdnet_b200: an activation exceeded the SPLIT16 range (|x| > 6e4) in an earlier frame, its logits "
                  "are invalid; use engine_mode='simt' (fp32 planes) for this checkpoint")

    @torch.no_grad()
    def _subnetwork(self, name, img):
        """`model.pretrainedK(img)` of the reference (the sub-network of path K: dilated ResNet, image -> c4 feature map
        [n, C4, H/8, W/8], resnet.py:204-215) on the engine's kernels.  Does not touch the FIFO."""
        k = int(name[len("pretrained"):] or 1)
        if not img.is_cuda:
            raise RuntimeError("tdnet_b200 runs on a CUDA (sm_100) device only; there is no CPU path. "
                               "Move the model and the input with .to('cuda').")
        if img.dtype != torch.float32 or img.dim() != 4 or img.shape[1] != 3:
            raise RuntimeError("expected an fp32 NCHW image batch [n,3,H,W] (Testing/dataloader.py:69-71)")
        img = img.contiguous()
        n, _, h, w = img.shape
        with torch.cuda.device(img.device):
            fill, key = self._fifo_fill, getattr(self, "_active_key", None)
            eng = self._engine(img, (n, 3, h, w))
            if key is not None and key == self._active_key:
                self._fifo_fill = fill                                 # a sub-network call is not a clip boundary
            plan = eng.backbone_plan(k)
            eng.run(plan, img.data_ptr(), 0, torch.cuda.current_stream(img.device).cuda_stream)
            out = plan.taps["c4"].torch().permute(0, 3, 1, 2).contiguous()
            if eng.tc:
                self._poll_range_flag(eng)
        return out

    def time_attention_op(self, frames, step, reps=8):
        """Average device time (ms) of the fused attention-propagation kernel of the big hop (the last hop of
        the path: all P queries against the P' keys of the newest FIFO entry), timed like time_dominant_op."""
        total = 0.0
        for r in range(reps):
            pos = (step + r) % self.PATHS
            name = self.arch.hop_modules(pos + 1)[-1] + ".attention"
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.forward(frames[(step + r) % len(frames)], pos_id=pos, _probe=(name, e0, e1))
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1)
        return total / reps

    def time_dominant_op(self, frames, step, reps=8):
        """Average device time (ms) of the frame's dominant kernel -- the last 3x3 conv of layer4, 154.6
        GFLOP at 1024x2048 for td4-psp18 -- bracketed by CUDA events while whole frames run."""
        total = 0.0
        for r in range(reps):
            pos = (step + r) % self.PATHS
            self.dominant_op_name = f"pretrained{pos + 1}.layer4.1.conv2"
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.forward(frames[(step + r) % len(frames)], pos_id=pos, _probe=(self.dominant_op_name, e0, e1))
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1)
        return total / reps

    # forward_path1..4 of the reference (td4_psp18.py:137-212, td2_psp50.py:112-143) return the LOW-resolution logits
    # [n, nclass, H/8, W/8] of one path; forward() adds the bilinear upsample (:227).  Same here: the frame runs through
    # the engine (FIFO advanced exactly as by forward()) and the head map is returned as a new NCHW tensor.
    def _forward_path(self, k, img):
        if not (1 <= k <= self.PATHS):
            raise AttributeError(f"{type(self).__name__} has no forward_path{k} (path_num = {self.PATHS})")
        self.forward(img, pos_id=k - 1)
        _, plan = self._last
        return plan.taps["head"].torch().permute(0, 3, 1, 2).contiguous()

    def forward_path1(self, img):
        return self._forward_path(1, img)

    def forward_path2(self, img):
        return self._forward_path(2, img)

    def forward_path3(self, img):
        return self._forward_path(3, img)

    def forward_path4(self, img):
        return self._forward_path(4, img)

    def forward_labels(self, img, pos_id=0):
        """Same frame step as forward(), but returns the uint8 label map [n, H, W] =
        forward(img, pos_id).max(1)[1] (Testing/test.py:61) computed by a fused upsample + arg-max kernel:
        the 159 MB fp32 logits tensor is never written."""
        return self.forward(img, pos_id, _labels=True)

    def forward_preview(self, img, pos_id=0, out_hw=None, u8=False):
        """What Testing/test.py:61-64 keeps of a frame: `output.max(1)[1]` as int8, resized with cv2.INTER_NEAREST to
        (W//4, H//4) -- returned as uint8 [n, H//4, W//4] (or `out_hw`).  Nearest resampling only selects pixels, so the
        final upsample + arg-max runs on those pixels alone (tdn_upsample_argmax_sampled), bit-consistent with
        forward_labels at the sampled positions."""
        return self.forward(img, pos_id, _u8=u8, _preview=out_hw or "quarter")

    def forward_u8(self, frame_u8, pos_id=0, labels=False):
        """Device-side frame ingest: `frame_u8` is the RGB camera frame as uint8 HWC
        [n,H,W,3] on the GPU; (x/255 - mean)/std of Testing/dataloader.py:52-53,66-67 is applied inside the
        stem kernel through an fp64-built table, bit-identical to feeding the normalised fp32 NCHW tensor."""
        return self.forward(frame_u8, pos_id, _labels=labels, _u8=True)

    def check_numeric_range(self):
        """Raise if any SPLIT16 activation ever exceeded the fp16 range guard (|x| > 6e4) since the
        engine was created.  Synchronises; meant for validation runs, not the frame loop."""
        for eng in self._engines.values():
            torch.cuda.synchronize(eng.device)
            if int(eng.range_flag[0]) != 0:
                eng.range_flag.zero_()
                raise RuntimeError(self._RANGE_MSG)

    @torch.no_grad()
    def forward(self, img, pos_id=0, _probe=None, _labels=False, _u8=False, _preview=None):
        if not img.is_cuda:
            raise RuntimeError("tdnet_b200 runs on a CUDA (sm_100) device only; there is no CPU path. "
                               "Move the model and the input with .to('cuda').")
        if _u8:
            if img.dtype != torch.uint8 or img.dim() != 4 or img.shape[3] != 3:
                raise RuntimeError("forward_u8 expects a uint8 HWC frame batch [n,H,W,3]")
            if len(self.arch.stems[1]) != 1:
                raise NotImplementedError("forward_u8 needs the single-conv stem (ResNet-18/34 backbones)")
            shape_nchw = (img.shape[0], 3, img.shape[1], img.shape[2])
        elif img.dtype != torch.float32 or img.dim() != 4 or img.shape[1] != 3:
            raise RuntimeError("expected an fp32 NCHW image batch [n,3,H,W] (Testing/dataloader.py:69-71)")
        if not (0 <= pos_id < self.PATHS):
            raise RuntimeError(f"pos_id must be in [0,{self.PATHS})")
        img = img.contiguous()
        n, _, h, w = shape_nchw if _u8 else img.shape
        # the library launches on the CURRENT device: make the input's device current for the whole frame
        with torch.cuda.device(img.device):
            return self._forward_on_device(img, pos_id, n, h, w, _probe, _labels, _u8, _preview)

    def _forward_on_device(self, img, pos_id, n, h, w, _probe, _labels, _u8, _preview):
        eng = self._engine(img, (n, 3, h, w))
        steady = self._fifo_fill >= self.arch.depth
        plan = eng.plan(pos_id + 1, steady)
        last_op = None
        if _preview is not None:
            ph, pw = (h // 4, w // 4) if _preview == "quarter" else _preview
            out = torch.empty((n, ph, pw), dtype=torch.uint8, device=img.device)
            last_op = eng.preview_op(plan, ph, pw)
        elif _labels:
            out = torch.empty((n, h, w), dtype=torch.uint8, device=img.device)
        else:
            out = torch.empty((n, self.nclass, h, w), dtype=torch.float32, device=img.device)
        plan.uses = getattr(plan, "uses", 0) + 1
        if self.use_cuda_graph and _probe is None and getattr(plan, "graph", None) is not None:
            eng.run_graphed(plan, img.data_ptr(), out.data_ptr(), labels=_labels, u8=_u8, last_op=last_op)
        else:
            eng.run(plan, img.data_ptr(), out.data_ptr(), torch.cuda.current_stream(img.device).cuda_stream, _probe,
                    labels=_labels, u8=_u8, last_op=last_op)
        # FIFO bookkeeping mirrors buffer_contral (slot j = j-th oldest frame once the FIFO is full); the observable
        # lists are materialised on access (Q_queue / K_queue / V_queue properties), not here in the frame loop
        self._fifo_fill = min(self._fifo_fill + 1, self.arch.depth)
        self._fifo_manual = None
        self._last = (eng, plan)
        if eng.tc:
            self._poll_range_flag(eng)
        return out
