"""B200-native stand-in for Training/ptsemseg/models/td2_fanet/td2_fa.py (class `td2_fa`, :16-280), inference only
(SURVEY.md 8f rank 4): the authors' lightweight TD2 variant -- two non-dilated ResNet sub-networks with fast-attention
(linear attention) FPN decoders, one attention-propagation hop from the previous frame, LayerNorm, FPN head.

Kept from the reference: constructor kwargs (`nclass, backbone, norm_layer, loss_fn, path_num, mdl_path, teacher`) and
asserts, the state-dict layout (`pretrained{1,2}.*`, `ffm_{32,16,8,4}_{1,2}.*`, `enc*`, `atn*`, `layer_norm*.ln.*`,
`head*`, `head_aux*`; loads with strict=True), `pretrained_init()` from a single-path FANet checkpoint
(ptsemseg/utils.py:35-66 `split_fanet_dict`), and `forward(f_img, lbl=None, pos_id=None)` where `f_img` is the pair
[previous frame, current frame] (td2_fa.py:88-89) and the result the fp32 logits [n, nclass, H, W] of the current
frame.  The reference keeps no state between calls: both sub-networks run on every call.

Not carried over: the training branch (loss, knowledge distillation against `teacher`) -- calling the module in
training mode raises -- and the `pdb.set_trace()` of the reference constructor (:81).
"""
from __future__ import annotations

import os
from collections import OrderedDict

import torch
import torch.nn as nn

from . import arch as A
from ._td_base import TDModel, _attach, _default_init_


class td2_fa(TDModel):  # noqa: N801
    ARCH, PATHS = "td2_fa", 2

    def __init__(self, nclass=21, backbone="resnet18", norm_layer=None, loss_fn=None, path_num=None, mdl_path=None,
                 teacher=None, ln_shape=(96, 192)):
        nn.Module.__init__(self)
        assert backbone == "resnet50" or backbone == "resnet34" or backbone == "resnet18"   # td2_fa.py:37
        assert path_num == 2                                                                # td2_fa.py:38
        self.loss_fn, self.fa_path, self.path_num = loss_fn, mdl_path, path_num
        self.norm_layer, self.nclass, self.backbone = norm_layer, nclass, backbone
        self.teacher = teacher
        self.arch = A.build_arch("td2_fa", backbone, nclass)
        self.expansion = self.arch.c_exp
        self.ln_shape = tuple(ln_shape)
        for key, (shape, kind) in A.parameter_table(self.arch, self.ln_shape).items():
            _attach(self, key, shape, kind)
        _default_init_(self)
        self._engines = {}
        self._fifo_fill, self._fifo_manual = 0, None             # no FIFO in this model; the shared base reports empty lists
        self.engine_mode = os.environ.get("TDNET_B200_ENGINE", "tc")
        self.use_cuda_graph = os.environ.get("TDNET_B200_CUDA_GRAPH", "1") != "0"
        self.pretrained_init()

    def pretrained_init(self):
        """td2_fa.py:246-274: initialise BOTH sub-networks from one single-path FANet checkpoint whose keys start with
        resnet. / ffm_32. / ffm_16. / ffm_8. / ffm_4. / clslayer_8. / clslayer_32. (ptsemseg/utils.py:35-66)."""
        if self.fa_path is None:
            return
        if not os.path.isfile(self.fa_path):
            print("No pretrained found at '{}'".format(self.fa_path))
            return
        print("Initializaing sub networks with pretrained '{}'".format(self.fa_path))
        model_state = torch.load(self.fa_path, map_location="cpu")
        groups = {"resnet": "pretrained{}", "ffm_32": "ffm_32_{}", "ffm_16": "ffm_16_{}", "ffm_8": "ffm_8_{}",
                  "ffm_4": "ffm_4_{}", "clslayer_8": "head{}", "clslayer_32": "head_aux{}"}
        own = self.state_dict()
        new = OrderedDict()
        for k, v in model_state.items():
            head, _, rest = k.partition(".")
            if head in groups:
                for idx in (1, 2):
                    new[groups[head].format(idx) + "." + rest] = v
        for prefix in {g.format(i) for g in groups.values() for i in (1, 2)}:     # strict per sub-module, as the reference
            want = {k for k in own if k.startswith(prefix + ".")}
            have = {k for k in new if k.startswith(prefix + ".")}
            if want != have:
                raise RuntimeError("Error(s) in loading state_dict for {}: missing {} unexpected {}".format(
                    prefix, sorted(want - have)[:4], sorted(have - want)[:4]))
        own.update(new)
        self.load_state_dict(own, strict=True)

    def reset(self):
        pass

    def forward_path1(self, f_img):
        """td2_fa.py:87-131: sub-network 1 on the current frame; full-resolution logits (the reference interpolates
        inside forward_path1/2 for this model)."""
        return self.forward(f_img, pos_id=0)

    def forward_path2(self, f_img):
        return self.forward(f_img, pos_id=1)

    def forward_path3(self, f_img):
        raise AttributeError("td2_fa has two paths")

    forward_path4 = forward_path3

    def forward_labels(self, f_img, pos_id=0):
        """uint8 label map [n, H, W] = forward(f_img, pos_id=pos_id).max(1)[1], fused upsample + arg-max."""
        return self.forward(f_img, pos_id=pos_id, _labels=True)

    def forward_preview(self, f_img, pos_id=0, out_hw=None):
        """uint8 [n, H//4, W//4]: the arg-max labels resized with cv2.INTER_NEAREST, only the sampled pixels computed."""
        return self.forward(f_img, pos_id=pos_id, _preview=out_hw or "quarter")

    def forward_u8(self, *a, **k):
        raise NotImplementedError("forward_u8 needs the fused ReLU stem; the FANet stem ends in LeakyReLU")

    @torch.no_grad()
    def forward(self, f2_img, lbl=None, pos_id=None, _probe=None, _labels=False, _preview=None):
        if self.training:
            raise RuntimeError("tdnet_b200.td2_fa implements the inference path only: call .eval() first "
                               "(the training branch, td2_fa.py:116-129, is out of scope)")
        if pos_id not in (0, 1):
            raise RuntimeError("Only Two Paths.")                                   # td2_fa.py:207
        prev, cur = f2_img[0], f2_img[1]
        for img in (prev, cur):
            if not img.is_cuda:
                raise RuntimeError("tdnet_b200 runs on a CUDA (sm_100) device only; there is no CPU path. "
                                   "Move the model and the input with .to('cuda').")
            if img.dtype != torch.float32 or img.dim() != 4 or img.shape[1] != 3:
                raise RuntimeError("expected fp32 NCHW image batches [n,3,H,W]")
        if prev.shape != cur.shape or prev.device != cur.device:
            raise RuntimeError("the two frames of the pair must have the same shape and device")
        prev, cur = prev.contiguous(), cur.contiguous()
        n, _, h, w = cur.shape
        with torch.cuda.device(cur.device):      # the library launches on the current device
            return self._forward_on_device(prev, cur, pos_id, n, h, w, _probe, _labels, _preview)

    def _forward_on_device(self, prev, cur, pos_id, n, h, w, _probe, _labels, _preview):
        eng = self._engine(cur, (n, 3, h, w))
        plan = eng.plan(pos_id + 1, True)
        last_op = None
        if _preview is not None:
            ph, pw = (h // 4, w // 4) if _preview == "quarter" else _preview
            out = torch.empty((n, ph, pw), dtype=torch.uint8, device=cur.device)
            last_op = eng.preview_op(plan, ph, pw)
        elif _labels:
            out = torch.empty((n, h, w), dtype=torch.uint8, device=cur.device)
        else:
            out = torch.empty((n, self.nclass, h, w), dtype=torch.float32, device=cur.device)
        plan.uses = getattr(plan, "uses", 0) + 1
        if self.use_cuda_graph and _probe is None and plan.uses > 1:
            eng.run_graphed(plan, cur.data_ptr(), out.data_ptr(), labels=_labels, img2_ptr=prev.data_ptr(),
                            last_op=last_op)
        else:
            eng.run(plan, cur.data_ptr(), out.data_ptr(), torch.cuda.current_stream(cur.device).cuda_stream, _probe,
                    labels=_labels, img2_ptr=prev.data_ptr(), last_op=last_op)
        self._last = (eng, plan)
        if eng.tc:
            self._poll_range_flag(eng)
        return out
