"""B200-native stand-in for Testing/model/pspnet/pspnet.py (class `pspnet`, :31-100): the single-path
PSPNet comparison model of the paper's speed table (Testing/TEST_README.md:31).

Same constructor kwargs, state-dict layout (`pretrained.*`, `head.conv5.*`; checkpoints load with
strict=True) and `forward(x, pos_id=None)`: only the LAST image of the batch is segmented (`x = x[-1:]`,
pspnet.py:74) and `pos_id` is ignored.  Backbone, pyramid pooling, head and upsample run on the same
sm_100a kernels as the TD paths (tdnet_b200/engine.py: `_build_pspnet_tail`); there is no FIFO.
"""
from ._td_base import TDModel
from .td4_psp18 import BatchNorm2d  # noqa: F401


class pspnet(TDModel):  # noqa: N801
    ARCH, PATHS = "pspnet", 1
    BACKBONES = ("resnet101", "resnet50", "resnet34", "resnet18")

    def __init__(self, nclass=21, norm_layer=BatchNorm2d, backbone="resnet101", dilated=True, aux=True,
                 multi_grid=True, model_path=None):
        super().__init__(nclass, norm_layer, backbone, dilated, aux, multi_grid, 1, model_path, ln_shape=(0, 0))

    def forward(self, x, pos_id=None, **kw):
        return super().forward(x[-1:], 0, **kw)

    def forward_labels(self, x, pos_id=None):
        return super().forward(x[-1:], 0, _labels=True)

    def forward_preview(self, x, pos_id=None, out_hw=None, u8=False):
        return super().forward(x[-1:], 0, _u8=u8, _preview=out_hw or "quarter")

    def forward_u8(self, frame_u8, pos_id=None, labels=False):
        return super().forward(frame_u8[-1:], 0, _labels=labels, _u8=True)

    def set_ln_shape(self, h8, w8):
        raise RuntimeError("pspnet has no LayerNorm")
