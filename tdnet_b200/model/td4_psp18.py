"""B200-native stand-in for Testing/model/pspnet/td4_psp18.py (class `td4_psp18`, :29-240)."""
from ._td_base import TDModel


class BatchNorm2d:  # noqa: N801 - name kept for API parity (td4_psp18.py:11-24); folded at prepare time
    def __init__(self, *a, **k):
        raise RuntimeError("BatchNorm is folded into the convolution epilogues; not a standalone module here")


class td4_psp18(TDModel):  # noqa: N801
    ARCH, PATHS = "td4_psp18", 4

    def __init__(self, nclass=21, norm_layer=BatchNorm2d, backbone="resnet18", dilated=True, aux=True,
                 multi_grid=True, path_num=None, model_path=None, ln_shape=(97, 193)):
        super().__init__(nclass, norm_layer, backbone, dilated, aux, multi_grid, path_num, model_path, ln_shape)
