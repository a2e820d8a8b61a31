"""B200-native stand-in for Testing/model/pspnet/td2_psp50.py (class `td2_psp50`, :29-166)."""
from ._td_base import TDModel
from .td4_psp18 import BatchNorm2d  # noqa: F401


class td2_psp50(TDModel):  # noqa: N801
    ARCH, PATHS = "td2_psp50", 2

    def __init__(self, nclass=21, norm_layer=BatchNorm2d, backbone="resnet50", dilated=True, aux=True,
                 multi_grid=True, path_num=None, model_path=None, ln_shape=(97, 193)):
        super().__init__(nclass, norm_layer, backbone, dilated, aux, multi_grid, path_num, model_path, ln_shape)
