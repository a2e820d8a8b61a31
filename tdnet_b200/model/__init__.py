"""Same surface as Testing/model/__init__.py:1-3: the *modules* td4_psp18, td2_psp50 and pspnet
(classes are model.td4_psp18.td4_psp18 / model.td2_psp50.td2_psp50 / model.pspnet.pspnet), plus td2_fa, the
stand-in for Training/ptsemseg/models/td2_fanet/td2_fa.py (class model.td2_fa.td2_fa)."""
from . import pspnet, td2_fa, td2_psp50, td4_psp18  # noqa: F401
