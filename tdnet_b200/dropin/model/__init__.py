"""Drop-in replacement for the reference package `Testing/model` (Testing/model/__init__.py:1-3).

Put this directory's parent (tdnet_b200/dropin) ahead of Testing/ on sys.path / PYTHONPATH and the
unmodified Testing/test.py runs on the B200-native kernels:

    PYTHONPATH=/path/to/repo/tdnet_b200/dropin:/path/to/repo python Testing/test.py --model td4-psp18 ...

`from model import td4_psp18, td2_psp50, pspnet` then resolves to the modules below; the classes keep
the reference's constructor, state-dict layout and forward(img, pos_id) signature.
"""
import os as _os
import sys as _sys

_repo = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
if _repo not in _sys.path:
    _sys.path.insert(0, _repo)

from tdnet_b200.model import pspnet, td2_psp50, td4_psp18  # noqa: E402,F401
