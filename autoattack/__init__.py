"""Drop-in for `from autoattack import AutoAttack` (AA_eval.py:17): the APGD-CE / APGD-T part of autoattack-0.1
on the B200 attack kernels.  Put this repository's root before any installed `autoattack` on PYTHONPATH."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
import revisiting_at_b200 as _pkg  # noqa: E402,F401
from revisiting_at_b200.autoattack import APGDAttack, APGDAttack_targeted, AutoAttack  # noqa: E402,F401

__all__ = ['AutoAttack', 'APGDAttack', 'APGDAttack_targeted']
