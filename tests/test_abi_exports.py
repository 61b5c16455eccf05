"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/b200at.h declares
(no compute calls: there is no GPU here)."""
import ctypes
import os
import subprocess

import pytest

import revisiting_at_b200
from revisiting_at_b200 import _abi


@pytest.fixture(scope='module')
def libpath():
    import __graft_entry__ as ge
    ge.build()
    assert os.path.exists(_abi.LIB_PATH)
    return _abi.LIB_PATH


def test_library_exports_every_declared_symbol(libpath):
    L = ctypes.CDLL(libpath)
    names = _abi.exported_symbols()
    assert len(names) >= 7
    for n in names:
        assert hasattr(L, n), f'{n} declared in include/b200at.h but not exported'
    L.b200at_abi_version.restype = ctypes.c_int
    assert L.b200at_abi_version() == _abi.ABI_VERSION


def test_binding_declares_every_symbol(libpath):
    L = _abi.lib()
    for n in _abi.exported_symbols():
        if n == 'b200at_abi_version':
            continue
        assert getattr(L, n).argtypes is not None, f'{n} has no ctypes signature in _abi.py'


def test_sass_is_sm100a(libpath):
    out = subprocess.run(['cuobjdump', '-lelf', libpath], capture_output=True, text=True).stdout
    assert 'sm_100a' in out, out


def test_product_never_imports_oracle_or_hostcheck():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, 'revisiting-at_b200')
    bad = []
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith('.py'):
                src = open(os.path.join(dp, f)).read()
                for needle in ('import oracle', 'from oracle', 'hostcheck'):
                    for line in src.splitlines():
                        if needle in line and not line.lstrip().startswith(('#', '"', "'")) and 'tests/hostcheck' not in line:
                            bad.append((f, line.strip()))
    assert not bad, bad
