"""The C-ABI library loads on a CPU-only box and exports every symbol ``include/lnst_b200.h``
declares (no compute calls here); bad arguments are rejected before any launch."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, PKG
from lnst import _lib


@pytest.fixture(scope='module')
def dll():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'lnst_b200.h')).read()
    return sorted(set(re.findall(r'\bint\s+(lnst_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_the_binding_table():
    names = declared_symbols()
    assert len(names) >= 30
    table = set(_lib.SIGNATURES)
    if any(n in names for n in _lib.CUDA_ONLY):
        table |= set(_lib.CUDA_ONLY)
    assert set(names) == table, set(names) ^ table


def test_every_declared_symbol_is_exported(dll):
    for name in declared_symbols():
        assert hasattr(dll, name), name
    assert dll.lnst_abi_version() == 1


def test_argument_errors_return_negative_without_launching(dll):
    dll.lnst_adam_step.restype = ctypes.c_int
    assert dll.lnst_adam_step(None, None, None, None, ctypes.c_int64(4), ctypes.c_float(0.1), ctypes.c_float(0.9),
                              ctypes.c_float(0.999), ctypes.c_float(1e-8), ctypes.c_float(1.0), None) < 0
    g = _lib.make_grid(3, [4, 4, 4], [4, 4, 4], 1, False)
    g.cell = 0.0
    assert dll.lnst_splat_sph_fwd(None, None, ctypes.c_int64(1), ctypes.byref(g), ctypes.c_float(1.0),
                                  ctypes.c_float(1.0), None, None, 1, ctypes.c_float(1000.0), None, None) < 0


def test_product_has_no_cpu_fallback():
    """Without a GPU the product loader must refuse (it never falls back to the oracle/emu)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    prev = _lib._lib
    _lib.set_for_testing(None)
    try:
        with pytest.raises(_lib.LnstError):
            _lib.get()
    finally:
        _lib.set_for_testing(prev)
    pkg_src = ''
    for f in os.listdir(os.path.join(PKG, 'lnst')):
        if f.endswith('.py'):
            pkg_src += open(os.path.join(PKG, 'lnst', f)).read()
    assert 'import oracle' not in pkg_src and 'from oracle' not in pkg_src and 'cpu_emu' not in pkg_src.replace(
        'tools/cpu_emu', '')
