import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'neural-flow-style_b200')
for p in (ROOT, PKG, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def emu_lib():
    """The SIMT kernels compiled for the CPU interpreter (tools/cpu_emu) -- test tooling."""
    sys.path.insert(0, os.path.join(ROOT, 'tools', 'cpu_emu'))
    import build_emu
    from lnst import _lib
    return _lib.Lib(build_emu.build(), 'emu')


@pytest.fixture(params=['emu', pytest.param('cuda', marks=pytest.mark.gpu)])
def dev(request):
    """Runs a kernel test twice: through the CPU interpreter here, and on the B200 (-m gpu)."""
    import torch
    from lnst import _lib
    prev = _lib._lib
    if request.param == 'emu':
        _lib.set_for_testing(request.getfixturevalue('emu_lib'))
        device = torch.device('cpu')
    else:
        _lib.set_for_testing(None)
        _lib.get()                      # raises if the CUDA library is missing
        device = torch.device('cuda:0')
    yield device
    if device.type == 'cuda':
        torch.cuda.synchronize()
    _lib.set_for_testing(prev)
