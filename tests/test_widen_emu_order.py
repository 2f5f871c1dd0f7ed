"""Order-dependence check of the SIMT kernels: the CPU interpreter re-runs a kernel-parity file with blocks and threads
scheduled in DESCENDING order between barriers (``LNST_EMU_ORDER=reverse``, tools/cpu_emu/cpu_emu.h).  A kernel that
relies on the unspecified execution order -- a missing ``__syncthreads``, one block reading what another wrote -- gives
different numbers there and fails its parity test.  (The whole emulator-backed suite was run this way by hand; this
keeps two representative files in the regular CPU run.)"""
import os
import subprocess
import sys

from conftest import ROOT


def test_kernel_parity_under_reversed_schedule():
    env = dict(os.environ, LNST_EMU_ORDER='reverse')
    r = subprocess.run([sys.executable, '-m', 'pytest', '-q', '-x', '-m', 'not gpu', '-p', 'no:cacheprovider',
                        'tests/test_widen_graphnet_kernels.py', 'tests/test_widen_resim.py', '-k', 'not oracle'],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert ' passed' in r.stdout
