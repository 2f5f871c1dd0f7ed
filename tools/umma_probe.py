#!/usr/bin/env python
"""Which smem-descriptor addressing lets tcgen05.mma read the 9 taps of a 3x3 convolution out of ONE
halo'd TMA patch?  (developer probe; GPU only; kernel umma_probe_k in csrc/conv_tc.cu)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'neural-flow-style_b200')]
import torch  # noqa: E402
from lnst import _lib  # noqa: E402
from lnst._lib import ptr  # noqa: E402

dev = torch.device('cuda:0')
lib = _lib.get()
g = torch.Generator().manual_seed(0)
x = torch.randint(-4, 5, (180, 64), generator=g).to(torch.bfloat16)
b = torch.randint(-4, 5, (16, 64), generator=g).to(torch.bfloat16)
xd, bd = x.to(dev), b.to(dev)
patch = x.float().reshape(18, 10, 64)
for pitched in (0, 1):
    for base_mode in (0, 1, 2):
        errs = []
        for ky in range(3):
            for kx in range(3):
                out = torch.full((128, 16), -777.0, device=dev)
                lib.call('lnst_umma_probe', ptr(xd), ptr(bd), ptr(out), pitched, ky, kx, base_mode,
                         _lib.stream_ptr(dev))
                torch.cuda.synchronize()
                a = patch[ky:ky + 16, kx:kx + 8].reshape(128, 64)
                want = a @ b.float().t()
                errs.append(float((out.cpu() - want).abs().max()))
        print('pitched=%d base_mode=%d  max err per tap (ky,kx row-major): %s' % (pitched, base_mode,
              ' '.join('%.0f' % e for e in errs)), flush=True)
