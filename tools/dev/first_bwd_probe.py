#!/usr/bin/env python
"""conv1_1 gray data gradient: one-GEMM-per-patch kernel against the per-tap halo kernel at C3's shape (GPU only)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in (ROOT, os.path.join(ROOT, 'neural-flow-style_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, _p)
import torch  # noqa: E402
from lnst import _lib, ops, synth, vgg  # noqa: E402

dev = torch.device('cuda:0')
lib = _lib.get()
net = vgg.LossNet(synth.vgg_weights(), 'vgg_19', dev, math='bf16x3')
g = ops.to_split(torch.randn(9, 200, 200, 64, device=dev))


def t(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


for mode in (1, 0, 1):
    lib.call('lnst_set_conv_first_col', mode)
    print('mode', mode, '%.1f us' % t(lambda: ops.conv_first_bwd_gray_x3_tc(g, net.tc.wd16_gray)), flush=True)
