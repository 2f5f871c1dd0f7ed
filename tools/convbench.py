#!/usr/bin/env python
"""Per-layer timing of the tensor-core convolutions at the C3 shapes (developer tool; GPU only)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'neural-flow-style_b200')]
import torch  # noqa: E402
from lnst import _lib, ops  # noqa: E402

dev = torch.device('cuda:0')
lib = _lib.get()


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


n = 9
layers = [('conv1_2', 200, 64, 64), ('conv2_1', 100, 64, 128), ('conv2_2', 100, 128, 128), ('conv3_1', 50, 128, 256),
          ('d conv3_1', 50, 256, 128), ('d conv2_1', 100, 128, 64)]
if len(sys.argv) > 1:
    n = int(sys.argv[1])
for name, hw, cin, cout in layers:
    x = torch.randn(n, hw, hw, cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(9, cout, cin, device=dev) / (3 * cin ** 0.5)).to(torch.bfloat16)
    b = torch.randn(cout, device=dev)
    y = torch.empty(n, hw, hw, cout, device=dev, dtype=torch.bfloat16)
    mask = torch.randn(n, hw, hw, cout, device=dev).to(torch.bfloat16)
    flops = 2.0 * n * hw * hw * 9 * cin * cout
    row = []
    for halo in (0, 1, 2):
        lib.call('lnst_set_conv_halo', halo)
        us = timeit(lambda: ops.conv3x3_bf16_tc(x, w, b, relu=True, y=y))
        usm = timeit(lambda: ops.conv3x3_bf16_tc(x, w, None, relu=False, mask=mask, y=y))
        row.append('halo=%d %7.1f us (%6.1f TF/s)  masked %7.1f us' % (halo, us, flops / us / 1e6, usm))
    print('%-10s %s' % (name, ' | '.join(row)), flush=True)
lib.call('lnst_set_conv_halo', 2)
g = torch.randn(n, 200, 200, 64, device=dev).to(torch.bfloat16)
wd = torch.randn(3, 3, 64, 3, device=dev)
wd16 = torch.zeros(9, 16, 64, dtype=torch.bfloat16, device=dev)
wd16[:, :3] = wd.permute(0, 1, 3, 2).reshape(9, 3, 64).to(torch.bfloat16)
print('conv_first_bwd  cuda-core %7.1f us | tensor-core %7.1f us' % (timeit(lambda: ops.conv_first_bwd(g, wd)),
                                                                    timeit(lambda: ops.conv_first_bwd_tc(g, wd16))))
