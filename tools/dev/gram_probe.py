#!/usr/bin/env python
"""Gram difference of split features at C3's two loss layers: TMEM-summed products against the split-row Gram (GPU only)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in (ROOT, os.path.join(ROOT, 'neural-flow-style_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, _p)
import torch  # noqa: E402
from lnst import _lib, ops  # noqa: E402

dev = torch.device('cuda:0')
lib = _lib.get()


def t(fn, reps=20):
    """device time per call: `reps` calls captured in one CUDA graph (eager launches are CPU-bound here)"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


for (n, h, w, ch) in ((9, 100, 100, 128), (9, 50, 50, 256)):
    F = ops.to_split(torch.relu(torch.randn(n, h, w, ch, device=dev)))
    Gs = torch.zeros(ch, ch, device=dev)
    loss = torch.zeros(n, device=dev)
    for mode in (1, 0, 1):
        lib.call('lnst_set_gram_split3', mode)
        print((n, h, w, ch), 'mode', mode, '%.1f us' % t(lambda: ops.gram_diff_bf16x3_tc(F, 2.0 * h * w * ch, Gs, 0.5, loss)), flush=True)

# Gram gradient (per-pixel GEMM with a per-image matrix, addend + ReLU mask in the epilogue)
for (n, h, w, ch) in ((9, 100, 100, 128), (9, 50, 50, 256)):
    F = ops.to_split(torch.relu(torch.randn(n, h, w, ch, device=dev)))
    Gs = torch.zeros(ch, ch, device=dev)
    loss = torch.zeros(n, device=dev)
    G, Gd2 = ops.gram_diff_bf16x3_tc(F, 2.0 * h * w * ch, Gs, 0.5, loss)
    add = ops.to_split(torch.randn(n, h, w, ch, device=dev))
    print((n, h, w, ch), 'gram_bwd x3 (addend, mask) %.1f us' % t(lambda: ops.gram_bwd_bf16x3_tc(F, Gd2, 0.3, add, 1, add)), flush=True)
    print((n, h, w, ch), 'gram_bwd x3 (no addend)    %.1f us' % t(lambda: ops.gram_bwd_bf16x3_tc(F, Gd2, 0.3, None, 1)), flush=True)
