#!/usr/bin/env python
"""Tolerance study: error of the loss history and of the stylised field against the
fp64 oracle as a function of the iteration count, for the three loss-network arithmetic modes of the engine
(conv_math = bf16 | bf16x3 | fp32) and for the fp32 oracle itself (the floor any fp32 implementation sits on).

    python tools/precision_study.py [--iters 50] [--res 24] > profiles/r2_precision_study.json      (on the B200)

The oracle is test infrastructure; this tool is a checker, not a product path."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'neural-flow-style_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=50)
    ap.add_argument('--res', type=int, default=24)
    ap.add_argument('--view-mode', default='allreduce')
    a = ap.parse_args()
    from helpers import smoke_cfg
    from lnst import synth
    from lnst.styler_3p import Styler
    from oracle.styler import Oracle3P
    import oracle.vgg
    checkpoints = sorted({k for k in (1, 3, 10, 20, 35, a.iters) if k <= a.iters})
    kw = dict(res=a.res, rotate=True, n_views=9, view_mode=a.view_mode, style_layer=['conv2_1', 'conv3_1'],
              w_style_layer=[0.5, 0.5])
    p, r = synth.smoke_particles(6000, 2, pad=8)
    sty = synth.style_image(a.res, a.res)
    out = {'config': dict(kw, iters=a.iters, particles=6000), 'checkpoints': checkpoints, 'modes': {}}
    refs = {}
    for k in checkpoints:                         # fp64 oracle = the truth the modes are measured against
        o = Oracle3P(smoke_cfg(iter=k, conv_math='fp32', **kw), oracle.vgg.synthetic_weights(), dtype=torch.float64)
        refs[k] = o.run({'p': p, 'r': r}, style_targets=[sty], view_mode=a.view_mode)

    def errors(run):
        rows = []
        for k in checkpoints:
            res = run(k)
            ref = refs[k]
            l_new, l_ref = np.asarray(res['l'][0], np.float64), np.asarray(ref['l'][0], np.float64)
            d_new, d_ref = np.asarray(res['d'], np.float64), np.asarray(ref['d'], np.float64)
            g_new = np.asarray(res['g_opt'][0], np.float64)
            g_ref = np.asarray(ref['g_opt'][0].numpy() if hasattr(ref['g_opt'][0], 'numpy') else ref['g_opt'][0], np.float64)
            rows.append({'iters': k,
                         'loss_rel_max': float(np.max(np.abs(l_new - l_ref) / np.abs(l_ref))),
                         'loss_rel_last': float(abs(l_new[-1] - l_ref[-1]) / abs(l_ref[-1])),
                         'field_max_rel': float(np.abs(d_new - d_ref).max() / np.abs(d_ref).max()),
                         'var_rel_l2': float(np.linalg.norm(g_new - g_ref) / max(np.linalg.norm(g_ref), 1e-30))})
        return rows

    def engine(math):
        def run(k):
            s = Styler(smoke_cfg(iter=k, conv_math=math, **kw), weights=synth.vgg_weights())
            s.style_img = sty
            return s.run({'p': p, 'r': r})
        return run

    def oracle32(k):
        o = Oracle3P(smoke_cfg(iter=k, conv_math='fp32', **kw), oracle.vgg.synthetic_weights())
        return o.run({'p': p, 'r': r}, style_targets=[sty], view_mode=a.view_mode)

    out['modes']['oracle_fp32'] = errors(oracle32)
    for math in ('fp32', 'bf16x3', 'bf16'):
        out['modes']['engine_' + math] = errors(engine(math))
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
