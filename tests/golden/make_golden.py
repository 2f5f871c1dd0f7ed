#!/usr/bin/env python
"""Regenerates the fixtures in this directory.

    python tests/golden/make_golden.py

The reference (TensorFlow 1.15) cannot run in this image, so the vectors come from two places:
  * ``warp_kat.json``   -- the ONLY known-answer the reference holds for this path: the 5x5 identity /
    zoom-in bilinear warp tables of the ``__main__`` docstring, ``transform.py:1865-1884``
    (transcribed; the oracle's ``interpolate`` must reproduce them).
  * ``*.npz``           -- outputs of the CPU oracle (``oracle/``, the line-by-line restatement of the
    reference) on small seeded scenes: they pin the oracle against drift and give the CUDA path
    fixed numbers to hit on the GPU box, where neither the reference nor this script's inputs exist.
Seeds, configurations and sizes are the ones of tests/test_styler_parity.py.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, 'neural-flow-style_b200'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

CASES = {
    # name: (kind, config overrides, particles)
    'density_sequential': ('3d', dict(res=12, iter=3, rotate=True, n_views=9, view_mode='sequential', conv_math='fp32',
                                      style_layer=['conv1_2', 'conv2_1'], w_style_layer=[0.5, 0.5]), 900),
    'density_allreduce': ('3d', dict(res=12, iter=3, rotate=True, n_views=9, view_mode='allreduce', conv_math='fp32',
                                     style_layer=['conv1_2', 'conv2_1'], w_style_layer=[0.5, 0.5]), 900),
    'density_tc_shapes': ('3d', dict(res=20, iter=3, rotate=True, n_views=9, view_mode='allreduce', conv_math='fp32',
                                     style_layer=['conv2_1', 'conv3_1'], w_style_layer=[0.5, 0.5]), 4000),
    'position_liquid': ('3p', dict(res=12, iter=3, conv_math='fp32', w_pressure=0.5, style_layer=['conv1_2'],
                                   w_style_layer=[1.0]), 700),
}


def run_case(name):
    """Oracle result of one case as a dict of numpy arrays (also used by tests/test_golden.py)."""
    from helpers import smoke_cfg, liquid_cfg
    from lnst import synth
    from oracle.styler import Oracle3P
    import oracle.vgg
    kind, kw, n = CASES[name]
    torch.manual_seed(0)
    sty = synth.style_image(kw['res'], kw['res'])
    if kind == '3d':
        p, r = synth.smoke_particles(n, 2, pad=4)
        out = Oracle3P(smoke_cfg(**kw), oracle.vgg.synthetic_weights()).run(
            {'p': p, 'r': r}, style_targets=[sty], view_mode=kw['view_mode'])
    else:
        p = synth.liquid_particles(n)
        out = Oracle3P(liquid_cfg(**kw), oracle.vgg.synthetic_weights()).run({'p': p}, style_targets=[sty])
    return {'l': np.asarray(out['l'], np.float64), 'g_opt': out['g_opt'][0].numpy(), 'd': np.asarray(out['d']),
            'r': np.asarray(out['r'])}


WARP_KAT = {
    'source': 'transform.py:1865-1884 (docstring of __main__)',
    'image': np.arange(25).reshape(5, 5).tolist(),
    'identity': np.arange(25).reshape(5, 5).tolist(),
    'zoom_in': [[6, 6.5, 7, 7.5, 8], [8.5, 9, 9.5, 10, 10.5], [11, 11.5, 12, 12.5, 13],
                [13.5, 14, 14.5, 15, 15.5], [16, 16.5, 17, 17.5, 18]],
}


def main():
    with open(os.path.join(HERE, 'warp_kat.json'), 'w') as f:
        json.dump(WARP_KAT, f, indent=1)
    for name in CASES:
        out = run_case(name)
        # the field is stored as float16-free full precision but only its occupied part matters; keep it whole
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
        print(name, 'loss', out['l'].ravel()[:3], 'bytes', os.path.getsize(os.path.join(HERE, name + '.npz')))


if __name__ == '__main__':
    main()
