"""BASELINE.json configs[0] at FULL size: dambreak2d single frame, 2-D 256x256 colour field
(base grid 64^2 x scale 4: domain 6.4 x 6.4, nsize 4 = 81 splat taps), VGG-19 conv1_1 style loss, 50 Adam iterations --
the reference's own CPU-runnable case.  Engine on the B200 through the C-ABI against the CPU oracle, same seeded
inputs.  (The CPU-interpreter variant runs 1 iteration -- it needs ~30 s per iteration at this size; 4 iterations were checked by hand.)"""
import os

import numpy as np
import pytest

from helpers import dam_cfg
from lnst import synth


def _c1(iters):
    return dam_cfg(resolution=[256, 256], domain=[6.4, 6.4], radius=0.025, nsize=4, support=4, iter=iters, lr=0.01,
                   octave_n=1, style_layer=['conv1_1'], w_style_layer=[1.0], w_style=1, w_tv=0, style_mask=False,
                   conv_math='fp32')


def test_c1_full_size(dev):
    from lnst.styler_2p import Styler
    from oracle.styler import Oracle2P
    import oracle.vgg
    if dev.type != 'cuda' and not os.environ.get('LNST_SLOW'):
        pytest.skip('~35 s on the CPU interpreter: set LNST_SLOW=1 (the GPU variant always runs)')
    iters = 50 if dev.type == 'cuda' else 1
    cfg = _c1(iters)
    p, r = synth.dam_particles_2d(cfg.domain)
    assert 2500 < p[0].shape[0] < 4500                      # N ~ 3.3k
    sty = synth.style_image(256, 256)
    c0 = np.random.RandomState(5).uniform(0.2, 0.8, (1, p[0].shape[0], 3)).astype(np.float32)
    new = Styler(cfg, weights=synth.vgg_weights(), device=dev)
    new.style_img = sty
    out = new.run({'p': p, 'r': r}, c_init=c0)
    ref = Oracle2P(_c1(iters), oracle.vgg.synthetic_weights()).run({'p': p, 'r': r}, style_targets=[sty], c_init=c0)
    np.testing.assert_allclose(out['l'][0], ref['l'][0], rtol=1e-3)
    # after 50 Adam steps a handful of elements whose gradient sits at the fp32 noise floor may have walked apart
    # (Adam normalises the step size): norm-based bounds, as SURVEY 8c states them (rel-L2 of the variables <= 1e-3
    # per 50 iterations on the fp32 path; 5e-3 here leaves room for the atomics' summation order)
    dd = np.abs(out['d'].astype(int) - ref['d'].astype(int))
    assert dd.mean() < 0.25 and np.percentile(dd, 99.9) <= 2
    c_got, c_ref = np.asarray(out['c'][0], np.float64), np.asarray(ref['c'][0], np.float64)
    assert np.linalg.norm(c_got - c_ref) <= 5e-3 * np.linalg.norm(c_ref)
    if iters > 1:
        assert ref['l'][0][-1] < ref['l'][0][0]             # the optimisation makes progress
