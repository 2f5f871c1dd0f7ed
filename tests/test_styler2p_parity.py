"""2-D colour mode (BASELINE config 1 shape, scaled down): ``lnst.styler_2p.Styler.run`` against
``oracle.styler.Oracle2P.run``.  Runs through the CPU interpreter here and on the B200 (-m gpu)."""
import numpy as np
import pytest

from helpers import make_cfg
from lnst import synth
from lnst.styler_2p import Styler
from oracle.styler import Oracle2P
import oracle.vgg


def cfg2d(**over):
    res = [24, 32]
    cell = 0.1
    base = dict(target_field='c', resolution=res, domain=[r * cell for r in res], radius=0.025, nsize=2,
                support=4, rest_density=1000, lr=0.01, iter=3, octave_n=1, window_sigma=0, frames_per_opt=1,
                style_layer=['conv1_1'], w_style_layer=[1.0], w_style=1, w_tv=0, conv_math='fp32')
    base.update(over)
    return make_cfg(**base)


def run_pair(kw, nf=1):
    c = cfg2d(num_frames=nf, **kw)
    p, r = synth.dam_particles_2d(c.domain, spacing=0.05, num_frames=nf)
    sty = synth.style_image(*c.resolution)
    new = Styler(cfg2d(num_frames=nf, **kw), weights=synth.vgg_weights())
    new.style_img = sty
    c_init = new.init_colors(p[0].shape[0])
    out_new = new.run({'p': p, 'r': r}, c_init=c_init)
    ref = Oracle2P(cfg2d(num_frames=nf, **kw), oracle.vgg.synthetic_weights())
    from lnst.util import octave_sizes, resize
    sizes = octave_sizes(c.resolution, c.octave_n, c.octave_scale)
    out_ref = ref.run({'p': p, 'r': r}, style_targets=[resize(sty, s) for s in sizes], c_init=c_init)
    return out_new, out_ref


def check(out_new, out_ref):
    for a, b in zip(out_new['l'], out_ref['l']):
        np.testing.assert_allclose(a, b, rtol=2e-4)
    for a, b in zip(out_new['g_opt'], out_ref['g_opt']):
        assert np.linalg.norm(a - b.numpy()) / np.linalg.norm(b.numpy()) < 2e-3
    for a, b in zip(out_new['c'], out_ref['c']):
        np.testing.assert_allclose(a, b, atol=2e-3)
    assert out_new['d'].shape == out_ref['d'].shape
    assert np.abs(out_new['d'].astype(int) - out_ref['d'].astype(int)).max() <= 1


def test_colour_mode_single_frame(dev):
    out_new, out_ref = run_pair(dict(w_tv=0.01, style_layer=['conv1_1', 'conv2_1'], w_style_layer=[0.5, 0.5]))
    check(out_new, out_ref)
    assert out_new['d'].shape == (1, 24, 32, 3) and out_new['d'].dtype == np.uint8


def test_colour_mode_frames_octaves(dev):
    out_new, out_ref = run_pair(dict(octave_n=2, octave_scale=1.5, window_sigma=1.0, iter=2), nf=3)
    check(out_new, out_ref)
    assert len(out_new['d_intm']) == 1 and out_new['d_intm'][0].shape == out_ref['d_intm'][0].shape


def test_colour_mode_style_mask(dev):
    """The dambreak2d driver's setting (test_dambreak2d.py:189): Gram of the feature times the density mask
    (bicubic-resized to the feature size), normalised by the mask area (styler_base.py:165-169)."""
    out_new, out_ref = run_pair(dict(style_mask=True, style_layer=['conv1_1', 'conv2_1'], w_style_layer=[0.5, 0.5]))
    check(out_new, out_ref)
    plain, _ = run_pair(dict(style_mask=False, style_layer=['conv1_1', 'conv2_1'], w_style_layer=[0.5, 0.5]))
    assert abs(plain['l'][0][0] - out_new['l'][0][0]) > 1e-3 * abs(plain['l'][0][0])      # the mask matters
