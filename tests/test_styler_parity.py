"""Loop-level parity: ``lnst.styler_3p.Styler.run`` against ``oracle.styler.Oracle3P.run`` on the
same seeded scene, weights and style target -- loss history, final variables, stylised field and
render.  Runs through the CPU interpreter here and on the B200 with ``-m gpu``.

Stated tolerances (fp32 path; atomics reorder sums, Adam amplifies round-off of tiny gradients):
  loss history  rel <= 2e-4;  stylised field d_out  max-abs <= 2e-4 * max|d|;
  variables after K Adam iterations  rel-L2 <= 2e-3.
"""
import numpy as np
import pytest
import torch

from helpers import smoke_cfg, liquid_cfg
from lnst import synth
from lnst.styler_3p import Styler
from oracle.styler import Oracle3P
import oracle.vgg


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def check_run(res_new, res_ref, tol_loss=2e-4, tol_var=2e-3, tol_field=2e-4):
    for lo_new, lo_ref in zip(res_new['l'], res_ref['l']):
        np.testing.assert_allclose(lo_new, lo_ref, rtol=tol_loss)
    for g_new, g_ref in zip(res_new['g_opt'], res_ref['g_opt']):
        g_ref = g_ref.numpy()
        assert np.array_equal(g_new == 0, g_ref == 0) or rel_l2(g_new, g_ref) < tol_var
        assert rel_l2(g_new, g_ref) < tol_var, rel_l2(g_new, g_ref)
    d_new, d_ref = res_new['d'], res_ref['d']
    assert d_new.shape == d_ref.shape
    assert np.abs(d_new - d_ref).max() <= tol_field * np.abs(d_ref).max()
    assert res_new['r'].shape == res_ref['r'].shape
    assert np.abs(res_new['r'].astype(int) - res_ref['r'].astype(int)).max() <= 1


def scene(res, n, nf=1):
    p, r = synth.smoke_particles(n, 2, pad=4, num_frames=nf)
    return {'p': p, 'r': r}


@pytest.mark.parametrize('view_mode', ['sequential', 'allreduce'])
def test_density_mode_multiview(dev, view_mode):
    res = 12
    kw = dict(res=res, iter=3, rotate=True, n_views=9, view_mode=view_mode, conv_math='fp32',
              style_layer=['conv1_2', 'conv2_1'], w_style_layer=[0.5, 0.5])
    params = scene(res, 900)
    sty = synth.style_image(res, res)
    W = synth.vgg_weights()
    new = Styler(smoke_cfg(**kw), weights=W)
    new.style_img = sty
    out_new = new.run(params)
    ref = Oracle3P(smoke_cfg(**kw), oracle.vgg.synthetic_weights())
    out_ref = ref.run(params, style_targets=[sty], view_mode=view_mode)
    check_run(out_new, out_ref)
    g = out_new['g_opt'][0]
    assert (g == 0).mean() > 0.02 and (g != 0).mean() > 0.3      # NaN-rule particles frozen, others moved


def test_density_mode_single_view_octaves_resize_tv(dev):
    res = 16
    kw = dict(res=res, iter=2, rotate=False, conv_math='fp32', octave_n=2, octave_scale=1.6, resize_scale=1.5,
              w_tv=0.01, lr_scale=2.0, style_layer=['conv2_1'], w_style_layer=[1.0])
    params = scene(res, 1500)
    W = synth.vgg_weights()
    new = Styler(smoke_cfg(**kw), weights=W)
    from lnst.util import octave_sizes
    sizes = octave_sizes([res] * 3, 2, 1.6)
    sty = synth.style_image(24, 24)
    new.style_img = sty
    out_new = new.run(params)
    ref = Oracle3P(smoke_cfg(**kw), oracle.vgg.synthetic_weights())
    from lnst.util import resize
    targets = [resize(sty, (int(np.float32(s[1]) * np.float32(1.5)), int(np.float32(s[2]) * np.float32(1.5)))) for s in sizes]
    out_ref = ref.run(params, style_targets=targets)
    check_run(out_new, out_ref)
    assert len(out_new['d_intm']) == 1 and out_new['d_intm'][0].shape == out_ref['d_intm'][0].shape


def test_position_mode_liquid(dev):
    res = 12
    kw = dict(res=res, iter=3, conv_math='fp32', w_pressure=0.5, style_layer=['conv1_2'], w_style_layer=[1.0])
    p = synth.liquid_particles(700)
    sty = synth.style_image(res, res)
    new = Styler(liquid_cfg(**kw), weights=synth.vgg_weights())
    new.style_img = sty
    out_new = new.run({'p': p})
    ref = Oracle3P(liquid_cfg(**kw), oracle.vgg.synthetic_weights())
    out_ref = ref.run({'p': p}, style_targets=[sty])
    check_run(out_new, out_ref, tol_var=5e-3)
    assert out_new['v'] is not None and np.abs(out_new['v'][0]).max() > 0


def test_multi_frame_temporal_smoothing(dev):
    res = 10
    kw = dict(res=res, iter=2, conv_math='fp32', num_frames=4, window_sigma=1.5, frames_per_opt=2,
              style_layer=['conv1_2'], w_style_layer=[1.0])
    params = scene(res, 500, nf=4)
    sty = synth.style_image(res, res)
    new = Styler(smoke_cfg(**kw), weights=synth.vgg_weights())
    new.style_img = sty
    out_new = new.run(params)
    ref = Oracle3P(smoke_cfg(**kw), oracle.vgg.synthetic_weights())
    out_ref = ref.run(params, style_targets=[sty])
    check_run(out_new, out_ref)
    assert out_new['d'].shape[0] == 4


def test_content_target_image_and_style(dev):
    """Content loss against a target IMAGE's feature (styler_base.py:137-141, 233-247) together with the
    style loss, single view, density mode."""
    res = 12
    kw = dict(res=res, iter=3, rotate=False, conv_math='fp32', w_content=0.7, w_content_amp=1.5, top_k=0,
              content_layer='conv2_1', style_layer=['conv1_2'], w_style_layer=[1.0])   # top_k > 0 is for class logits only
    params = scene(res, 900)
    sty = synth.style_image(res, res)
    con = synth.style_image(res, res, seed=11)
    new = Styler(smoke_cfg(**kw), weights=synth.vgg_weights())
    new.style_img, new.content_img = sty, con
    out_new = new.run(params)
    ref = Oracle3P(smoke_cfg(**kw), oracle.vgg.synthetic_weights())
    out_ref = ref.run(params, style_targets=[sty], content_targets=[con])
    check_run(out_new, out_ref)
    off = Oracle3P(smoke_cfg(**dict(kw, w_content=0)), oracle.vgg.synthetic_weights()).run(params, style_targets=[sty])
    assert abs(off['l'][0][0] - out_ref['l'][0][0]) > 1e-3 * abs(out_ref['l'][0][0])   # the content term matters


def test_edge_all_padding_and_out_of_domain_particles(dev):
    """Empty input in the reference's sense: every row is padding (p = -1, test_smokegun.py:48) or lies
    outside the domain -- nothing is splatted, the render is empty, the loss is the style target's own Gram
    energy and the variables stay zero (the NaN gradients of empty cells are swallowed, styler_3p.py:360)."""
    res = 10
    kw = dict(res=res, iter=2, rotate=False, conv_math='fp32', style_layer=['conv1_2'], w_style_layer=[1.0],
              render_liquid=True)          # liquid render: no division by an all-zero image maximum
    p = [np.concatenate([-np.ones((40, 3), np.float32), np.full((10, 3), 1.5, np.float32)])]
    r = [np.random.RandomState(3).rand(50, 2).astype(np.float32)]
    sty = synth.style_image(res, res)
    new = Styler(smoke_cfg(**kw), weights=synth.vgg_weights())
    new.style_img = sty
    out_new = new.run({'p': p, 'r': r})
    ref = Oracle3P(smoke_cfg(**kw), oracle.vgg.synthetic_weights()).run({'p': p, 'r': r}, style_targets=[sty])
    np.testing.assert_allclose(out_new['l'][0], ref['l'][0], rtol=2e-4)
    assert np.abs(out_new['d']).max() == 0.0 and np.abs(ref['d']).max() == 0.0
    assert np.abs(out_new['g_opt'][0]).max() == 0.0 and np.abs(ref['g_opt'][0].numpy()).max() == 0.0


def test_edge_ragged_frames_padded_like_the_drivers(dev):
    """Frames with different particle counts, padded to the longest with p = -1 rows as the drivers do
    (test_smokegun.py:60-65): the padding must not contribute and must not move."""
    res = 10
    kw = dict(res=res, iter=2, conv_math='fp32', num_frames=3, window_sigma=1.0, frames_per_opt=1,
              style_layer=['conv1_2'], w_style_layer=[1.0])
    p, r = synth.smoke_particles(300, 2, pad=0, num_frames=3)
    for t, keep in enumerate((300, 220, 150)):                   # ragged: frame t really has `keep` particles
        p[t] = p[t].copy()
        p[t][keep:] = -1.0
    sty = synth.style_image(res, res)
    new = Styler(smoke_cfg(**kw), weights=synth.vgg_weights())
    new.style_img = sty
    out_new = new.run({'p': p, 'r': r})
    ref = Oracle3P(smoke_cfg(**kw), oracle.vgg.synthetic_weights()).run({'p': p, 'r': r}, style_targets=[sty])
    check_run(out_new, ref)
    # padding rows never get a gradient; only the temporal filter can leak neighbours' updates into them
    assert np.abs(out_new['g_opt'][0][300:]).max(initial=0.0) == 0.0
    assert np.allclose(out_new['p'][2][150:], -1.0)
