"""Golden vectors produced by THE REFERENCE'S OWN CODE (tests/golden/ref_*.npz).

``tests/golden/make_reference_golden.py`` executes the unmodified ``/root/reference`` modules
(``styler_3p.Styler.run``, ``styler_2p.Styler.run``, ``transform``, ``vgg``, ``styler_base``) on top of
``oracle/tfshim`` -- a stand-in for the absent third-party TensorFlow 1.15 -- and stores what
``Styler(config).run(params)`` returns.  These tests hold

  * the CPU oracle (``-m "not gpu"``): pins the checker itself against the reference, and
  * the CUDA path through the C-ABI (``-m gpu``), plus the kernel sources under the CPU interpreter,

to those numbers, on inputs re-generated from seeds (neither the reference nor the shim is needed
at test time).  Tolerances (fp32 path): loss rel 2e-4, field 2e-4 of its max, uint8 images +-1,
velocities / colours 2e-3 of their max.  The oracle is held to loss 2e-5, field 1e-4 (measured: 2e-6 .. 5e-5; Adam's first steps amplify fp32 rounding).
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, 'golden')
sys.path.insert(0, GOLD)
import make_reference_golden as M  # noqa: E402

CASES_3D = ['density_noview', 'density_sequential', 'density_resize_tv_content', 'density_octaves_poisson',
            'density_sequence', 'density_interp_both', 'density_reg_content_image', 'position_clip_vgg16',
            'position_liquid', 'position_smoke_views']
CASES_2D = ['colour_2d', 'colour_2d_mask', 'colour_2d_frames', 'colour_2d_batch', 'colour_2d_mask_on_ref']


def _ref(name):
    return dict(np.load(os.path.join(GOLD, 'ref_%s.npz' % name)))


def _style_targets(cfg):
    from lnst import synth
    if not cfg.w_style:
        return None
    res = cfg.resolution
    hw = [int(int(s) * cfg.resize_scale) for s in res[-2:]] if not np.isclose(cfg.resize_scale, 1) else list(res[-2:])
    return [synth.style_image(hw[0], hw[1])] * cfg.octave_n


def _content_targets(cfg):
    from lnst import synth
    if not getattr(cfg, 'content_image', False):
        return None
    return [synth.style_image(int(cfg.resolution[-2]), int(cfg.resolution[-1]), seed=11)] * cfg.octave_n


def _check(out, want, kind, ltol=2e-4, ftol=2e-4):
    l_got = np.asarray([np.asarray(o, np.float64) for o in out['l']])
    np.testing.assert_allclose(l_got, want['l'], rtol=ltol)
    d = np.asarray(out['d'])
    assert d.shape == want['d'].shape
    if kind == '2c':
        assert np.abs(d.astype(int) - want['d'].astype(int)).max() <= 1
        c = np.asarray([np.asarray(x) for x in out['c']])
        assert np.abs(c - want['c']).max() <= 2e-3 * np.abs(want['c']).max()
    else:
        assert np.abs(d - want['d']).max() <= ftol * np.abs(want['d']).max()
        r = np.asarray(out['r'])
        assert r.shape == want['r'].shape
        assert np.abs(r.astype(int) - want['r'].astype(int)).max() <= 1
        if 'v' in want:
            v = np.asarray([np.asarray(x) for x in out['v']])
            assert np.abs(v - want['v']).max() <= 2e-3 * np.abs(want['v']).max()
    for k in want:
        if k.startswith('d_intm'):
            got = np.asarray(out['d_intm'][int(k[6:])])
            assert got.shape == want[k].shape
            assert np.abs(got.astype(int) - want[k].astype(int)).max() <= 1


# ---- the oracle against the reference --------------------------------------------------------------
@pytest.mark.parametrize('name', CASES_3D)
def test_oracle_matches_reference_run_3d(name):
    import oracle.vgg
    from oracle.styler import Oracle3P
    cfg, params = M.case_inputs(name)
    model = 'vgg_16' if '16' in cfg.network else 'vgg_19'
    out = Oracle3P(cfg, oracle.vgg.synthetic_weights(model)).run(params, style_targets=_style_targets(cfg),
                                                                content_targets=_content_targets(cfg),
                                                                view_mode='sequential')
    _check(out, _ref(name), M.CASES[name][0], ltol=2e-5, ftol=1e-4)


@pytest.mark.parametrize('name', CASES_2D)
def test_oracle_matches_reference_run_2d(name):
    import oracle.vgg
    from oracle.styler import Oracle2P
    cfg, params = M.case_inputs(name)
    out = Oracle2P(cfg, oracle.vgg.synthetic_weights()).run(params, style_targets=_style_targets(cfg))
    _check(out, _ref(name), '2c', ltol=2e-5)


# ---- the engine (CPU interpreter of the kernel sources here, CUDA through the C-ABI on the B200) ---
@pytest.mark.parametrize('name', CASES_3D)
def test_engine_matches_reference_run_3d(name, dev):
    from lnst import synth
    from lnst.styler_3p import Styler
    cfg, params = M.case_inputs(name)
    cfg.conv_math = 'fp32'
    cfg.view_mode = 'sequential'
    st = Styler(cfg, weights=synth.vgg_weights('vgg_16' if '16' in cfg.network else 'vgg_19'))
    tg = _style_targets(cfg)
    if tg is not None:
        st.style_img = tg[0]
    ct = _content_targets(cfg)
    if ct is not None:
        st.content_img = ct[0]
    out = st.run(params)
    _check(out, _ref(name), M.CASES[name][0])


@pytest.mark.parametrize('name', CASES_2D)
def test_engine_matches_reference_run_2d(name, dev):
    from lnst import synth
    from lnst.styler_2p import Styler
    cfg, params = M.case_inputs(name)
    cfg.conv_math = 'fp32'
    st = Styler(cfg, weights=synth.vgg_weights())
    tg = _style_targets(cfg)
    if tg is not None:
        st.style_img = tg[0]
    out = st.run(params)
    _check(out, _ref(name), '2c')


# ---- operator level: transform.py functions run under the TF stand-in vs the oracle's restatement --
def _close(got, want, tol=2e-6, what=''):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert np.array_equal(np.isnan(got), np.isnan(want)), '%s: NaN pattern differs' % what
    ok = ~np.isnan(want)
    scale = np.abs(want[ok]).max() if ok.any() else 0.0
    err = np.abs(got[ok] - want[ok]).max() if ok.any() else 0.0
    assert err <= tol * max(scale, 1e-30), '%s: max err %.3e vs scale %.3e' % (what, err, scale)


def test_oracle_operators_match_reference_transform():
    import torch
    from oracle import transform as T
    I = M.ops_inputs()
    ref = dict(np.load(os.path.join(GOLD, 'ref_ops.npz')))
    q = torch.tensor(I['q'])
    _close(T.cubic_w(q, 0.37, True), ref['W3'], what='W 3-D')
    _close(T.cubic_w(q, 0.37, False), ref['W2'], what='W 2-D')

    res, dom = [7, 9, 8], [0.7, 0.9, 0.8]
    for clip in (False, True):
        p = torch.tensor(I['p3'][None], requires_grad=True)
        y = T.p2g(p, dom, res, 0.025, 1000.0, 1, is_2d=False, clip=clip, support=4)
        (y[0, ..., 0] * torch.tensor(I['g3'])).sum().backward()
        _close(y.detach(), ref['p2g3_clip%d' % clip], what='p2g clip=%s' % clip)
        _close(p.grad, ref['p2g3_clip%d_grad' % clip], tol=1e-5, what='p2g grad clip=%s' % clip)

    res2, dom2 = [10, 14], [1.0, 1.4]
    pc = torch.tensor(I['pc'][None], requires_grad=True)
    y = T.p2g(torch.tensor(I['p2'][None]), dom2, res2, 0.025, 1000.0, 2, pc=pc, pd=torch.tensor(I['pd'][None]),
              is_2d=True, clip=False, support=4)
    (y[0] * torch.tensor(I['g2'])).sum().backward()
    _close(y.detach(), ref['p2g2_colour'], what='p2g colour')
    _close(pc.grad, ref['p2g2_colour_grad'], tol=1e-5, what='p2g colour grad')
    _close(T.p2g(torch.tensor(I['p2'][None]), dom2, res2, 0.025, 1000.0, 2, is_2d=True, clip=False),
           ref['p2g2_gray'], what='p2g 2-D gray')

    for k, support in enumerate((4.0, 2.0)):
        x = torch.tensor(I['xw'][None], requires_grad=True)
        y = T.p2g_wavg(torch.tensor(I['pw'][None]), x, [8, 8, 8], [8, 8, 8], 0.5, 1, is_2d=False, clip=False,
                       support=support)
        (y[0, ..., 0] * torch.tensor(I['gw'])).sum().backward()
        _close(y.detach(), ref['wavg_s%d' % k], what='p2g_wavg')
        assert np.isnan(ref['wavg_s%d_grad' % k]).any()                 # the NaN rule is exercised
        _close(x.grad, ref['wavg_s%d_grad' % k], tol=1e-5, what='p2g_wavg grad')

    mats = [m for m in ref['rotate_mats']]
    _close(T.rotate(torch.tensor(I['vol'][None, ..., None]), mats), ref['rotate'], what='rotate')
    _close(T.advect(torch.tensor(I['adv2_d'][None]), torch.tensor(I['adv2_v'][None]), is_3d=False), ref['advect2'],
           what='advect 2-D')
    _close(T.advect(torch.tensor(I['adv3_d'][None]), torch.tensor(I['adv3_v'][None]), is_3d=True), ref['advect3'],
           what='advect 3-D')

    for st in ('uniform', 'poisson', 'both'):
        rng = np.random.RandomState(123)
        for rep in range(2):
            m, _ = T.rot_mat(-5, 5, 5, -10, 10, 10, sample_type=st, rng=rng, nv=9)
            np.testing.assert_array_equal(np.asarray(m), ref['views_%s_%d' % (st, rep)])


def test_engine_view_sampling_matches_reference():
    """lnst.transform (the product's host-side view sampler) draws the same matrices in the same RNG order."""
    from lnst import transform as LT
    ref = dict(np.load(os.path.join(GOLD, 'ref_ops.npz')))
    for st in ('uniform', 'poisson', 'both'):
        rng = np.random.RandomState(123)
        for rep in range(2):
            m, _ = LT.rot_mat(-5, 5, 5, -10, 10, 10, sample_type=st, rng=rng, nv=9)
            np.testing.assert_allclose(np.asarray(m), ref['views_%s_%d' % (st, rep)], rtol=0, atol=1e-15)
