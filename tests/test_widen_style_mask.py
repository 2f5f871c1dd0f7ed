"""Style mask in the 3-D styler (reference styler_base.py:165-169 with d_gray = the normalised render): the mask
depends on the optimised density, so the loss gradient also flows through the mask (bicubic resize backward) and
through the masked area in the Gram denominator.  Kernel parity + the reference's own run (ref_density_style_mask)."""
import os
import sys

import numpy as np
import pytest
import torch

from lnst import ops
from oracle import loss as L
from test_kernel_parity import close

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
sys.path.insert(0, GOLD)


@pytest.mark.parametrize('H,W,oh,ow', [(12, 12, 6, 6), (9, 14, 4, 7), (5, 6, 10, 9)])
def test_resize_bicubic_bwd(dev, H, W, oh, ow):
    rng = np.random.RandomState(H + ow)
    x = torch.tensor(rng.rand(2, H, W, 1).astype(np.float32), requires_grad=True)
    y = L.bicubic_legacy(x, oh, ow)
    close(ops.resize_bicubic_fwd(x.detach().to(dev), oh, ow), y, tol=2e-6, what='bicubic fwd')
    g = torch.tensor(rng.randn(*y.shape).astype(np.float32))
    (y * g).sum().backward()
    close(ops.resize_bicubic_bwd(g.to(dev), H, W), x.grad, tol=3e-6, what='bicubic bwd')


def test_rowdot(dev):
    rng = np.random.RandomState(0)
    a, b = rng.randn(37, 70).astype(np.float32), rng.randn(37, 70).astype(np.float32)
    s = torch.tensor([2.5])
    out = torch.ones(37).to(dev)
    ops.rowdot(torch.tensor(a).to(dev), torch.tensor(b).to(dev), out, scalar=s.to(dev), scale=-0.5, accumulate=True)
    np.testing.assert_allclose(out.cpu().numpy(), 1 + (a * b).sum(1) - 1.25, rtol=2e-5, atol=2e-5)
    ops.rowdot(torch.tensor(a).to(dev), torch.tensor(b).to(dev), out)
    np.testing.assert_allclose(out.cpu().numpy(), (a * b).sum(1), rtol=2e-5, atol=2e-5)


def test_engine_matches_reference_style_mask_run(dev):
    import make_reference_golden as M
    import test_reference_golden as TR
    from lnst import synth
    from lnst.styler_3p import Styler
    name = 'density_style_mask'
    cfg, params = M.case_inputs(name)
    cfg.conv_math, cfg.view_mode = 'fp32', 'sequential'
    st = Styler(cfg, weights=synth.vgg_weights(), device=dev)
    st.style_img = TR._style_targets(cfg)[0]
    TR._check(st.run(params), dict(np.load(os.path.join(GOLD, 'ref_%s.npz' % name))), '3d')


def test_oracle_matches_reference_style_mask_run():
    import make_reference_golden as M
    import test_reference_golden as TR
    import oracle.vgg
    from oracle.styler import Oracle3P
    name = 'density_style_mask'
    cfg, params = M.case_inputs(name)
    out = Oracle3P(cfg, oracle.vgg.synthetic_weights()).run(params, style_targets=TR._style_targets(cfg),
                                                          view_mode='sequential')
    TR._check(out, dict(np.load(os.path.join(GOLD, 'ref_%s.npz' % name))), '3d', ltol=2e-5, ftol=1e-4)


def test_style_mask_with_resize_scale_matches_oracle(dev):
    """mask from the un-resized render, net input resized (styler_base.py:35-38 after styler_3p.py:161)"""
    from helpers import smoke_cfg
    from lnst import synth
    from lnst.styler_3p import Styler
    from oracle.styler import Oracle3P
    import oracle.vgg
    kw = dict(res=12, iter=2, rotate=False, style_mask=True, resize_scale=1.5, conv_math='fp32',
              style_layer=['conv1_2', 'conv2_1'], w_style_layer=[0.5, 0.5])
    p, r = synth.smoke_particles(700, 2)
    sty = synth.style_image(18, 18)
    new = Styler(smoke_cfg(**kw), weights=synth.vgg_weights(), device=dev)
    new.style_img = sty
    out = new.run({'p': p, 'r': r})
    ref = Oracle3P(smoke_cfg(**kw), oracle.vgg.synthetic_weights()).run({'p': p, 'r': r}, style_targets=[sty])
    np.testing.assert_allclose(out['l'][0], ref['l'][0], rtol=3e-4)
    assert np.abs(out['d'] - ref['d']).max() <= 3e-4 * np.abs(ref['d']).max()


def test_transport_between_frames(dev):
    """StylerBase._transport (reference styler_base.py:59-74): repeated order-1 advection forward / backward in time,
    recursive and single-step, 2-D and 3-D, against the oracle's advect."""
    from helpers import smoke_cfg
    from lnst import synth
    from lnst.styler_3p import Styler
    from oracle import transform as T
    st = Styler(smoke_cfg(res=8, conv_math='fp32'), weights=synth.vgg_weights(), device=dev)
    rng = np.random.RandomState(2)
    for shape in ([9, 11], [6, 7, 5]):
        dim = len(shape)
        g = rng.rand(*shape, 2).astype(np.float32)
        v = rng.uniform(-0.2, 0.2, [4] + shape + [dim]).astype(np.float32)

        def ref(a, b, recursive):
            x = torch.tensor(g)[None]
            adv = lambda f, u: T.advect(f, torch.tensor(u)[None], is_3d=dim == 3)
            if a < b:
                if recursive:
                    for i in range(a, b):
                        x = adv(x, v[i])
                else:
                    x = adv(x, v[a] * (b - a))
            elif a > b:
                if recursive:
                    for i in reversed(range(b, a)):
                        x = adv(x, -v[i])
                else:
                    x = adv(x, -v[a - 1] * (a - b))
            return x[0].numpy()

        for a, b in ((0, 3), (3, 1), (2, 2)):
            for rec in (True, False):
                np.testing.assert_allclose(st._transport(g, v, a, b, recursive=rec), ref(a, b, rec), rtol=0, atol=3e-6)


# ---- v_batch > 1 (config.py:69; styler_3p.py:329-352): groups of views per Adam step -----------------------------
def test_engine_matches_reference_vbatch_run(dev):
    import make_reference_golden as M
    import test_reference_golden as TR
    from lnst import synth
    from lnst.styler_3p import Styler
    name = 'density_vbatch'
    cfg, params = M.case_inputs(name)
    cfg.conv_math, cfg.view_mode = 'fp32', 'sequential'
    st = Styler(cfg, weights=synth.vgg_weights(), device=dev)
    st.style_img = TR._style_targets(cfg)[0]
    TR._check(st.run(params), dict(np.load(os.path.join(GOLD, 'ref_%s.npz' % name))), '3d')


def test_oracle_matches_reference_vbatch_run():
    import make_reference_golden as M
    import test_reference_golden as TR
    import oracle.vgg
    from oracle.styler import Oracle3P
    name = 'density_vbatch'
    cfg, params = M.case_inputs(name)
    out = Oracle3P(cfg, oracle.vgg.synthetic_weights()).run(params, style_targets=TR._style_targets(cfg),
                                                          view_mode='sequential')
    TR._check(out, dict(np.load(os.path.join(GOLD, 'ref_%s.npz' % name))), '3d', ltol=2e-5, ftol=1e-4)


# ---- batch_size > 1 in the 3-D styler, rotate off (styler_3p.py:42,304-363,409-431) ----------------------------------
@pytest.mark.parametrize('name', ['density_batch', 'position_batch_pressure'])
def test_engine_matches_reference_batch_run_3d(dev, name):
    import make_reference_golden as M
    import test_reference_golden as TR
    from lnst import synth
    from lnst.styler_3p import Styler
    cfg, params = M.case_inputs(name)
    cfg.conv_math, cfg.view_mode = 'fp32', 'sequential'
    st = Styler(cfg, weights=synth.vgg_weights(), device=dev)
    st.style_img = TR._style_targets(cfg)[0]
    TR._check(st.run(params), dict(np.load(os.path.join(GOLD, 'ref_%s.npz' % name))), M.CASES[name][0])


@pytest.mark.parametrize('name', ['density_batch', 'position_batch_pressure'])
def test_oracle_matches_reference_batch_run_3d(name):
    import make_reference_golden as M
    import test_reference_golden as TR
    import oracle.vgg
    from oracle.styler import Oracle3P
    cfg, params = M.case_inputs(name)
    out = Oracle3P(cfg, oracle.vgg.synthetic_weights()).run(params, style_targets=TR._style_targets(cfg),
                                                          view_mode='sequential')
    TR._check(out, dict(np.load(os.path.join(GOLD, 'ref_%s.npz' % name))), M.CASES[name][0], ltol=2e-5, ftol=1e-4)


def test_batch_run_3d_octaves_matches_oracle(dev):
    """batches over two octaves: intermediate octave renders with the batch's joint normalisation (interp > 1 cannot be
    combined with batch_size > 1 in the reference: one needs an odd, the other an even frame count)"""
    from helpers import smoke_cfg
    from lnst import synth
    from lnst.styler_3p import Styler
    from oracle.styler import Oracle3P
    import oracle.vgg
    kw = dict(res=12, iter=2, rotate=False, num_frames=4, batch_size=2, frames_per_opt=4, window_sigma=0.7,
              octave_n=2, octave_scale=1.5, w_style=0, w_content=1.0, content_layer='conv1_2', content_channel=3,
              conv_math='fp32')
    p, r = synth.smoke_particles(500, 2, num_frames=4)
    new = Styler(smoke_cfg(**kw), weights=synth.vgg_weights(), device=dev)
    out = new.run({'p': p, 'r': r})
    ref = Oracle3P(smoke_cfg(**kw), oracle.vgg.synthetic_weights()).run({'p': p, 'r': r})
    np.testing.assert_allclose(out['l'][0], ref['l'][0], rtol=3e-4)
    np.testing.assert_allclose(out['l'][1], ref['l'][1], rtol=3e-4)
    rd = np.asarray(ref['d'])
    rd = rd.reshape(out['d'].shape)
    assert np.abs(out['d'] - rd).max() <= 3e-4 * np.abs(rd).max()
    assert out['d_intm'][0].shape == ref['d_intm'][0].shape
    assert np.abs(out['d_intm'][0].astype(int) - ref['d_intm'][0].astype(int)).max() <= 1
    assert np.abs(out['r'].astype(int) - ref['r'].astype(int)).max() <= 1


def test_engine_matches_reference_style_mask_on_ref_run(dev):
    """styler_base.py:171-173 in 3-D: the style feature is masked by the render as well, so the target Gram moves with
    the optimised density and the mask's gradient has a style-side term."""
    import make_reference_golden as M
    import test_reference_golden as TR
    from lnst import synth
    from lnst.styler_3p import Styler
    name = 'density_style_mask_on_ref'
    cfg, params = M.case_inputs(name)
    cfg.conv_math, cfg.view_mode = 'fp32', 'sequential'
    st = Styler(cfg, weights=synth.vgg_weights(), device=dev)
    st.style_img = TR._style_targets(cfg)[0]
    TR._check(st.run(params), dict(np.load(os.path.join(GOLD, 'ref_%s.npz' % name))), '3d')


def test_style_mask_denominators_stay_on_the_device(dev):
    """3-D style mask: the Gram denominators 2 C * area(mask) are device tensors (no read-back per step), the step may
    replay from a CUDA graph, and the results equal the host-denominator path."""
    from helpers import smoke_cfg
    from lnst import synth
    from lnst.styler_3p import Styler
    kw = dict(res=12, iter=4, rotate=False, style_mask=True, conv_math='fp32', style_layer=['conv1_2', 'conv2_1'],
              w_style_layer=[0.5, 0.5])
    p, r = synth.smoke_particles(700, 2)
    sty = synth.style_image(12, 12)
    outs = []
    for on_device in (True, False):
        st = Styler(smoke_cfg(**kw), weights=synth.vgg_weights(), device=dev)
        assert getattr(st, 'cuda_graphs', True)                  # round 1 switched graphs off for style masks
        st.mask_areas_on_device = on_device
        if not on_device:
            st.cuda_graphs = False                               # the host path reads the areas back every step
        st.style_img = sty
        masks = st.style_masks_for(torch.rand(1, 12, 12, device=dev), (12, 12), device_areas=on_device)
        assert torch.is_tensor(masks['conv1_2'][1]) == on_device
        outs.append(st.run({'p': p, 'r': r}))
    np.testing.assert_allclose(outs[0]['l'][0], outs[1]['l'][0], rtol=2e-6)
    assert np.abs(outs[0]['d'] - outs[1]['d']).max() <= 2e-6 * np.abs(outs[1]['d']).max()
