"""Pin the oracle: the reference's one known-answer (5x5 warp tables) plus brute-force
definitions of every restated op on tiny grids, and fp64 gradchecks."""
import math

import numpy as np
import pytest
import torch

import oracle
from oracle import transform as T, render as R, vgg as V, loss as L
from oracle.adam import TFAdam


# ---- reference known-answer test: transform.py:1865-1884 ---------------------------------
def test_warp_kat_identity_and_zoom():
    img = torch.arange(25, dtype=torch.float32).reshape(1, 5, 5, 1)
    g = T.mgrid(5, 5).reshape(1, 2, -1)
    ident = T.interpolate(img, [g[:, 0], g[:, 1]]).reshape(5, 5)
    assert torch.equal(ident, img.reshape(5, 5))
    z = 0.5 * g   # theta = identity*0.5: matrix .5 I, translation 0
    zoom = T.interpolate(img, [z[:, 0], z[:, 1]]).reshape(5, 5).numpy()
    expect = np.array([[6, 6.5, 7, 7.5, 8], [8.5, 9, 9.5, 10, 10.5], [11, 11.5, 12, 12.5, 13],
                       [13.5, 14, 14.5, 15, 15.5], [16, 16.5, 17, 17.5, 18]], np.float32)
    np.testing.assert_allclose(zoom, expect, rtol=0, atol=1e-5)


# ---- splat vs dense O(N*V) definition -----------------------------------------------------
def _dense_splat(p, res, domain, radius, support, nsize, weights=None, mass=1.0):
    """Direct evaluation: every (particle, cell) pair within the (2nsize+1)^3 stencil."""
    D, H, W = res
    cs = domain[0] / res[0]
    h = radius * support
    sigma = 8 / math.pi / h ** 3
    out = np.zeros(res)
    wm = np.zeros(res)
    for n, pt in enumerate(p):
        pd = pt * np.array(domain, np.float64)
        if not np.all((pd >= 0) & (pd < np.array(domain))):
            continue
        idx = np.floor(pd / cs).astype(int)
        for dz in range(-nsize, nsize + 1):
            for dy in range(-nsize, nsize + 1):
                for dx in range(-nsize, nsize + 1):
                    c = idx + np.array([dz, dy, dx])
                    if np.any(c < 0) or np.any(c >= np.array(res)):
                        continue
                    q = np.linalg.norm(pd - (c + 0.5) * cs) / h
                    if q > 1:
                        w = 0.0
                    elif q <= 0.5:
                        w = sigma * (6 * (q ** 3 - q ** 2) + 1)
                    else:
                        w = sigma * 2 * (1 - q) ** 3
                    x = 1.0 if weights is None else weights[n]
                    out[c[0], H - 1 - c[1], c[2]] += mass * w * x
                    wm[c[0], H - 1 - c[1], c[2]] += w
    return out, wm


def test_p2g_matches_dense_definition():
    rng = np.random.RandomState(0)
    res, domain = [6, 8, 7], [3.0, 4.0, 3.5]
    p = rng.uniform(-0.05, 1.05, size=(60, 3))
    radius, support, rho = 0.25, 4, 1000.0
    got = T.p2g(torch.tensor(p[None], dtype=torch.float64), domain, res, radius, rho, 1, is_2d=False,
                clip=False, support=support)[0, ..., 0].numpy()
    mass = 0.8 * (2 * radius) ** 3 * rho
    want, _ = _dense_splat(p, res, domain, radius, support, 1, mass=mass)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-9)


def test_p2g_wavg_matches_dense_definition():
    rng = np.random.RandomState(1)
    res, domain = [6, 6, 6], [6.0, 6.0, 6.0]
    p = rng.uniform(0.1, 0.9, size=(40, 3))
    x = rng.uniform(0, 1, size=(40,))
    got = T.p2g_wavg(torch.tensor(p[None], dtype=torch.float64),
                     torch.tensor(x[None, :, None], dtype=torch.float64), domain, res, 0.5, 1,
                     is_2d=False, clip=False, support=2)[0, ..., 0].numpy()
    num, wm = _dense_splat(p, res, domain, 0.5, 2, 1, weights=x)
    want = np.where(wm > 1e-6, num / np.where(wm > 1e-6, wm, 1), num)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)


def test_p2g_wavg_nan_gradient_trap():
    """transform.py:1703 -- an isolated particle has empty target cells => NaN gradient."""
    p = torch.tensor([[[0.5, 0.5, 0.5]]], dtype=torch.float64)
    x = torch.ones(1, 1, 1, dtype=torch.float64, requires_grad=True)
    out = T.p2g_wavg(p, x, [8, 8, 8], [8, 8, 8], 0.5, 1, is_2d=False, clip=False, support=2)
    out.sum().backward()
    assert torch.isnan(x.grad).all()


def test_p2g_2d_colour_and_flip():
    p = torch.tensor([[[0.26, 0.66]]], dtype=torch.float64)     # (y,x)
    pc = torch.tensor([[[0.2, 0.5, 1.0]]], dtype=torch.float64)
    pd = torch.tensor([[[900.0]]], dtype=torch.float64)
    res, dom = [4, 5], [0.4, 0.5]
    out = T.p2g(p, dom, res, 0.025, 1000.0, 1, pc=pc, pd=pd, is_2d=True, clip=False)[0]
    gray = T.p2g(p, dom, res, 0.025, 1000.0, 1, is_2d=True, clip=False)[0, ..., 0]
    # particle in cell (y=1,x=3) -> after the y flip its own cell is row H-1-1 = 2
    assert gray[2, 3] == gray.max() and gray[2, 3] > 0
    np.testing.assert_allclose(out[..., 1].numpy(), gray.numpy() * 0.5 / 900.0, rtol=1e-12)


def test_p2g_gradcheck_positions():
    rng = np.random.RandomState(2)
    p = torch.tensor(rng.uniform(0.2, 0.8, size=(1, 5, 3)), dtype=torch.float64, requires_grad=True)
    f = lambda q: T.p2g(q, [0.4] * 3, [4, 4, 4], 0.025, 1000.0, 1, is_2d=False, clip=False)
    assert torch.autograd.gradcheck(f, (p,), eps=1e-7, atol=1e-3, rtol=1e-4)


# ---- smooth / render ---------------------------------------------------------------------
def test_smooth3_matches_explicit_stencil():
    rng = np.random.RandomState(3)
    d = rng.randn(1, 5, 6, 4, 1)
    got = R.smooth3(torch.tensor(d), 3)[0, ..., 0].numpy()
    k1 = np.array([1.0, 3.0, 1.0])
    pad = np.pad(d[0, ..., 0], 1)
    want = np.zeros_like(got)
    for a in range(3):
        for b in range(3):
            for c in range(3):
                want += k1[a] * k1[b] * k1[c] / 125.0 * pad[a:a + 5, b:b + 6, c:c + 4]
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)


def test_render_smoke_and_liquid_explicit_loop():
    rng = np.random.RandomState(4)
    d = rng.uniform(0, 2, size=(1, 7, 3, 4, 1))
    tau = 0.3
    img = np.zeros((3, 4))
    for h in range(3):
        for w in range(4):
            for i in range(7):
                img[h, w] += d[0, i, h, w, 0] * math.exp(-tau * d[0, i:, h, w, 0].sum())
    got = R.render(torch.tensor(d), tau, False)[0, ..., 0].numpy()
    np.testing.assert_allclose(got, img / img.max(), rtol=1e-12)
    liq = R.render(torch.tensor(d), tau, True)[0, ..., 0].numpy()
    np.testing.assert_allclose(liq, 1 - np.exp(-tau * d[0, ..., 0].sum(0)), rtol=1e-12)


def test_rotate_identity_and_grid_sample_equivalence():
    rng = np.random.RandomState(5)
    d = torch.tensor(rng.rand(1, 5, 6, 7, 1))
    out = T.rotate(d, [np.identity(3)])
    np.testing.assert_allclose(out.numpy(), d.numpy(), atol=1e-12)
    Rm = np.matmul(T.rot_y_3d(10.0), T.rot_z_3d(-5.0))
    out = T.rotate(d, [Rm])[0, ..., 0]
    # same thing through torch.grid_sample(border, align_corners=True); grid is (x=W,y=H,z=D)
    g = T.mgrid(5, 6, 7, dtype=torch.float64).reshape(3, -1)
    g = torch.tensor(Rm) @ g
    grid = torch.stack([g[2], g[1], g[0]], -1).reshape(1, 5, 6, 7, 3)
    want = torch.nn.functional.grid_sample(d.permute(0, 4, 1, 2, 3), grid, mode='bilinear',
                                           padding_mode='border', align_corners=True)[0, 0]
    np.testing.assert_allclose(out.numpy(), want.numpy(), atol=1e-12)


def test_views_uniform_default_is_nine():
    mats, views = T.rot_mat(-5, 5, 5, -10, 10, 10, sample_type='uniform')
    assert len(mats) == 9
    assert [v['phi'] for v in views][:3] == [-5.0, -5.0, -5.0]
    np.testing.assert_allclose(mats[4], np.identity(3), atol=1e-15)


def test_views_poisson_count_and_range():
    rng = np.random.RandomState(123)
    mats, views = T.rot_mat(-5, 5, 5, -10, 10, 10, sample_type='poisson', rng=rng, nv=9)
    assert len(mats) == 9
    for v in views:
        assert -5 <= v['phi'] <= 5 and -10 <= v['theta'] <= 10


def test_resize_bilinear_legacy_coordinates():
    x = torch.arange(4, dtype=torch.float64).reshape(1, 1, 4, 1)
    y = R.resize_bilinear_legacy(x, 1, 6)[0, 0, :, 0].numpy()
    # src = dst*4/6 ; clamp upper index
    src = np.arange(6) * 4 / 6
    want = np.interp(src, np.arange(4), np.arange(4))
    np.testing.assert_allclose(y, want, atol=1e-12)


def test_advect_zero_velocity_identity_and_shift():
    rng = np.random.RandomState(6)
    d = torch.tensor(rng.rand(1, 6, 5, 1))
    v = torch.zeros(1, 6, 5, 2, dtype=torch.float64)
    np.testing.assert_allclose(T.advect(d, v).numpy(), d.numpy(), atol=1e-12)
    v[..., 0] = 2.0 / 5.0      # one cell along axis 0 in normalised units (len 6 -> 2/(6-1))
    out = T.advect(d, v).numpy()
    np.testing.assert_allclose(out[0, 1:], d.numpy()[0, :-1], atol=1e-12)
    np.testing.assert_allclose(out[0, 0], d.numpy()[0, 0], atol=1e-12)   # border clamp


# ---- loss net / losses / optimiser ---------------------------------------------------------
def test_vgg_endpoints_shapes_and_avgpool():
    w = V.synthetic_weights()
    x = torch.rand(1, 16, 12, 3) * 255
    ep = V.forward(x, w, upto='conv3_1')
    assert ep['conv1_2'].shape == (1, 16, 12, 64)
    assert ep['conv2_1'].shape == (1, 8, 6, 128)
    assert ep['conv3_1'].shape == (1, 4, 3, 256)
    a = ep['conv1_2'][0]
    want = (a[0::2, 0::2] + a[1::2, 0::2] + a[0::2, 1::2] + a[1::2, 1::2]) / 4
    np.testing.assert_allclose(ep['pool1'][0].numpy(), want.numpy(), rtol=1e-5, atol=1e-4)
    assert (ep['conv3_1'] >= 0).all()


def test_gram_and_style_loss_vs_einsum():
    rng = np.random.RandomState(7)
    f = torch.tensor(rng.randn(1, 3, 4, 5))
    fs = torch.tensor(rng.randn(1, 2, 2, 5))
    total, _ = L.style_loss([f], [fs], [0.7], 1)
    G = np.einsum('hwc,hwd->cd', f[0].numpy(), f[0].numpy()) / (2 * 3 * 4 * 5)
    Gs = np.einsum('hwc,hwd->cd', fs[0].numpy(), fs[0].numpy()) / (2 * 2 * 2 * 5)
    np.testing.assert_allclose(float(total), 0.7 * ((G - Gs) ** 2).sum(), rtol=1e-12)


def test_tf_adam_first_steps_closed_form():
    opt = TFAdam()
    var = torch.zeros(3, dtype=torch.float64)
    g = torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)
    v1 = opt.step(var, g, 0.1)
    # t=1: m=(1-b1)g, v=(1-b2)g^2, lr_t = lr*sqrt(1-b2)/(1-b1)
    lr_t = 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9)
    want = -lr_t * (0.1 * g) / (torch.sqrt(0.001 * g * g) + 1e-8)
    np.testing.assert_allclose(v1.numpy(), want.numpy(), rtol=2e-5)  # fp32 beta powers
    v2 = opt.step(v1, g, 0.1)
    assert (v2.abs() > v1.abs()).all()


def test_octave_sizes():
    from oracle.styler import octave_sizes
    assert octave_sizes([200, 300, 200], 2, 1.8) == [[111, 166, 111], [200, 300, 200]]
    assert octave_sizes([512, 1024], 3, 1.7) == [[177, 354], [301, 602], [512, 1024]]
