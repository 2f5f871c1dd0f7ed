"""Kernel-level parity: every C-ABI entry point against the oracle on the same seeded inputs.

Each test runs twice -- through the CPU interpreter of the kernel sources here (``dev=emu``)
and on the B200 through the real library (``dev=cuda``, ``-m gpu``).  Tolerances are fp32
round-off scale (atomics reorder sums): max-abs <= 1e-5 * max|ref| unless stated.
"""
import numpy as np
import pytest
import torch

from lnst import _lib, ops
from oracle import transform as T, render as R, vgg as V, loss as L
from oracle.adam import TFAdam


def close(got, want, tol=1e-5, what=''):
    got = got.detach().cpu().double()
    want = want.detach().cpu().double()
    assert got.shape == want.shape, (got.shape, want.shape)
    nan_g, nan_w = torch.isnan(got), torch.isnan(want)
    assert torch.equal(nan_g, nan_w), '%s NaN pattern differs' % what
    scale = want[~nan_w].abs().max().item() if (~nan_w).any() else 0.0
    err = (got[~nan_w] - want[~nan_w]).abs().max().item() if (~nan_w).any() else 0.0
    assert err <= tol * max(scale, 1e-30), '%s max err %.3e vs scale %.3e' % (what, err, scale)


def particles(n, seed, lo=-0.02, hi=1.02, pad=3):
    rng = np.random.RandomState(seed)
    p = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    p[:pad] = -1.0           # the drivers' padding rows (test_smokegun.py:48)
    return torch.tensor(p)


# ---- splats ---------------------------------------------------------------------------------
@pytest.mark.parametrize('res,nsize', [([8, 10, 9], 1), ([6, 6, 6], 2)])
def test_splat_sph_fwd_bwd(dev, res, nsize):
    domain = [r * 0.1 for r in res]
    p = particles(300, 0)
    disp = torch.tensor(np.random.RandomState(1).uniform(-0.01, 0.01, (300, 3)).astype(np.float32))
    radius, support, rho = 0.025, 4, 1000.0
    grid = _lib.make_grid(3, res, domain, nsize, False)
    h = radius * support
    scale = 0.8 * (2 * radius) ** 3 * rho / rho
    out = ops.splat_sph_fwd(p.to(dev), disp.to(dev), grid, h, scale)
    pv = (p + disp).clone().requires_grad_(True)
    want = T.p2g(pv[None], domain, res, radius, rho, nsize, is_2d=False, clip=False, support=support) / rho
    close(out, want[0, ..., 0], what='p2g fwd')
    g_out = torch.tensor(np.random.RandomState(2).randn(*res).astype(np.float32))
    (want[0, ..., 0] * g_out).sum().backward()
    g_p = ops.splat_sph_bwd_pos(p.to(dev), disp.to(dev), grid, h, scale, g_out.to(dev))
    close(g_p, pv.grad, tol=2e-5, what='p2g bwd')


def test_splat_sph_clip_mode(dev):
    res, domain = [6, 6, 6], [0.6, 0.6, 0.6]
    p = particles(100, 3, lo=-0.1, hi=1.1, pad=0)
    grid = _lib.make_grid(3, res, domain, 1, True)
    out = ops.splat_sph_fwd(p.to(dev), None, grid, 0.1, 1.0)
    want = T.p2g(p[None], domain, res, 0.025, 1.0, 1, is_2d=False, clip=True, support=4) / (0.8 * 0.05 ** 3)
    close(out, want[0, ..., 0], what='p2g clip')


def test_splat_2d_color(dev):
    res, domain = [12, 16], [1.2, 1.6]
    rng = np.random.RandomState(4)
    p = torch.tensor(rng.uniform(0.05, 0.95, (150, 2)).astype(np.float32))
    pc = torch.tensor(rng.uniform(0, 1, (150, 3)).astype(np.float32), requires_grad=True)
    pd = torch.tensor((1000 * (1 + 0.02 * rng.randn(150, 1))).astype(np.float32))
    radius, rho, nsize = 0.025, 1000.0, 2
    grid = _lib.make_grid(2, res, domain, nsize, False)
    scale = 0.8 * (2 * radius) ** 2 * rho
    h = radius * 4
    out = ops.splat_sph_fwd(p.to(dev), None, grid, h, scale, pc=pc.detach().to(dev), pd=pd.to(dev), rest_density=rho)
    want = T.p2g(p[None], domain, res, radius, rho, nsize, pc=pc[None], pd=pd[None], is_2d=True, clip=False)
    close(out, want[0], what='p2g colour')
    gray = ops.splat_sph_fwd(p.to(dev), None, grid, h, scale / rho)
    close(gray, T.p2g(p[None], domain, res, radius, rho, nsize, is_2d=True, clip=False)[0, ..., 0] / rho, what='2d gray')
    g_out = torch.tensor(rng.randn(*res, 3).astype(np.float32))
    (want[0] * g_out).sum().backward()
    g_pc = ops.splat_sph_bwd_color(p.to(dev), grid, h, scale, pd.to(dev), 3, rho, g_out.to(dev))
    close(g_pc, pc.grad, what='colour bwd')


def test_splat_wavg_fwd_bwd_with_nan_rule(dev):
    res = [10, 10, 10]
    domain = [10, 10, 10]
    rng = np.random.RandomState(5)
    # a dense blob (no empty target cells inside) plus isolated particles (NaN-gradient rule)
    blob = rng.uniform(0.35, 0.65, (1500, 3))
    lone = np.array([[0.1, 0.1, 0.1], [0.9, 0.12, 0.5], [-1, -1, -1]])
    p = torch.tensor(np.concatenate([blob, lone]).astype(np.float32))
    n = p.shape[0]
    r = torch.tensor(np.stack([rng.uniform(0.2, 1, n), rng.uniform(-0.1, 0.1, n)], -1).astype(np.float32))
    var = torch.tensor(rng.uniform(-1.3, 1.3, (n, 2)).astype(np.float32), requires_grad=True)
    hs = [0.5 * 4, 0.5 * 4 / 2]
    grid = _lib.make_grid(3, res, domain, 1, False)
    wmap = ops.splat_wavg_wmap(p.to(dev), grid, hs)
    num = torch.empty_like(wmap)
    out = torch.empty(res, dtype=torch.float32, device=dev)
    ops.splat_wavg_fwd(p.to(dev), r.to(dev), var.detach().to(dev), grid, hs, wmap, num, out)
    x = r + torch.clamp(var, -1, 1)
    want = 0
    for k in range(2):
        want = want + T.p2g_wavg(p[None], x[None, :, k:k + 1], domain, res, 0.5, 1, is_2d=False, clip=False,
                                 support=4 / 2 ** k)
    close(out, want[0, ..., 0], what='wavg fwd')
    g_out = torch.tensor(rng.randn(*res).astype(np.float32))
    (want[0, ..., 0] * g_out).sum().backward()
    g_var = torch.empty(n, 2, dtype=torch.float32, device=dev)
    ops.splat_wavg_bwd(p.to(dev), var.detach().to(dev), grid, hs, wmap, g_out.to(dev), g_var)
    assert torch.isnan(var.grad).any() and not torch.isnan(var.grad).all()
    close(g_var, var.grad, tol=2e-5, what='wavg bwd')
    # density-mode fast path: unrolled 27-cell stencil + precomputed d out/d num (same NaN rule)
    g_fast = torch.empty(n, 2, dtype=torch.float32, device=dev)
    ops.splat_wavg_bwd_coef(p.to(dev), var.detach().to(dev), grid, hs, ops.splat_wavg_coef(wmap), g_out.to(dev), g_fast)
    close(g_fast, var.grad, tol=2e-5, what='wavg bwd (coef)')


# ---- field ----------------------------------------------------------------------------------
@pytest.mark.parametrize('shape,k', [((7, 9, 11), 3), ((5, 4, 6), 0), ((3, 3, 5), 1)])
def test_smooth3_relu_fwd_bwd(dev, shape, k):
    rng = np.random.RandomState(6)
    d = rng.randn(*shape).astype(np.float32)
    d[0] = 0.0                                   # exact zeros: maximum() passes the gradient there
    d = torch.tensor(d, requires_grad=True)
    want = R.field_post(d[None, ..., None], k)[0, ..., 0]
    out = torch.empty(shape, dtype=torch.float32, device=dev)
    ops.smooth3_relu_fwd(d.detach().to(dev), out, k)
    close(out, want, what='smooth fwd')
    g = torch.tensor(rng.randn(*shape).astype(np.float32))
    (want * g).sum().backward()
    g_in = torch.empty(shape, dtype=torch.float32, device=dev)
    ops.smooth3_relu_bwd(g.to(dev), out, g_in, k)
    close(g_in, d.grad, what='smooth bwd')


# ---- rotate / render ------------------------------------------------------------------------
def _rots():
    mats, _ = T.rot_mat(-5, 5, 5, -10, 10, 10, sample_type='uniform')
    return mats


def test_rotate_fwd(dev):
    rng = np.random.RandomState(7)
    vol = torch.tensor(rng.rand(9, 8, 10).astype(np.float32))
    mats = _rots()[:3] + [np.matmul(T.rot_y_3d(40.0), T.rot_z_3d(25.0))]
    rot = torch.tensor(np.asarray(mats), dtype=torch.float32).reshape(-1, 9)
    out = ops.rotate_fwd(vol.to(dev), rot.to(dev))
    want = T.rotate(vol[None, ..., None], mats)[..., 0]
    close(out, want, what='rotate')


@pytest.mark.parametrize('liquid', [False, True])
@pytest.mark.parametrize('rotated', [False, True, 'merge-off', 'wide'])
def test_raymarch_fwd_bwd(dev, liquid, rotated):
    """Fused rotate + render against transform.rotate + the cumsum render of the oracle.  'wide': rows
    longer than a warp and not a multiple of it (lanes of one warp straddle rows), large and small
    angles, a ray with zero incoming gradient; 'merge-off': the plain eight-atomics backward."""
    from lnst import _lib
    rng = np.random.RandomState(8)
    D, H, W = (11, 6, 45) if rotated == 'wide' else (9, 7, 8)
    vol = torch.tensor((rng.rand(D, H, W) * (rng.rand(D, H, W) > 0.3)).astype(np.float32), requires_grad=True)
    tau = 0.2
    mats = (_rots()[:2] + [np.matmul(T.rot_y_3d(33.0), T.rot_z_3d(-21.0))]) if rotated else None
    if rotated == 'wide':
        mats = mats + [np.identity(3), np.matmul(T.rot_y_3d(-80.0), T.rot_z_3d(100.0))]
    nv = len(mats) if rotated else 1
    rot = torch.tensor(np.asarray(mats), dtype=torch.float32).reshape(-1, 9).to(dev) if rotated else None
    img = torch.empty(nv, H, W, dtype=torch.float32, device=dev)
    stot = torch.empty_like(img)
    ops.raymarch_fwd(vol.detach().to(dev), rot, tau, liquid, img, stot)
    dr = T.rotate(vol[None, ..., None], mats) if rotated else vol[None, ..., None]
    if liquid:
        want = 1.0 - torch.exp(-dr.sum(1) * tau)
    else:
        cs = torch.flip(torch.cumsum(torch.flip(dr, [1]), 1), [1])
        want = (dr * torch.exp(-cs * tau)).sum(1)
    want = want[..., 0]
    close(img, want, what='raymarch fwd')
    g = torch.tensor(rng.randn(nv, H, W).astype(np.float32))
    g[:, 2, 3:6] = 0.0
    (want * g).sum().backward()
    g_vol = torch.zeros(D, H, W, dtype=torch.float32, device=dev)
    _lib.get().call('lnst_set_raymarch_merge', 0 if rotated == 'merge-off' else 1)
    try:
        ops.raymarch_bwd(vol.detach().to(dev), rot, tau, liquid, stot, g.to(dev), g_vol)
    finally:
        _lib.get().call('lnst_set_raymarch_merge', 1)
    close(g_vol, vol.grad, tol=2e-5, what='raymarch bwd')


def test_normalize_fwd_bwd_incl_max_gradient(dev):
    rng = np.random.RandomState(9)
    img = torch.tensor(rng.rand(3, 6, 5).astype(np.float32))
    img[1, 2, 3] = img[1].max()                  # a tie inside image 1 (a second arg-max)
    img[1, 0, 0] = img[1].max()
    img.requires_grad_(True)
    stats = torch.empty(6, dtype=torch.float32, device=dev)
    gray = torch.empty(3, 6, 5, dtype=torch.float32, device=dev)
    ops.image_max(img.detach().to(dev), stats)
    ops.normalize_fwd(img.detach().to(dev), stats, gray)
    want = torch.stack([img[v] / torch.amax(img[v]) for v in range(3)])
    close(gray, want, what='normalise')
    g = torch.tensor(rng.randn(3, 6, 5).astype(np.float32))
    (want * g).sum().backward()
    dots = torch.empty(3, dtype=torch.float32, device=dev)
    g_img = torch.empty_like(gray)
    ops.normalize_bwd(img.detach().to(dev), stats, g.to(dev), dots, g_img)
    close(g_img, img.grad, tol=2e-5, what='normalise bwd')


def test_resize_and_net_input(dev):
    rng = np.random.RandomState(10)
    x = torch.tensor(rng.rand(2, 6, 8, 1).astype(np.float32), requires_grad=True)
    oh, ow = R.resized_hw(6, 8, 1.5)
    y = ops.resize_bilinear_fwd(x.detach().to(dev), oh, ow)
    want = R.resize_bilinear_legacy(x, oh, ow)
    close(y, want, what='resize fwd')
    g = torch.tensor(rng.randn(*want.shape).astype(np.float32))
    (want * g).sum().backward()
    close(ops.resize_bilinear_bwd(g.to(dev), 6, 8), x.grad, what='resize bwd')
    d_img = torch.empty(2, oh, ow, 3, dtype=torch.float32, device=dev)
    xin = torch.empty_like(d_img)
    ops.to_net_input_fwd(y, 255.0, d_img, xin)
    w_img = R.to_loss_net_input(want.detach(), 1.0, 'd')
    close(d_img, w_img, what='d_img')
    close(xin, V.preprocess(w_img), what='net input')
    gg = torch.tensor(rng.randn(2, oh, ow, 3).astype(np.float32)).to(dev)
    g_gray = torch.empty(2, oh, ow, 1, dtype=torch.float32, device=dev)
    ops.to_net_input_bwd(gg, 1, 255.0, g_gray)
    close(g_gray[..., 0], gg.cpu().sum(-1) * 255.0, what='net input bwd')


# ---- loss net -------------------------------------------------------------------------------
@pytest.mark.parametrize('cin,cout,H,W', [(3, 64, 9, 7), (64, 32, 6, 10), (16, 70, 5, 5)])
def test_conv3x3_f32_fwd_and_dgrad(dev, cin, cout, H, W):
    rng = np.random.RandomState(11)
    x = torch.tensor(rng.randn(2, H, W, cin).astype(np.float32), requires_grad=True)
    w = torch.tensor((rng.randn(3, 3, cin, cout) / np.sqrt(9 * cin)).astype(np.float32))
    b = torch.tensor(rng.randn(cout).astype(np.float32))
    y = ops.conv3x3_f32(x.detach().to(dev), w.to(dev), b.to(dev), relu=True)
    want = torch.relu(torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, padding=1))
    want = want.permute(0, 2, 3, 1)
    close(y, want, tol=2e-5, what='conv fwd')
    g = torch.tensor(rng.randn(*want.shape).astype(np.float32))
    (want * g).sum().backward()
    gm = (g * (want > 0)).detach()               # ReLU backward applied by the producer
    wd = w.flip(0, 1).permute(0, 1, 3, 2).contiguous()
    gx = ops.conv3x3_f32(gm.to(dev), wd.to(dev), None, relu=False)
    close(gx, x.grad, tol=2e-5, what='conv dgrad')


def test_avgpool_fwd_bwd_odd_size(dev):
    rng = np.random.RandomState(12)
    x = torch.tensor(rng.randn(2, 7, 9, 5).astype(np.float32), requires_grad=True)
    want = torch.nn.functional.avg_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    close(ops.avgpool2_fwd(x.detach().to(dev)), want, what='pool fwd')
    g = torch.tensor(rng.randn(*want.shape).astype(np.float32))
    (want * g).sum().backward()
    close(ops.avgpool2_bwd(g.to(dev), None, x.shape), x.grad, what='pool bwd')


def test_gram_loss_fwd_bwd(dev):
    rng = np.random.RandomState(13)
    P, ch = 150, 70
    F = torch.tensor(np.maximum(rng.randn(P, ch), 0).astype(np.float32), requires_grad=True)
    Fs = torch.tensor(np.maximum(rng.randn(40, ch), 0).astype(np.float32))
    Gs = torch.empty(ch, ch, dtype=torch.float32, device=dev)
    loss = torch.zeros(1, dtype=torch.float32, device=dev)
    ops.gram_diff(Fs.to(dev), 2.0 * 40 * ch, None, 0.0, Gs, None)
    close(Gs, Fs.t() @ Fs / (2.0 * 40 * ch), what='style gram')
    G = torch.empty_like(Gs)
    ops.gram_diff(F.detach().to(dev), 2.0 * P * ch, Gs, 0.7, G, loss)
    want, _ = L.style_loss([F.reshape(1, 10, 15, ch)], [Fs.reshape(1, 5, 8, ch)], [0.7], 1)
    close(loss, want.reshape(1), tol=2e-5, what='style loss')
    want.backward()
    gF = torch.empty(P, ch, dtype=torch.float32, device=dev)
    ops.gram_bwd(F.detach().to(dev), G, 0.7 * 4.0 / (2.0 * P * ch), 0.0, 1, gF)
    close(gF, F.grad * (F > 0), tol=2e-5, what='gram bwd')


def test_content_and_tv_loss(dev):
    rng = np.random.RandomState(14)
    F = torch.tensor(np.maximum(rng.randn(1, 6, 5, 12), 0).astype(np.float32), requires_grad=True)
    want = L.content_loss(F, 4) * 1.5
    want.backward()
    loss = torch.zeros(1, dtype=torch.float32, device=dev)
    gF = torch.empty(30, 12, dtype=torch.float32, device=dev)
    ops.content_loss(F.detach().reshape(30, 12).to(dev), 4, 1.5, loss, gF, 0.0, 0)
    close(loss, want.reshape(1), what='content')
    close(gF, F.grad.reshape(30, 12), what='content bwd')
    img = torch.tensor((rng.rand(1, 7, 6, 3) * 255).astype(np.float32), requires_grad=True)
    tv = L.tv_loss(img) * 0.01
    tv.backward()
    loss.zero_()
    g = torch.empty(7, 6, 3, dtype=torch.float32, device=dev)
    ops.tv_loss(img.detach()[0].to(dev), 0.01, loss, g)
    close(loss, tv.reshape(1), what='tv')
    close(g, img.grad[0], what='tv bwd')


# ---- active-box variants: same results inside the box, nothing touched outside it ------------------
def _box_volume(rng, shape, lo, hi):
    vol = np.zeros(shape, np.float32)
    sl = tuple(slice(a, b + 1) for a, b in zip(lo, hi))
    vol[sl] = rng.rand(*[b - a + 1 for a, b in zip(lo, hi)]) * (rng.rand(*[b - a + 1 for a, b in zip(lo, hi)]) > 0.3)
    return vol, sl


@pytest.mark.parametrize('liquid', [False, True])
@pytest.mark.parametrize('case', ['rot-inner', 'rot-face', 'plain'])
def test_raymarch_box_equals_full_inside(dev, liquid, case):
    """Rays marched only through the interval that can touch the active box: image bit-identical to
    the full march (the density is zero outside), gradient identical inside the box, zero outside.
    'rot-face': the box touches volume faces, where clamped out-of-volume samples still land in it."""
    rng = np.random.RandomState(21)
    D, H, W = 14, 9, 37
    lo, hi = ((4, 2, 9), (10, 6, 27)) if case != 'rot-face' else ((0, 2, 20), (9, 8, 36))
    vol_np, sl = _box_volume(rng, (D, H, W), lo, hi)
    vol = torch.tensor(vol_np).to(dev)
    box = _lib.make_box(lo, hi)
    mats = _rots()[:3] + [np.matmul(T.rot_y_3d(35.0), T.rot_z_3d(-50.0)), np.identity(3)]
    rot = None if case == 'plain' else torch.tensor(np.asarray(mats), dtype=torch.float32).reshape(-1, 9).to(dev)
    nv = 1 if rot is None else rot.shape[0]
    tau = 0.3
    img_f, st_f = torch.empty(nv, H, W, device=dev), torch.empty(nv, H, W, device=dev)
    img_b, st_b = torch.empty(nv, H, W, device=dev), torch.empty(nv, H, W, device=dev)
    ops.raymarch_fwd(vol, rot, tau, liquid, img_f, st_f)
    ops.raymarch_fwd(vol, rot, tau, liquid, img_b, st_b, box)
    assert torch.equal(img_f.cpu(), img_b.cpu()) and torch.equal(st_f.cpu(), st_b.cpu())
    g = torch.tensor(rng.randn(nv, H, W).astype(np.float32)).to(dev)
    g_full = torch.zeros(D, H, W, device=dev)
    g_box = torch.zeros(D, H, W, device=dev)
    ops.raymarch_bwd(vol, rot, tau, liquid, st_f, g, g_full)
    ops.raymarch_bwd(vol, rot, tau, liquid, st_f, g, g_box, box)
    close(g_box[sl], g_full[sl], tol=2e-5, what='box gradient inside')
    if case == 'plain':
        outside = g_box.clone()
        outside[sl] = 0
        assert float(outside.abs().max()) == 0.0


@pytest.mark.parametrize('liquid', [False, True])
def test_raymarch_bricks_skip_only_zeros(dev, liquid):
    """Occupancy bricks (built like Styler._workspace does): rays also skip the empty bricks at both ends of
    their box interval.  Image bit-identical; gradient identical on every active voxel."""
    rng = np.random.RandomState(23)
    D, H, W = 56, 52, 60
    z, y, x = np.meshgrid(np.arange(D), np.arange(H), np.arange(W), indexing='ij')
    # two blobs in opposite corners: their bounding box is most of the volume, most bricks are empty
    reach = (((z - 12) / 7.0) ** 2 + ((y - 11) / 6.0) ** 2 + ((x - 13) / 8.0) ** 2 <= 1.0) | \
            (((z - 44) / 6.0) ** 2 + ((y - 40) / 7.0) ** 2 + ((x - 47) / 6.0) ** 2 <= 1.0)          # "wmap > 0"
    mp = torch.nn.functional.max_pool3d
    o = torch.tensor(reach.astype(np.float32))[None, None]
    active = mp(o, 3, 1, 1)
    bricks = (mp(mp(mp(active, 5, 1, 2), 4, 4, 0, ceil_mode=True), 3, 1, 1)[0, 0] > 0).to(torch.uint8).contiguous()
    active = active[0, 0] > 0
    vol_np = (rng.rand(D, H, W) * (rng.rand(D, H, W) > 0.3)).astype(np.float32) * active.numpy()
    idx = np.argwhere(reach)
    lo, hi = np.maximum(idx.min(0) - 1, 0), np.minimum(idx.max(0) + 1, [D - 1, H - 1, W - 1])
    box = _lib.make_box(lo.tolist(), hi.tolist())
    vol = torch.tensor(vol_np).to(dev)
    mats = _rots()[:3] + [np.matmul(T.rot_y_3d(35.0), T.rot_z_3d(-50.0)), np.matmul(T.rot_y_3d(90.0), T.rot_z_3d(0.0))]
    rot = torch.tensor(np.asarray(mats), dtype=torch.float32).reshape(-1, 9).to(dev)
    nv, tau = rot.shape[0], 0.3
    img_f, st_f = torch.empty(nv, H, W, device=dev), torch.empty(nv, H, W, device=dev)
    img_b, st_b = torch.empty(nv, H, W, device=dev), torch.empty(nv, H, W, device=dev)
    ops.raymarch_fwd(vol, rot, tau, liquid, img_f, st_f)
    iv = ops.ray_intervals(rot, (D, H, W), box, bricks.to(dev))
    iv_box = ops.ray_intervals(rot, (D, H, W), box, None)
    n_box = (iv_box[..., 1] - iv_box[..., 0] + 1).clamp(min=0).sum().item()
    n_br = (iv[..., 1] - iv[..., 0] + 1).clamp(min=0).sum().item()
    assert 0 < n_br < 0.9 * n_box                                       # the bricks shorten the intervals
    ops.raymarch_fwd(vol, rot, tau, liquid, img_b, st_b, box, iv)
    assert torch.equal(img_f.cpu(), img_b.cpu()) and torch.equal(st_f.cpu(), st_b.cpu())
    g = torch.tensor(rng.randn(nv, H, W).astype(np.float32)).to(dev)
    g_full, g_br = torch.zeros(D, H, W, device=dev), torch.zeros(D, H, W, device=dev)
    ops.raymarch_bwd(vol, rot, tau, liquid, st_f, g, g_full)
    ops.raymarch_bwd(vol, rot, tau, liquid, st_f, g, g_br, box, iv)
    close(g_br.cpu()[active], g_full.cpu()[active], tol=2e-5, what='brick gradient on active voxels')
    assert float(g_br.abs().sum()) < float(g_full.abs().sum())          # work was actually skipped


def test_smooth_fill_and_wavg_box_variants(dev):
    rng = np.random.RandomState(22)
    shape, lo, hi = (9, 11, 13), (2, 1, 3), (6, 9, 10)
    box = _lib.make_box(lo, hi)
    d_np, sl = _box_volume(rng, shape, (3, 2, 4), (5, 8, 9))           # support one voxel inside the box
    d_np = d_np - 0.2 * (d_np > 0)
    d = torch.tensor(d_np).to(dev)
    full, part = torch.empty(shape, device=dev), torch.full(shape, 7.0, device=dev)
    ops.smooth3_relu_fwd(d, full, 3)
    ops.smooth3_relu_fwd(d, part, 3, box)
    bsl = tuple(slice(a, b + 1) for a, b in zip(lo, hi))
    assert torch.equal(part[bsl].cpu(), full[bsl].cpu())
    rest = part.clone()
    rest[bsl] = 7.0
    assert bool((rest == 7.0).all())                                  # nothing written outside the box
    g = torch.zeros(shape)
    g[bsl] = torch.tensor(rng.randn(*[b - a + 1 for a, b in zip(lo, hi)]).astype(np.float32))
    g = g.to(dev)
    gf, gp = torch.empty(shape, device=dev), torch.zeros(shape, device=dev)
    ops.smooth3_relu_bwd(g, full, gf, 3)
    ops.smooth3_relu_bwd(g, full, gp, 3, box)
    assert torch.equal(gp[bsl].cpu(), gf[bsl].cpu())
    ops.fill_box(part, box, 0.0)
    assert float(part[bsl].abs().max()) == 0.0 and float(part[0, 0, 0]) == 7.0
    # weighted-average splat: combine only inside the box, workspace left zero
    res, domain = [10, 10, 10], [10, 10, 10]
    p = torch.tensor(rng.uniform(0.35, 0.65, (800, 3)).astype(np.float32)).to(dev)
    r = torch.tensor(rng.uniform(0.2, 1, (800, 2)).astype(np.float32)).to(dev)
    var = torch.tensor(rng.uniform(-1.3, 1.3, (800, 2)).astype(np.float32)).to(dev)
    hs = [2.0, 1.0]
    grid = _lib.make_grid(3, res, domain, 1, False)
    wmap = ops.splat_wavg_wmap(p, grid, hs)
    occ = (wmap > 0).any(0).reshape(res)
    idx = torch.nonzero(occ)
    wb = _lib.make_box(idx.min(0).values.tolist(), idx.max(0).values.tolist())
    num = torch.zeros_like(wmap)
    out_f, out_b = torch.empty(res, device=dev), torch.zeros(res, device=dev)
    ops.splat_wavg_fwd(p, r, var, grid, hs, wmap, torch.empty_like(wmap), out_f)
    for _ in range(2):                                                # second pass: num was left clean
        ops.splat_wavg_fwd(p, r, var, grid, hs, wmap, num, out_b, wb)
        close(out_b, out_f, what='wavg box')
        assert float(num.abs().max()) == 0.0


# ---- optimiser / glue -----------------------------------------------------------------------
def test_adam_matches_tf_formula_with_nan(dev):
    rng = np.random.RandomState(15)
    var0 = torch.tensor(rng.randn(50, 2).astype(np.float32))
    var = var0.clone().to(dev)
    m = torch.zeros_like(var)
    v = torch.zeros_like(var)
    ref = TFAdam()
    want = var0.clone()
    for t in range(1, 4):
        g = torch.tensor(rng.randn(50, 2).astype(np.float32))
        g[3, 1] = float('nan')
        lr_t = 0.1 * np.sqrt(1 - np.float32(0.999) ** t) / (1 - np.float32(0.9) ** t)
        ops.adam_step(var, g.to(dev), m, v, lr_t)
        want = ref.step(want, g, 0.1)
    close(var, want, what='adam')
    # like the TF variable, the element stays NaN (with m, v) until the loop re-assigns it from the host iterate
    assert torch.isnan(var[3, 1]).item() and torch.isnan(m[3, 1]).item()


def test_adam_device_step_counter(dev):
    """lnst_adam_step_dev: beta powers and lr_t live on the device (CUDA-graph replayable)."""
    rng = np.random.RandomState(17)
    var0 = torch.tensor(rng.randn(33, 3).astype(np.float32))
    var = var0.clone().to(dev)
    m, v = torch.zeros_like(var), torch.zeros_like(var)
    state = torch.tensor([0.9, 0.999, 0.0], dtype=torch.float32).to(dev)
    ref = TFAdam()
    want = var0.clone()
    for t in range(1, 6):
        g = torch.tensor(rng.randn(33, 3).astype(np.float32))
        ops.adam_step_dev(var, g.to(dev), m, v, state, 0.05, gscale=0.5)
        want = ref.step(want, g * 0.5, 0.05)
    close(var, want, what='adam dev')
    np.testing.assert_allclose(state.cpu().numpy()[:2], [np.float32(0.9) ** 6, np.float32(0.999) ** 6], rtol=1e-6)


@pytest.mark.parametrize('apply', [False, True])
def test_fused_adam_iterate_equals_the_separate_kernels(dev, apply):
    """lnst_adam_iterate_dev = clone + lnst_adam_step_dev + lnst_iterate_delta (+ lnst_axpy), NaN rule included."""
    rng = np.random.RandomState(23)
    n, c = 37, 2
    g0 = torch.tensor(rng.randn(n, c).astype(np.float32))
    mask = torch.tensor(rng.rand(n, c).astype(np.float32))
    ref_var, ref_m, ref_v = g0.clone().to(dev), torch.zeros(n, c).to(dev), torch.zeros(n, c).to(dev)
    ref_state = torch.tensor([0.9, 0.999, 0.0]).to(dev)
    ref_g = g0.clone().to(dev)
    g_opt, m, v = g0.clone().to(dev), torch.zeros(n, c).to(dev), torch.zeros(n, c).to(dev)
    state = torch.tensor([0.9, 0.999, 0.0]).to(dev)
    for t in range(3):
        grad = torch.tensor(rng.randn(n, c).astype(np.float32))
        grad[5, 1] = float('nan')
        # separate kernels
        ref_var = ref_g.clone()
        ops.adam_step_dev(ref_var, grad.to(dev), ref_m, ref_v, ref_state, 0.05, gscale=0.25)
        ref_delta = ops.iterate_delta(ref_var, 1.0, ref_g, mask.to(dev), c, torch.empty_like(ref_var))
        # fused
        before = g_opt.clone()
        var, delta = ops.adam_iterate_dev(g_opt, grad.to(dev), m, v, state, 0.05, 0.25, mask.to(dev), c,
                                          torch.empty_like(g_opt), torch.empty_like(g_opt), apply)
        assert torch.equal(torch.isnan(var), torch.isnan(ref_var)) and torch.isnan(var[5, 1])
        assert torch.equal(torch.nan_to_num(var), torch.nan_to_num(ref_var))
        assert torch.equal(delta, ref_delta)
        assert delta[5, 1].item() == (-before[5, 1].cpu() * mask[5, 0]).item()      # NaN iterate counts as 0
        ops.axpy(ref_g, ref_delta, 1.0)
        if not apply:
            assert torch.equal(g_opt, before)
            ops.axpy(g_opt, delta, 1.0)
        assert torch.equal(g_opt, ref_g)
    assert torch.equal(torch.nan_to_num(m), torch.nan_to_num(ref_m)) and torch.equal(state, ref_state)


def test_iterate_glue_and_temporal_gauss(dev):
    from scipy.ndimage import gaussian_filter
    rng = np.random.RandomState(16)
    a = torch.tensor(rng.randn(20, 2).astype(np.float32))
    a[2, 0] = float('nan')
    acc = torch.empty(20, 2, dtype=torch.float32, device=dev)
    ops.iterate_accumulate(acc, a.to(dev), 1)
    ops.iterate_accumulate(acc, a.to(dev), 0)
    close(acc, 2 * torch.nan_to_num(a), what='accumulate')
    g_opt = torch.tensor(rng.randn(20, 2).astype(np.float32))
    r = torch.tensor(rng.rand(20, 2).astype(np.float32))
    delta = torch.empty_like(acc)
    ops.iterate_delta(acc, 0.5, g_opt.to(dev), r.to(dev), 2, delta)
    close(delta, (torch.nan_to_num(a) - g_opt) * r[:, 0:1], what='delta')
    x = rng.randn(7, 11, 3).astype(np.float32)
    for sigma in (0.8, 3.0):
        y = ops.temporal_gauss(torch.tensor(x).to(dev), sigma)
        close(y, torch.tensor(gaussian_filter(x, sigma=(sigma, 0, 0))), tol=1e-5, what='temporal gauss')


@pytest.mark.parametrize('dim', [2, 3])
def test_advect(dev, dim):
    rng = np.random.RandomState(17)
    shape = (6, 7) if dim == 2 else (5, 6, 4)
    d = torch.tensor(rng.rand(*shape, 2).astype(np.float32))
    vel = torch.tensor(rng.uniform(-0.5, 0.5, shape + (dim,)).astype(np.float32))
    out = ops.advect(d.to(dev), vel.to(dev))
    want = T.advect(d[None], vel[None], is_3d=(dim == 3))[0]
    close(out, want, what='advect')
