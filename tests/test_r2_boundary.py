"""Round-2 additions to the boundary: library queries, the standalone rotate gradient, and the
pressure / density regularisers as kernels -- against torch autograd on the same inputs (emu here, cuda on the B200)."""
import ctypes

import numpy as np
import pytest
import torch

from lnst import _lib, ops
from oracle import transform as T
from test_kernel_parity import close


def test_version_and_workspace_queries():
    import os
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    dll = ctypes.CDLL(_lib.LIB_PATH)
    dll.lnst_version.restype = ctypes.c_char_p
    assert dll.lnst_version().decode().startswith('lnst-b200')
    dll.lnst_workspace_bytes.restype = ctypes.c_int64
    dll.lnst_workspace_bytes.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int64), ctypes.c_int32]
    dims = (ctypes.c_int64 * 2)(200 ** 3, 2)
    assert dll.lnst_workspace_bytes(b'splat_wavg_fwd', dims, 2) == 4 * 2 * 200 ** 3
    dims = (ctypes.c_int64 * 2)(200 * 200, 9)
    assert dll.lnst_workspace_bytes(b'raymarch_bwd', dims, 2) == 8 * 9 * 200 * 200
    assert dll.lnst_workspace_bytes(b'smooth3_relu_fwd', None, 0) == 0
    assert dll.lnst_workspace_bytes(b'no_such_op', None, 0) == -1
    assert dll.lnst_workspace_bytes(None, None, 0) == -1


def test_rotate_bwd_is_the_transpose_of_rotate_fwd(dev):
    rng = np.random.RandomState(5)
    D, H, W = 7, 9, 8
    vol = torch.tensor(rng.rand(D, H, W).astype(np.float32))
    from lnst.transform import rot_mat
    mats, _ = rot_mat(-20, 20, 20, -30, 30, 30, sample_type='uniform', rng=rng, nv=None)
    rot = torch.tensor(np.asarray(mats, np.float64).reshape(-1, 9), dtype=torch.float32)
    g_out = torch.tensor(rng.randn(rot.shape[0], D, H, W).astype(np.float32))
    got = ops.rotate_bwd(g_out.to(dev), rot.to(dev))
    v = vol.double().clone().requires_grad_(True)
    out = T.rotate(v[None, ..., None], [m for m in np.asarray(mats)])          # [nv, D, H, W, 1]
    (out[..., 0] * g_out.double()).sum().backward()
    close(got, v.grad.float(), tol=2e-6, what='rotate bwd')
    # adjoint identity <rotate(vol), g> = <vol, rotate_bwd(g)>
    fwd = ops.rotate_fwd(vol.to(dev), rot.to(dev))
    lhs = float((fwd.double() * g_out.to(dev).double()).sum())
    rhs = float((vol.to(dev).double() * got.double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)


def test_pressure_reg_accumulates(dev):
    rng = np.random.RandomState(2)
    d = rng.rand(6, 5, 7).astype(np.float32) * 2
    d[rng.rand(*d.shape) < 0.3] = 0.0
    g0 = rng.randn(*d.shape).astype(np.float32)
    loss = torch.tensor([0.25, 0.5, 7.0]).to(dev)
    g = torch.tensor(g0).to(dev)
    w, n_terms = 0.5, 2
    ops.pressure_reg(torch.tensor(d).to(dev), 1.0, w, n_terms * w * 2.0 / d.size, loss, n_terms, g)
    dt = torch.tensor(d, dtype=torch.float64, requires_grad=True)
    pr = torch.where(dt > 0, dt - 1, torch.zeros_like(dt))
    val = w * (pr * pr).mean()
    (n_terms * val).backward()
    close(loss, torch.tensor([0.25 + float(val), 0.5 + float(val), 7.0]), tol=2e-6, what='pressure loss')
    close(g, torch.tensor(g0).double() + dt.grad, tol=2e-6, what='pressure grad')


@pytest.mark.parametrize('n', [0, 1, 3000])
def test_density_reg(dev, n):
    rng = np.random.RandomState(4)
    var = (rng.randn(n, 2) * 0.8).astype(np.float32)              # some elements outside [-1, 1]
    g0 = rng.randn(n, 2).astype(np.float32)
    loss = torch.tensor([1.0, 2.0]).to(dev)
    grad = torch.tensor(g0).to(dev)
    w, n_terms = 1e-3, 1
    ops.density_reg(torch.tensor(var).to(dev), w, n_terms * w, loss, n_terms, grad)
    vt = torch.tensor(var, dtype=torch.float64, requires_grad=True)
    dv = torch.clamp(vt, -1, 1)
    val = w * (dv.sum() ** 2 + 1e3 * (-torch.log(dv.abs() + 1e-6)).sum())
    val.backward()
    close(loss, torch.tensor([1.0 + float(val), 2.0]), tol=3e-6, what='density loss')
    if n:
        close(grad, torch.tensor(g0).double() + vt.grad, tol=3e-6, what='density grad')


def _blob(D, H, W, seed=0, frac=0.55):
    rng = np.random.RandomState(seed)
    z, y, x = np.meshgrid(np.linspace(-1, 1, D), np.linspace(-1, 1, H), np.linspace(-1, 1, W), indexing='ij')
    rr = (z / frac) ** 2 + (y / (frac * 1.1)) ** 2 + (x / frac) ** 2
    occ = rr < 1
    v = np.where(occ, rng.rand(D, H, W) * (1.05 - rr), 0).astype(np.float32)
    return v, occ


@pytest.mark.parametrize('shape', [(16, 16, 16), (12, 18, 20)])
def test_exact_ray_intervals_leave_the_render_and_its_gradient_unchanged(dev, shape):
    """lnst_ray_intervals_exact: per ray the first / last sample whose footprint touches the support -- images stay bit
    identical to the full march, the volume gradient is the same wherever the support (the only place it is read) is."""
    import torch.nn.functional as Fn
    D, H, W = shape
    v, occ = _blob(D, H, W, seed=D)
    vol = torch.tensor(v).to(dev)
    act = torch.tensor(occ.astype(np.float32))[None, None]            # support of the volume itself (no blur here)
    touch = (Fn.max_pool3d(Fn.pad(act, (0, 1, 0, 1, 0, 1)), 2, 1, 0)[0, 0] > 0).to(torch.uint8).contiguous().to(dev)
    from lnst.transform import rot_mat
    mats, _ = rot_mat(-5, 5, 5, -10, 10, 10, sample_type='uniform', rng=np.random.RandomState(0), nv=None)
    rot = torch.tensor(np.asarray(mats, np.float64).reshape(-1, 9), dtype=torch.float32).to(dev)
    nv = rot.shape[0]
    iv = ops.ray_intervals_exact(rot, shape, None, touch)
    lo, hi = iv[..., 0].cpu().numpy(), iv[..., 1].cpu().numpy()
    assert ((hi - lo + 1).clip(0).sum()) < 0.8 * nv * D * H * W     # it does cut
    outs = []
    for use in (None, iv):
        img, stot = torch.empty(nv, H, W).to(dev), torch.empty(nv, H, W).to(dev)
        ops.raymarch_fwd(vol, rot, 0.05, False, img, stot, None, use)
        g_img = torch.tensor(np.random.RandomState(1).randn(nv, H, W).astype(np.float32)).to(dev)
        g_vol = torch.zeros(D, H, W).to(dev)
        ops.raymarch_bwd(vol, rot, 0.05, False, stot, g_img, g_vol, None, use)
        outs.append((img.cpu(), stot.cpu(), g_vol.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    m = torch.tensor(occ)
    close(outs[1][2][m], outs[0][2][m], tol=1e-5, what='gradient on the support')


def test_fused_image_glue_unrotated_march(dev):
    """the same for the un-rotated march (rot = None: one view along z, box-limited)"""
    rng = np.random.RandomState(5)
    D, H, W = 10, 9, 12
    vol = torch.tensor((rng.rand(D, H, W) * (rng.rand(D, H, W) > 0.4)).astype(np.float32)).to(dev)
    tau = 0.3
    for box in (None, _lib.make_box((1, 2, 3), (8, 7, 10))):
        v = vol.clone()
        if box is not None:                                  # density only inside the box, like the styler's volumes
            keep = torch.zeros_like(v)
            keep[1:9, 2:8, 3:11] = 1
            v = v * keep
        img0, st0 = torch.empty(1, H, W, device=dev), torch.empty(1, H, W, device=dev)
        ops.raymarch_fwd(v, None, tau, False, img0, st0, box)
        stats0 = ops.image_max(img0, torch.empty(2, device=dev))
        gray0 = ops.normalize_fwd(img0, stats0, torch.empty_like(img0))
        img1, st1, stats1 = torch.empty(1, H, W, device=dev), torch.empty(1, H, W, device=dev), torch.zeros(2, device=dev)
        ops.raymarch_fwd(v, None, tau, False, img1, st1, box, stats=stats1)
        gray1 = ops.normalize_ties_fwd(img1, stats1, torch.empty_like(img1))
        assert torch.equal(img0, img1) and torch.equal(st0, st1) and torch.equal(stats0, stats1) and torch.equal(gray0, gray1)
        g_gray = torch.tensor(rng.randn(1, H, W).astype(np.float32)).to(dev)
        dots = torch.empty(1, device=dev)
        g_img = ops.normalize_bwd(img0, stats0, g_gray, dots, torch.empty_like(g_gray))
        want = ops.raymarch_bwd(v, None, tau, False, st0, g_img, torch.zeros(D, H, W, device=dev), box)
        got = ops.raymarch_bwd(v, None, tau, False, st0, g_gray, torch.zeros(D, H, W, device=dev), box, norm=(img0, stats0, dots))
        assert torch.equal(want, got)


@pytest.mark.parametrize('shape', [(9, 7, 8), (11, 6, 45), (14, 33, 37)])
def test_fused_image_glue_equals_the_separate_calls(dev, shape):
    """lnst_raymarch_fwd_max_* + lnst_normalize_ties_fwd against the march, lnst_image_max and lnst_normalize_fwd; and
    lnst_raymarch_bwd_norm_box against lnst_normalize_bwd followed by the march: same maxima, ties, gray image, gradient."""
    rng = np.random.RandomState(shape[2])
    D, H, W = shape
    vol_np = (rng.rand(D, H, W) * (rng.rand(D, H, W) > 0.4)).astype(np.float32)
    vol = torch.tensor(vol_np).to(dev)
    mats = [np.identity(3), np.matmul(T.rot_y_3d(5.0), T.rot_z_3d(-10.0)), np.matmul(T.rot_y_3d(-33.0), T.rot_z_3d(21.0))]
    rot = torch.tensor(np.asarray(mats), dtype=torch.float32).reshape(-1, 9).to(dev)
    nv, tau = len(mats), 0.3
    img0, st0 = torch.empty(nv, H, W, device=dev), torch.empty(nv, H, W, device=dev)
    ops.raymarch_fwd(vol, rot, tau, False, img0, st0)
    stats0 = ops.image_max(img0, torch.empty(2 * nv, device=dev))
    gray0 = ops.normalize_fwd(img0, stats0, torch.empty_like(img0))
    img1, st1 = torch.empty(nv, H, W, device=dev), torch.empty(nv, H, W, device=dev)
    stats1 = torch.zeros(2 * nv, device=dev)
    ops.raymarch_fwd(vol, rot, tau, False, img1, st1, stats=stats1)
    gray1 = ops.normalize_ties_fwd(img1, stats1, torch.empty_like(img1))
    assert torch.equal(img0, img1) and torch.equal(st0, st1)
    assert torch.equal(stats0, stats1), (stats0, stats1)
    # ties: a hand-made image whose maximum is attained three times in view 1
    im = torch.tensor(rng.rand(nv, H, W).astype(np.float32)).to(dev)
    im[1, 0, 0] = im[1, H - 1, W - 1] = im[1, 2, 3] = 2.0
    sa = ops.image_max(im, torch.empty(2 * nv, device=dev))
    sb = torch.zeros(2 * nv, device=dev)
    sb[0::2] = sa[0::2]
    gb = ops.normalize_ties_fwd(im, sb, torch.empty_like(im))
    assert torch.equal(sa, sb) and float(sb[3]) == 3.0
    assert torch.equal(gb, ops.normalize_fwd(im, sa, torch.empty_like(im)))
    assert torch.equal(gray0, gray1)
    g_gray = torch.tensor(rng.randn(nv, H, W).astype(np.float32)).to(dev)
    dots = torch.empty(nv, device=dev)
    g_img = ops.normalize_bwd(img0, stats0, g_gray, dots, torch.empty_like(g_gray))
    want = ops.raymarch_bwd(vol, rot, tau, False, st0, g_img, torch.zeros(D, H, W, device=dev))
    got = ops.raymarch_bwd(vol, rot, tau, False, st0, g_gray, torch.zeros(D, H, W, device=dev), norm=(img0, stats0, dots))
    if dev.type == 'cpu':
        assert torch.equal(want, got)
    else:                                                    # atomics: the summation order differs from launch to launch
        close(got, want, tol=2e-6, what='fused normalisation gradient')
