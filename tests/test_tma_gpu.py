"""TMA-tiled volume kernels (csrc/tiles_tma.cu) against the SIMT kernels they shadow -- GPU only.  The SIMT kernels are
the ones pinned against the oracle / the reference's vectors (test_kernel_parity.py, test_reference_golden.py, here and on
the CPU interpreter); these tests pin the TMA versions to them: ray-march images bit for bit, the rest to round-off."""
import numpy as np
import pytest
import torch

from lnst import _lib, ops
from lnst.transform import rot_mat

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


@pytest.fixture(autouse=True)
def cuda_lib():
    prev = _lib._lib
    _lib.set_for_testing(None)
    lib = _lib.get()
    assert lib.has_tma, 'TMA entry points missing from the CUDA library'
    yield
    torch.cuda.synchronize()
    ops.USE_TMA, ops.USE_TMA_BWD = True, False
    _lib.set_for_testing(prev)


def both(fn):
    """run fn() with the SIMT and with the TMA kernels"""
    outs = []
    for tma in (False, True):
        ops.USE_TMA = ops.USE_TMA_BWD = tma
        n0 = _lib.get().launches
        outs.append(fn())
    ops.USE_TMA, ops.USE_TMA_BWD = True, False
    return outs


def blob(D, H, W, seed=0, frac=0.6):
    """smooth non-negative density that is exactly zero outside a centred ellipsoid, and that ellipsoid's box"""
    rng = np.random.RandomState(seed)
    z, y, x = np.meshgrid(np.linspace(-1, 1, D), np.linspace(-1, 1, H), np.linspace(-1, 1, W), indexing='ij')
    rr = (z / frac) ** 2 + (y / (frac * 1.1)) ** 2 + (x / frac) ** 2
    v = np.where(rr < 1, rng.rand(D, H, W) * (1 - rr), 0).astype(np.float32)
    nz = np.nonzero(v)
    lo = [max(int(a.min()) - 1, 0) for a in nz]
    hi = [min(int(a.max()) + 1, s - 1) for a, s in zip(nz, (D, H, W))]
    return torch.tensor(v).to(DEV), _lib.make_box(lo, hi)


def views(phi, theta, n=None):
    rng = np.random.RandomState(1)
    mats, _ = rot_mat(-phi, phi, phi, -theta, theta, theta, sample_type='uniform', rng=rng, nv=n)
    return torch.tensor(np.asarray(mats, np.float64).reshape(-1, 9), dtype=torch.float32).to(DEV)


@pytest.mark.parametrize('shape', [(24, 24, 24), (20, 28, 36), (200, 200, 200), (9, 5, 8), (64, 64, 132)])
@pytest.mark.parametrize('use_box', [False, True])
def test_smooth3_tma_equals_simt(shape, use_box):
    D, H, W = shape
    d, box = blob(D, H, W, seed=D)
    d = d - 0.05 * (d > 0)                                   # some negative pre-activations for the ReLU marker
    box = box if use_box else None
    g = torch.tensor(np.random.RandomState(3).randn(D, H, W).astype(np.float32)).to(DEV)

    def run():
        ds = ops.smooth3_relu_fwd(d, torch.zeros_like(d), 3, box)
        gd = ops.smooth3_relu_bwd(g, ds, torch.zeros_like(d), 3, box)
        return ds, gd
    (ds0, gd0), (ds1, gd1) = both(run)
    assert torch.equal(torch.signbit(ds0), torch.signbit(ds1))        # the -0.0 markers
    assert (ds0 - ds1).abs().max() <= 1e-6 * ds0.abs().max()
    assert (gd0 - gd1).abs().max() <= 2e-6 * gd0.abs().max()


@pytest.mark.parametrize('shape,phi,theta', [((24, 24, 24), 5, 10), ((200, 200, 200), 5, 10), ((20, 28, 36), 5, 10),
                                             ((40, 40, 40), 40, 60),      # too oblique for the slab box: gather path
                                             ((64, 64, 64), 12, 15), ((33, 30, 32), 0, 0)])
@pytest.mark.parametrize('liquid', [False, True])
def test_raymarch_fwd_tma_is_bit_identical(shape, phi, theta, liquid):
    D, H, W = shape
    vol, box = blob(D, H, W, seed=H)
    rot = views(phi, theta) if phi or theta else torch.eye(3).reshape(1, 9).to(DEV)
    nv = rot.shape[0]
    for bx in (None, box):
        def run():
            img = torch.empty(nv, H, W, device=DEV)
            stot = torch.empty(nv, H, W, device=DEV)
            iv = ops.ray_intervals(rot, (D, H, W), bx, None) if bx is not None else None
            ops.raymarch_fwd(vol, rot, 0.01 if not liquid else 0.2, liquid, img, stot, bx, iv)
            return img, stot
        (i0, s0), (i1, s1) = both(run)
        assert torch.equal(i0, i1) and torch.equal(s0, s1)
        assert float(i0.max()) > 0


@pytest.mark.parametrize('shape,phi,theta', [((24, 24, 24), 5, 10), ((200, 200, 200), 5, 10), ((20, 28, 36), 5, 10),
                                             ((40, 40, 40), 40, 60), ((64, 64, 64), 12, 15), ((33, 30, 32), 0, 0),
                                             ((48, 48, 48), 2, 25)])
def test_raymarch_bwd_tma_equals_simt(shape, phi, theta):
    """gradient of the rotated render w.r.t. the volume: TMA-slab kernel (4 x 8 warp patches, depth / x / y merges of the
    corner atomics) against the SIMT kernel -- same sums in another order"""
    D, H, W = shape
    vol, box = blob(D, H, W, seed=H)
    rot = views(phi, theta) if phi or theta else torch.eye(3).reshape(1, 9).to(DEV)
    nv = rot.shape[0]
    g_img = torch.tensor(np.random.RandomState(2).randn(nv, H, W).astype(np.float32)).to(DEV)
    g_img[:, ::5, ::3] = 0.0                                  # rays without a cotangent are skipped
    for bx in (None, box):
        iv = ops.ray_intervals(rot, (D, H, W), bx, None) if bx is not None else None
        img = torch.empty(nv, H, W, device=DEV)
        stot = torch.empty(nv, H, W, device=DEV)
        ops.raymarch_fwd(vol, rot, 0.03, False, img, stot, bx, iv)

        def run():
            g_vol = torch.zeros(D, H, W, device=DEV)
            ops.raymarch_bwd(vol, rot, 0.03, False, stot, g_img, g_vol, bx, iv)
            return g_vol
        g0, g1 = both(run)
        assert float(g0.abs().max()) > 0
        assert (g0 - g1).abs().max() <= 2e-5 * g0.abs().max(), float((g0 - g1).abs().max() / g0.abs().max())


@pytest.mark.parametrize('res,nk,n', [((24, 24, 24), 2, 5000), ((20, 28, 36), 1, 3000), ((64, 48, 40), 3, 40000),
                                      ((200, 200, 200), 2, 300000)])
def test_splat_gather_equals_scatter(res, nk, n):
    """p2g_wavg forward as a gather over per-cell lists (TMA-stored tiles) against the atomics kernel: same weights,
    another summation order; particles on the faces, padding rows (p = -1), variables outside [-1, 1] and NaN."""
    from lnst import synth
    D, H, W = res
    rng = np.random.RandomState(n)
    p = rng.uniform(0.15, 0.85, (n, 3)).astype(np.float32)
    p[:50] = rng.uniform(0.0, 1.0, (50, 3))                   # some next to / on the faces
    p[50:60] = -1.0                                           # padding rows
    p[60:64] = [[0.0, 0.0, 0.0], [0.999999, 0.5, 0.5], [0.5, 0.999999, 0.5], [0.5, 0.5, 0.999999]]
    r = rng.uniform(0.2, 1.0, (n, nk)).astype(np.float32)
    var = (rng.randn(n, nk) * 0.7).astype(np.float32)
    var[100:104] = np.nan
    grid = _lib.make_grid(3, res, [float(v) for v in res], 1, False)
    hs = [2.0 / (2 ** k) for k in range(nk)]
    pt, rt, vt = (torch.tensor(a).to(DEV) for a in (p, r, var))
    wmap = ops.splat_wavg_wmap(pt, grid, hs)
    num = torch.zeros(nk, D * H * W, device=DEV)
    want = ops.splat_wavg_fwd(pt, rt, vt, grid, hs, wmap, num, torch.zeros(D, H, W, device=DEV), None)
    lists = ops.cell_lists(pt, grid)
    assert lists is not None
    got = ops.splat_wavg_fwd_gather(lists, rt, vt, grid, hs, torch.zeros(D, H, W, device=DEV), None)
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    ok = ~torch.isnan(want)
    assert (got[ok] - want[ok]).abs().max() <= 2e-5 * want[ok].abs().max()
    # box variant: only the box is written, and it holds the same values
    occ = (wmap > 0).any(0).reshape(D, H, W)
    idx = torch.nonzero(occ)
    lo, hi = idx.min(0).values.tolist(), idx.max(0).values.tolist()
    box = _lib.make_box(lo, hi)
    got_b = ops.splat_wavg_fwd_gather(lists, rt, vt, grid, hs, torch.zeros(D, H, W, device=DEV), box)
    assert torch.equal(torch.isnan(got_b), torch.isnan(want))
    assert (got_b[ok] - want[ok]).abs().max() <= 2e-5 * want[ok].abs().max()


@pytest.mark.parametrize('shape,amp', [((24, 24, 24), 1.5), ((128, 128, 128), 2.0), ((20, 28, 36), 0.5), ((16, 16, 32), 6.0)])
def test_advect_tma_equals_simt(shape, amp):
    """order-1 semi-Lagrangian advection of a scalar field with the source tiles staged by TMA against the gather kernel;
    velocities up to `amp` cells (6 cells: beyond the staged reach, the global-memory branch)"""
    D, H, W = shape
    rng = np.random.RandomState(D + W)
    d = torch.tensor(rng.rand(D, H, W, 1).astype(np.float32)).to(DEV)
    cells = np.array([D - 1, H - 1, W - 1], np.float32)
    vel = torch.tensor((rng.uniform(-amp, amp, (D, H, W, 3)) * 2.0 / cells).astype(np.float32)).to(DEV)
    o0, o1 = both(lambda: ops.advect(d, vel))
    assert (o0 - o1).abs().max() <= 1e-6 * o0.abs().max()
    for reach in (1, 4):
        ops.ADVECT_REACH = reach
        try:
            assert (ops.advect(d, vel) - o0).abs().max() <= 1e-6 * o0.abs().max()
        finally:
            ops.ADVECT_REACH = 2


@pytest.mark.parametrize('n,H,W', [(2, 13, 9), (3, 50, 64), (1, 200, 200), (2, 16, 16), (1, 33, 47)])
@pytest.mark.parametrize('split', [False, True])
def test_conv_first_bwd_gray_direct(n, H, W, split):
    """conv1_1's data gradient w.r.t. the gray render on the CUDA cores (TMA-staged patch) against an fp64 transposed
    convolution of the same bf16 / split-bf16 cotangent"""
    from lnst import synth, vgg
    gen = torch.Generator().manual_seed(H * 7 + W)
    net = vgg.LossNet(synth.vgg_weights(), 'vgg_19', DEV, math='bf16x3' if split else 'bf16')
    w = net.w['conv1_1'].double().cpu()
    g32 = torch.randn(n, H, W, 64, generator=gen)
    g_dev = ops.to_split(g32.to(DEV)) if split else g32.to(torch.bfloat16).to(DEV)
    g_val = ops.from_split(g_dev).cpu().double() if split else g_dev.float().cpu().double()
    gx = torch.nn.functional.conv_transpose2d(g_val.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding=1)
    want = 255.0 * gx.sum(1)
    got = ops.conv_first_bwd_gray_direct(g_dev, split, net.tc.wg_gray).cpu().double()
    err = (got - want).abs().max().item()
    assert err <= 3e-6 * want.abs().max().item(), (err, want.abs().max().item())
