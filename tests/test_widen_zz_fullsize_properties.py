"""Size-independent properties of the hot-path operators at BASELINE.json's headline size (C3: 200^3 volume, N = 2^20
particles x 2 kernels, 9 views) -- where an oracle comparison would take minutes, the domain's own invariants are
checked through the C-ABI instead:

  * the weighted-average splat reproduces constants (partition of unity) and is linear in the particle values;
  * 3x3x3 smoothing: <S d, g> == <d, S^T g> (adjoint identity) and it preserves the mean of an interior blob;
  * emission/absorption ray-march: closed form for a uniform slab, identity rotation == no rotation, zero field ->
    zero image, backward == central finite difference along a random direction;
  * TF-Adam: a zero gradient moves nothing; one step moves every element by lr (|g| >> eps).

On the B200 (``dev=cuda``) the sizes are the full ones; on the CPU interpreter here the same properties run at 20^3."""
import numpy as np
import torch

from lnst import _lib, ops, synth
from oracle import transform as T


def _sizes(dev):
    return (200, 1 << 20) if dev.type == 'cuda' else (20, 3000)


def _rel(a, b):
    a, b = float(a), float(b)
    return abs(a - b) / max(abs(a), abs(b), 1e-30)


def test_wavg_splat_partition_of_unity_and_linearity(dev):
    res, n = _sizes(dev)
    p, r = synth.smoke_particles(n, 2)
    p = torch.tensor(p[0]).to(dev)
    grid = _lib.make_grid(3, [res] * 3, [res] * 3, 1, False)
    hs = [0.5 * 4, 0.5 * 2]                                   # radius * support / kernel_scale^k (styler_3p.py:79-82)
    wmap = ops.splat_wavg_wmap(p, grid, hs)
    cells = res ** 3
    num = torch.empty(2, cells, device=dev)
    c = torch.tensor([0.7, -0.2]).to(dev)
    const = torch.ones(n, 2, device=dev) * c
    out = ops.splat_wavg_fwd(p, const, None, grid, hs, wmap, num, torch.empty(res, res, res, device=dev))
    covered = ((wmap[0] > 1e-6) & (wmap[1] > 1e-6)).reshape(res, res, res)
    assert int(covered.sum()) > 0.05 * cells
    assert float((out[covered] - 0.5).abs().max()) < 2e-5      # a weighted average of a constant is the constant
    outside = (wmap.sum(0) == 0).reshape(res, res, res)
    assert float(out[outside].abs().max()) == 0.0
    r1 = torch.tensor(r[0]).to(dev)
    r2 = torch.rand(n, 2, device=dev)
    o1 = ops.splat_wavg_fwd(p, r1, None, grid, hs, wmap, num, torch.empty_like(out)).clone()
    o2 = ops.splat_wavg_fwd(p, r2, None, grid, hs, wmap, num, torch.empty_like(out)).clone()
    o12 = ops.splat_wavg_fwd(p, (r1 + r2).contiguous(), None, grid, hs, wmap, num, torch.empty_like(out))
    assert float((o12 - (o1 + o2)).abs().max()) < 2e-5 * float(o12.abs().max())


def test_smoothing_adjoint_and_mean(dev):
    res, _ = _sizes(dev)
    g0 = torch.Generator().manual_seed(3)
    d = torch.zeros(res, res, res)
    q = res // 4
    d[q:3 * q, q:3 * q, q:3 * q] = torch.rand(2 * q, 2 * q, 2 * q, generator=g0) + 0.1      # positive, away from the faces
    g = torch.randn(res, res, res, generator=g0)
    d, g = d.to(dev), g.to(dev)
    out = ops.smooth3_relu_fwd(d, torch.empty_like(d), 3)
    gin = ops.smooth3_relu_bwd(g, out, torch.empty_like(d), 3)
    lhs = (out.double() * g.double()).sum()
    # maximum(.,0) passes the gradient wherever the smoothed value is >= 0, i.e. everywhere here: S^T g
    rhs = (d.double() * gin.double()).sum()
    assert _rel(lhs, rhs) < 1e-5
    assert _rel(out.double().sum(), d.double().sum()) < 1e-6    # the kernel sums to 1 and nothing reaches a face


def test_raymarch_invariants(dev):
    res, _ = _sizes(dev)
    tau, rho = 0.01, 0.37
    vol = torch.full((res, res, res), rho, device=dev)
    img, stot = torch.empty(1, res, res, device=dev), torch.empty(1, res, res, device=dev)
    ops.raymarch_fwd(vol, None, tau, False, img, stot)
    m = np.arange(1, res + 1, dtype=np.float64)
    want = float((rho * np.exp(-tau * rho * m)).sum())          # I = sum_m rho exp(-tau rho m), styler_3p.py:154-157
    # 200 fp32 additions of the same constant round in the same direction: a few 1e-6 relative on S, hence on I
    assert float((img - want).abs().max()) < 5e-5 * want
    assert float((stot - rho * res).abs().max()) < 5e-5 * rho * res
    ops.raymarch_fwd(vol, None, tau, True, img, stot)
    assert float((img - (1 - np.exp(-tau * rho * res))).abs().max()) < 5e-5      # liquid: 1 - exp(-tau sum d)
    eye = torch.eye(3).reshape(1, 9).to(dev)
    blob = torch.rand(res, res, res, generator=torch.Generator().manual_seed(5)).to(dev)
    a, b = torch.empty(1, res, res, device=dev), torch.empty(1, res, res, device=dev)
    ops.raymarch_fwd(blob, None, tau, False, a, stot)
    ops.raymarch_fwd(blob, eye, tau, False, b, stot)
    assert float((a - b).abs().max()) < 2e-5 * float(a.abs().max())
    ops.raymarch_fwd(torch.zeros_like(blob), eye, tau, False, b, stot)
    assert float(b.abs().max()) == 0.0


def test_raymarch_backward_is_the_directional_derivative(dev):
    res, _ = _sizes(dev)
    tau = 0.01 if dev.type == 'cuda' else 0.1
    gen = torch.Generator().manual_seed(7)
    vol = (torch.rand(res, res, res, generator=gen) * 0.5 + 0.25).to(dev)
    v = torch.randn(res, res, res, generator=gen).to(dev)
    mats, _ = T.rot_mat(-5, 5, 5, -10, 10, 10, sample_type='uniform')
    rot = torch.tensor(np.asarray(mats), dtype=torch.float32).reshape(-1, 9).to(dev)
    nv = rot.shape[0]
    w = torch.rand(nv, res, res, generator=gen).to(dev)

    def L(x):
        img, st = torch.empty(nv, res, res, device=dev), torch.empty(nv, res, res, device=dev)
        ops.raymarch_fwd(x.contiguous(), rot, tau, False, img, st)
        return (img.double() * w.double()).sum(), st

    _, st = L(vol)
    g_vol = torch.zeros_like(vol)
    ops.raymarch_bwd(vol, rot, tau, False, st, w, g_vol)
    analytic = (g_vol.double() * v.double()).sum()
    eps = 1e-2
    numeric = (L(vol + eps * v)[0] - L(vol - eps * v)[0]) / (2 * eps)
    assert _rel(analytic, numeric) < 2e-2, (float(analytic), float(numeric))


def test_adam_fixed_points(dev):
    _, n = _sizes(dev)
    var = torch.rand(n, 2, device=dev)
    v0 = var.clone()
    m, v = torch.zeros_like(var), torch.zeros_like(var)
    state = torch.tensor([0.9, 0.999, 0.0]).to(dev)
    ops.adam_step_dev(var, torch.zeros_like(var), m, v, state, 0.1)
    assert torch.equal(var.cpu(), v0.cpu()) and float(m.abs().max()) == 0 and float(v.abs().max()) == 0
    g = torch.where(torch.rand(n, 2, device=dev) > 0.5, 3.0, -2.0)
    ops.adam_step_dev(var, g, m, v, state, 0.1)                 # second step of this optimizer: t = 2
    m2, v2 = 0.1 * g, 0.001 * g * g                             # first non-zero gradient: m = (1-b1) g, v = (1-b2) g^2
    lr_t = 0.1 * np.sqrt(1 - 0.999 ** 2) / (1 - 0.9 ** 2)
    want = v0 - lr_t * m2 / (v2.sqrt() + 1e-8)
    assert float((var - want).abs().max()) < 1e-5
