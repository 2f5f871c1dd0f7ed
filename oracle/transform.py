"""Oracle restatement of the reference's differentiable fluid ops (``transform.py``).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  All tensors are torch CPU tensors; every
function is differentiable through torch autograd, which reproduces the TF-1.15 gradient
rules named in the docstrings.
"""
import math

import numpy as np
import torch


# --------------------------------------------------------------------------------------
# SPH kernel
# --------------------------------------------------------------------------------------
def cubic_w(q, h, is_3d):
    """Cubic-spline kernel, ``transform.py:1233-1245``.

    sigma * (q<=.5 ? 6(q^3-q^2)+1 : 2(1-q)^3), zero for q>1.  ``tf.where`` routes the
    gradient through the selected branch only; torch.where does the same.
    """
    sigma = 8.0 / math.pi / h ** 3 if is_3d else 40.0 / 7.0 / math.pi / h ** 2
    inner = 6.0 * (q ** 3 - q ** 2) + 1.0
    outer = 2.0 * (1.0 - q) ** 3
    w = sigma * torch.where(q <= 0.5, inner, outer)
    return torch.where(q > 1, torch.zeros_like(q), w)


def _safe_norm(sq):
    """sqrt with a zero (instead of NaN/inf) derivative at exactly 0.

    TF gives NaN for d sqrt(0) * 0 (``transform.py:1416``); the reference later swallows it
    with ``np.nan_to_num`` (``styler_3p.py:360``).  A particle sitting exactly on a cell
    centre is measure-zero for real data, so the oracle and the CUDA kernels both define
    the derivative as 0 there (documented deviation).
    """
    pos = sq > 0
    return torch.where(pos, torch.sqrt(torch.where(pos, sq, torch.ones_like(sq))), torch.zeros_like(sq))


def _cell_setup(p, domain, res, clip, eps):
    """Shared front part of p2g / p2g_wavg: ``transform.py:1316-1343`` / ``1583-1610``."""
    dt = p.dtype
    domain_ = torch.as_tensor([float(d) for d in domain], dtype=dt)
    res_ = torch.as_tensor([float(r) for r in res], dtype=dt)
    p = p * domain_
    if clip:
        p = torch.minimum(torch.clamp(p, min=0.0), domain_ - eps)
        valid = torch.ones(p.shape[:-1], dtype=torch.bool)
    else:
        valid = ((p >= 0) & (p < domain_)).all(dim=-1)
    cell = (domain_ / res_)[0]
    idx_f = torch.floor(p / cell)
    idx = idx_f.to(torch.int64)
    r = p - (idx_f + 0.5) * cell  # offset from the centre of the particle's own cell
    return p, valid, cell, idx, r


def _scatter(out_flat, res, idx, shift, upd, valid):
    """One ``tf.scatter_nd`` of the reference with GPU semantics: out-of-range targets are
    dropped, duplicates are summed (``transform.py:1430``)."""
    tgt = idx + torch.as_tensor(shift, dtype=torch.int64)
    ok = valid.clone()
    flat = torch.zeros(idx.shape[:-1], dtype=torch.int64)
    for a, n in enumerate(res):
        ok &= (tgt[..., a] >= 0) & (tgt[..., a] < n)
        flat = flat * n + tgt[..., a]
    # batch offset
    nb = idx.shape[0]
    vol = int(np.prod(res))
    flat = flat + torch.arange(nb, dtype=torch.int64).reshape(nb, 1) * vol
    ok_f = ok.reshape(-1)
    flat_f = flat.reshape(-1)[ok_f]
    if upd.dim() == idx.dim() - 1:  # scalar per particle
        out_flat.index_put_((flat_f,), upd.reshape(-1)[ok_f], accumulate=True)
    else:
        c = upd.shape[-1]
        out_flat.index_put_((flat_f,), upd.reshape(-1, c)[ok_f], accumulate=True)


def p2g(p, domain, res, radius, rest_density, nsize, pc=None, pd=None, is_2d=True,
        eps=1e-6, clip=True, support=4):
    """SPH particle->grid splat, ``transform.py:1310-1453``.

    p [B,N,dim] normalised positions ((z,)y,x order).  Returns [B,*res,1] (or [B,*res,C] with
    colours ``pc`` [B,N,C], divided by ``pd`` [B,N,1] or rest_density).  H axis flipped at the
    end (``:1404`` 2-D, ``:1452`` 3-D).
    """
    dim = 2 if is_2d else 3
    res = [int(r) for r in res]
    B = p.shape[0]
    _, valid, cell, idx, r = _cell_setup(p, domain, res, clip, eps)
    h = radius * support
    volume = 0.8 * (2 * radius) ** dim
    mass = volume * rest_density
    nch = 1 if pc is None else pc.shape[-1]
    out = torch.zeros(B * int(np.prod(res)), nch, dtype=p.dtype)
    rng = range(-nsize, nsize + 1)
    shifts = [(a, b) for a in rng for b in rng] if is_2d else \
             [(a, b, c) for a in rng for b in rng for c in rng]
    for s in shifts:
        d = r - torch.as_tensor(s, dtype=p.dtype) * cell
        q = _safe_norm((d ** 2).sum(-1)) / h
        w = cubic_w(q, h, is_3d=not is_2d)
        if pc is None:
            upd = (mass * w).unsqueeze(-1)
        else:
            upd = mass * w.unsqueeze(-1) * pc
            upd = upd / (rest_density if pd is None else pd)
        _scatter(out, res, idx, s, upd, valid)
    out = out.reshape([B] + res + [nch])
    return torch.flip(out, dims=[1 if is_2d else 2])


def p2g_wavg(p, x, domain, res, radius, nsize, is_2d=True, eps=1e-6, clip=True, support=4):
    """Weighted-average splat, ``transform.py:1577-1704`` (called with kernel='cubic',
    ``styler_3p.py:83``).  out = where(wmap>eps, sum(W x)/sum(W), sum(W x)), H flipped.

    NB the ``where`` has the classic TF NaN-gradient trap: cells with wmap == 0 produce
    0/0 = NaN in the gradient of the unselected branch (``:1703``), which reaches every
    particle that has such a cell among its (2 nsize+1)^dim targets.  torch.where / torch
    division reproduce this exactly; the Styler loop then zeroes those variables through
    ``nan_to_num`` (``styler_3p.py:360``).
    """
    res = [int(r) for r in res]
    B = p.shape[0]
    _, valid, cell, idx, r = _cell_setup(p, domain, res, clip, eps)
    h = radius * support
    nch = x.shape[-1]
    vol = int(np.prod(res))
    wmap = torch.zeros(B * vol, 1, dtype=p.dtype)
    num = torch.zeros(B * vol, nch, dtype=p.dtype)
    rng = range(-nsize, nsize + 1)
    shifts = [(a, b) for a in rng for b in rng] if is_2d else \
             [(a, b, c) for a in rng for b in rng for c in rng]
    for s in shifts:
        d = r - torch.as_tensor(s, dtype=p.dtype) * cell
        q = _safe_norm((d ** 2).sum(-1)) / h
        w = cubic_w(q, h, is_3d=not is_2d)
        _scatter(wmap, res, idx, s, w.unsqueeze(-1), valid)
        _scatter(num, res, idx, s, w.unsqueeze(-1) * x, valid)
    flip = [1 if is_2d else 2]
    wmap = torch.flip(wmap.reshape([B] + res + [1]), dims=flip)
    num = torch.flip(num.reshape([B] + res + [nch]), dims=flip)
    return torch.where(wmap > eps, num / wmap, num)


# --------------------------------------------------------------------------------------
# warps: rotate / advect
# --------------------------------------------------------------------------------------
def linspace_tf(n, dtype, low=-1.0, high=1.0):
    """tf.linspace as TF-1.15's CPU/GPU kernel computes it: start + step*i in T
    (``transform.py:175``)."""
    if n == 1:
        return torch.full((1,), low, dtype=dtype)
    step = torch.tensor((high - low), dtype=dtype) / torch.tensor(float(n - 1), dtype=dtype)
    return torch.tensor(low, dtype=dtype) + step * torch.arange(n, dtype=dtype)


def mgrid(*lens, dtype=torch.float32):
    """``transform.py:152-177``: stacked 'ij' meshgrid of linspace(-1,1,len)."""
    coords = [linspace_tf(n, dtype) for n in lens]
    return torch.stack(torch.meshgrid(*coords, indexing='ij'))


def interpolate(img, coords):
    """Edge-clamped multilinear sampling, ``transform.py:280-341`` (2-D) / ``343-433`` (3-D).

    img [B,*lens,C]; coords: list of dim tensors [B,M] in [-1,1] units.  Returns [B,M,C].
    floor / floor+1 are clamped to [0,len-1]; weights use the *unclamped* coordinate minus
    the *clamped* lower index.  Gradient reaches ``img`` (gather -> scatter-add) and the
    coordinates (through the weights), as in TF.
    """
    B = img.shape[0]
    lens = list(img.shape[1:-1])
    C = img.shape[-1]
    dim = len(lens)
    flat = img.reshape(B, -1, C)
    lo, hi, frac = [], [], []
    for a in range(dim):
        L = lens[a]
        x = (coords[a] + 1.0) * (float(L) - 1.0) * 0.5
        x0 = torch.floor(x).to(torch.int64)
        x1 = x0 + 1
        x0 = x0.clamp(0, L - 1)
        x1 = x1.clamp(0, L - 1)
        lo.append(x0)
        hi.append(x1)
        frac.append(x - x0.to(x.dtype))
    out = 0
    for corner in range(1 << dim):
        idx = torch.zeros_like(lo[0])
        w = torch.ones_like(frac[0])
        for a in range(dim):
            bit = (corner >> (dim - 1 - a)) & 1
            idx = idx * lens[a] + (hi[a] if bit else lo[a])
            w = w * (frac[a] if bit else (1.0 - frac[a]))
        val = torch.gather(flat, 1, idx.unsqueeze(-1).expand(-1, -1, C))
        out = out + w.unsqueeze(-1) * val
    return out


def rotate(d, rot_mats):
    """``transform.py:611-628``: tile d over the rotations, rotate the [-1,1]^3 grid by R
    and resample.  d [B,D,H,W,C], rot_mats [n_rot,3,3] -> [B*n_rot,D,H,W,C].

    Follows tf.tile ordering: d tiled n_rot times (so output b' uses d[b' % B]) and R tiled
    B times (output b' uses R[b' % n_rot]); identical for the reference's B == 1.
    """
    B, D, H, W, C = d.shape
    R = torch.as_tensor(np.asarray(rot_mats), dtype=d.dtype).reshape(-1, 3, 3)
    n_rot = R.shape[0]
    nb = B * n_rot
    dd = d.repeat(n_rot, 1, 1, 1, 1)
    rr = R.repeat(B, 1, 1)
    g = mgrid(D, H, W, dtype=d.dtype).reshape(1, 3, -1).expand(nb, -1, -1)
    g = torch.matmul(rr, g)
    out = interpolate(dd, [g[:, 0], g[:, 1], g[:, 2]])
    return out.reshape(nb, D, H, W, C)


def advect(d, vel, is_3d=False):
    """Semi-Lagrangian back-trace, order 1 only, ``transform.py:557-609``.

    d [1,X,Y,(Z),C]; vel [1,X,Y,(Z),dim] in normalised [-1,1] units, channel i <-> axis i.
    The MacCormack branch (order 2) is non-functional at HEAD and is not restated.
    """
    lens = list(d.shape[1:-1])
    g = mgrid(*lens, dtype=d.dtype).unsqueeze(0)
    perm = [0, len(lens) + 1] + list(range(1, len(lens) + 1))
    g = g - vel.permute(*perm)
    g = g.reshape(1, len(lens), -1)
    out = interpolate(d, [g[:, a] for a in range(len(lens))])
    return out.reshape(d.shape)


# --------------------------------------------------------------------------------------
# host-side view sampling
# --------------------------------------------------------------------------------------
def rot_z_3d(deg):
    """``transform.py:640-648`` -- acts on axes (0,1) = (D,H)."""
    a = deg / 180.0 * np.pi
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def rot_y_3d(deg):
    """``transform.py:650-658`` -- acts on axes (0,2) = (D,W)."""
    a = deg / 180.0 * np.pi
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]])


def rot_mat_uniform(phi0, phi1, phi_unit, theta0, theta1, theta_unit):
    """``transform.py:750-768`` (the float linspace count at ``:754-755`` is cast to int)."""
    if phi_unit == 0:
        phi = [(phi1 - phi0) / 2]
    else:
        phi = np.linspace(phi0, phi1, int(np.abs(phi1 - phi0) / float(phi_unit) + 1), endpoint=True)
    if theta_unit == 0:
        theta = [(theta1 - theta0) / 2]
    else:
        theta = np.linspace(theta0, theta1, int(np.abs(theta1 - theta0) / float(theta_unit) + 1),
                            endpoint=True)
    return [{'phi': a, 'theta': b} for a in phi for b in theta]


class _Poisson:
    """Bridson Poisson-disc sampler with the rng call order of ``transform.py:14-150``."""
    _NB = [(-1, -2), (0, -2), (1, -2), (-2, -1), (-1, -1), (0, -1), (1, -1), (2, -1), (-2, 0),
           (-1, 0), (1, 0), (2, 0), (-2, 1), (-1, 1), (0, 1), (1, 1), (2, 1), (-1, 2), (0, 2),
           (1, 2), (0, 0)]

    def __init__(self, rng, width, height, r, k=30):
        self.rng, self.w, self.h, self.r, self.k = rng, width, height, r, k
        self.a = r / np.sqrt(2)
        self.nx, self.ny = int(width / self.a) + 1, int(height / self.a) + 1
        self.cells = {}
        self.pts = []

    def _cc(self, pt):
        return int(pt[0] // self.a), int(pt[1] // self.a)

    def _ok(self, pt):
        cx, cy = self._cc(pt)
        for dx, dy in self._NB:
            key = (cx + dx, cy + dy)
            if not (0 <= key[0] < self.nx and 0 <= key[1] < self.ny):
                continue
            j = self.cells.get(key)
            if j is not None:
                o = self.pts[j]
                if (o[0] - pt[0]) ** 2 + (o[1] - pt[1]) ** 2 < self.r ** 2:
                    return False
        return True

    def _near(self, ref):
        i = 0
        while i < self.k:
            rho = self.rng.uniform(self.r, 2 * self.r)
            th = self.rng.uniform(0, 2 * np.pi)
            pt = ref[0] + rho * np.cos(th), ref[1] + rho * np.sin(th)
            if not (0 < pt[0] < self.w and 0 < pt[1] < self.h):
                continue
            if self._ok(pt):
                return pt
            i += 1
        return False

    def sample(self):
        pt = (self.rng.uniform(0, self.w), self.rng.uniform(0, self.h))
        self.pts = [pt]
        self.cells[self._cc(pt)] = 0
        active = [0]
        while active:
            j = self.rng.choice(active)
            new = self._near(self.pts[j])
            if new:
                self.pts.append(new)
                active.append(len(self.pts) - 1)
                self.cells[self._cc(new)] = len(self.pts) - 1
            else:
                active.remove(j)
        return self.pts


def rot_mat_poisson(phi0, phi1, phi_unit, theta0, theta1, theta_unit, rng):
    """``transform.py:724-748``."""
    if phi_unit == 0:
        h, phi0 = 1, -0.5
    else:
        h = phi1 - phi0
    if theta_unit == 0:
        w, theta0 = 1, -0.5
    else:
        w = theta1 - theta0
    r = max(phi_unit, theta_unit) / 2
    pts = _Poisson(rng, width=w, height=h, r=r).sample()
    return [{'phi': s[1] + phi0, 'theta': s[0] + theta0} for s in pts]


def rot_mat(phi0, phi1, phi_unit, theta0, theta1, theta_unit, sample_type='uniform', rng=None,
            nv=None):
    """``transform.py:689-722``: views -> R = R_y(theta) R_z(phi)."""
    if 'uniform' in sample_type:
        views = rot_mat_uniform(phi0, phi1, phi_unit, theta0, theta1, theta_unit)
    else:
        if 'poisson' in sample_type:
            pu, tu = phi_unit, theta_unit
            views = rot_mat_poisson(phi0, phi1, pu, theta0, theta1, tu, rng)
            views += rot_mat_uniform(phi0, phi1, 0, theta0, theta1, 0)
        else:  # both
            pu, tu = phi_unit * 2, theta_unit * 2
            views = rot_mat_uniform(phi0, phi1, phi_unit, theta0, theta1, theta_unit)
            views += rot_mat_poisson(phi0, phi1, pu, theta0, theta1, tu, rng)
        if nv is not None:
            if len(views) > nv:
                views = views[len(views) - nv:]
            elif len(views) < nv:
                extra = rot_mat_poisson(phi0, phi1, pu, theta0, theta1, tu, rng)
                views += extra[:nv - len(views)]
    mats = [np.matmul(rot_y_3d(v['theta']), rot_z_3d(v['phi'])) for v in views]
    return mats, views


# --------------------------------------------------------------------------------------
# grid -> particle gathers (resimulation / data-prep stage)
# --------------------------------------------------------------------------------------
def _g2p_axis(pos, length, taps):
    """Clamped tap indices + fractional offset along one axis (``transform.py:796-838`` cubic,
    ``:1128-1162`` linear).  The reference re-assigns x1 (x0 for linear) by ``clip_by_value``
    BEFORE computing ``dx = x - (x1 + 0.5)`` (``:999``, ``:1200``): the offset is measured from the
    clamped anchor, so particles outside the grid extrapolate.  Restated as is."""
    x = pos * float(length)
    f = torch.floor(x - 0.5).to(torch.int64)
    first = f - 1 if taps == 4 else f
    idx = [torch.clamp(first + i, 0, length - 1) for i in range(taps)]
    anchor = idx[1] if taps == 4 else idx[0]
    return idx, x - (anchor.to(pos.dtype) + 0.5)


def _hermite(A, B, C, D, t):
    """``transform.py:972-978``."""
    a = A * (-0.5) + B * 1.5 + C * (-1.5) + D * 0.5
    b = A + B * (-2.5) + C * 2.0 + D * (-0.5)
    c = A * (-0.5) + C * 0.5
    return a * t * t * t + b * t * t + c * t + B


def g2p_cubic(g, p, is_2d=True):
    """Catmull-Rom grid->particle sampling at cell centres, ``transform.py:778-1108``.
    g [1,n0,n1,(n2),C], p [1,N,dim] normalised in the grid's axis order -> [1,N,C]."""
    dims = list(g.shape[1:-1])
    C = g.shape[-1]
    flat = g.reshape(-1, C)
    ix, tx = _g2p_axis(p[0, :, 0], dims[0], 4)
    iy, ty = _g2p_axis(p[0, :, 1], dims[1], 4)
    tx, ty = tx[:, None], ty[:, None]
    if is_2d:
        rows = [_hermite(*[flat[ix[a] * dims[1] + iy[b]] for a in range(4)], tx) for b in range(4)]
        out = _hermite(*rows, ty)
    else:
        iz, tz = _g2p_axis(p[0, :, 2], dims[2], 4)
        tz = tz[:, None]
        planes = []
        for e in range(4):
            rows = [_hermite(*[flat[(ix[a] * dims[1] + iy[b]) * dims[2] + iz[e]] for a in range(4)], tx)
                    for b in range(4)]
            planes.append(_hermite(*rows, ty))
        out = _hermite(*planes, tz)
    return out[None]


def g2p_linear(g, p, is_2d=True):
    """Bi/tri-linear grid->particle sampling at cell centres, ``transform.py:1110-1231``."""
    dims = list(g.shape[1:-1])
    C = g.shape[-1]
    flat = g.reshape(-1, C)
    ix, tx = _g2p_axis(p[0, :, 0], dims[0], 2)
    iy, ty = _g2p_axis(p[0, :, 1], dims[1], 2)
    out = 0
    if is_2d:
        for a in range(2):
            for b in range(2):
                w = (tx if a else 1.0 - tx) * (ty if b else 1.0 - ty)
                out = out + w[:, None] * flat[ix[a] * dims[1] + iy[b]]
    else:
        iz, tz = _g2p_axis(p[0, :, 2], dims[2], 2)
        for a in range(2):
            for b in range(2):
                for e in range(2):
                    w = (tx if a else 1.0 - tx) * (ty if b else 1.0 - ty) * (tz if e else 1.0 - tz)
                    out = out + w[:, None] * flat[(ix[a] * dims[1] + iy[b]) * dims[2] + iz[e]]
    return out[None]


def g2p(g, p, is_2d=True, is_linear=False):
    """``transform.py:771-776``."""
    return g2p_linear(g, p, is_2d) if is_linear else g2p_cubic(g, p, is_2d)
