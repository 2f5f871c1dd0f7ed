"""Oracle restatement of the resimulation (data-prep) step, ``test_smokegun_resim.py:17-217``
(class ``SimG2P``): RK4 particle advection through the velocity grid, Adam optimisation of a
per-particle displacement against the SPH pressure loss, seeding of new particles where the advected
set does not cover the density, multi-scale density sampling.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  PARITY: pinned by ``tests/golden/ref_resim.npz``
(the reference's own ``SimG2P`` graph and ``optimize`` / ``naive_adv`` methods run on ``oracle/tfshim``;
only ``sample``'s hard-coded source window ``d[76:124,231:279,16:64]`` (``:117``) is parameterised,
because it only fits the 200x300x200 demo grid).
"""
import numpy as np
import torch

from . import transform as T
from .adam import TFAdam

REF_SRC_REGION = ((76, 124), (231, 279), (16, 64))     # test_smokegun_resim.py:117-119


def sample(d, disc=1, threshold=0, p0=None, p_id=None, src_region=REF_SRC_REGION):
    """``SimG2P.sample`` (``:110-153``): one particle per sub-cell of every source-window voxel with
    d > threshold; ids continue from the previous maximum; normalised (z,y,x) in [0,1]."""
    (z0, z1), (y0, y1), (x0, x1) = src_region
    pid = np.where(d[z0:z1, y0:y1, x0:x1] > threshold)
    pid = np.array(pid).transpose([1, 0]).astype(np.float64)
    pid += np.array([z0, y0, x0])
    cell_size = 1 / disc
    offset = cell_size / 2
    p = []
    for i in range(disc):
        for j in range(disc):
            for k in range(disc):
                p.append(pid + offset + np.array([cell_size * i, cell_size * j, cell_size * k]))
    p = np.concatenate(p, axis=0)
    p = np.stack([p[:, 0] / d.shape[0], p[:, 1] / d.shape[1], p[:, 2] / d.shape[2]], axis=-1)
    if len(p) > 0:
        if p_id is None:
            p_id = np.arange(p.shape[0])
        else:
            p_id0 = p_id[-1] + 1
            p_id = np.concatenate([p_id, np.arange(p_id0, p_id0 + p.shape[0])])
        if p0 is not None:
            p = np.concatenate([p0, p], axis=0)
    return p, p_id


class OracleSimG2P:
    def __init__(self, config, dtype=torch.float32, src_region=REF_SRC_REGION):
        self.c = config
        self.dt = dtype
        self.src_region = src_region

    def _t(self, a):
        return torch.as_tensor(np.asarray(a), dtype=self.dt)

    def advect(self, p, u, time_step=0.5):
        """``:36-55``: RK4 velocity sampling, x_adv = x + v * 0.5."""
        x, u = self._t(p)[None], self._t(u)[None]
        v = T.g2p(u, x, is_2d=False)
        v1 = T.g2p(u, x + v * 0.5, is_2d=False)
        v2 = T.g2p(u, x + v1 * 0.5, is_2d=False)
        v3 = T.g2p(u, x + v2, is_2d=False)
        v = (v + v1 * 2 + v2 * 2 + v3) / 6
        return (x + v * time_step)[0]

    def pressure_loss(self, x_hat):
        """``:66-74``."""
        c = self.c
        d_rec = T.p2g(x_hat, c.domain, c.resolution, c.radius, c.rest_density, c.nsize, is_2d=False,
                      clip=False, support=4)
        pressure = torch.where(d_rec > 0, d_rec - c.rest_density, torch.zeros_like(d_rec))
        return (pressure ** 2).mean()

    def multiscale(self, x_hat, d):
        """``:82-106``: per-octave residual sampling; returns r_smp [N,octave_n], d_smp, d_diff."""
        c = self.c
        d = self._t(d)[None, ..., None]
        r, d_hat = [], None
        for o in range(c.octave_n):
            d_ = d - torch.flip(d_hat, dims=[2]) if o > 0 else d
            r_ = T.g2p(d_, x_hat, is_2d=False)
            r.append(r_)
            new = T.p2g_wavg(x_hat, r_, c.domain, c.resolution, c.radius, c.nsize, is_2d=False, clip=False,
                             support=c.support / c.octave_scale ** o)
            d_hat = new + d_hat if o > 0 else new
        r_smp = torch.cat(r, dim=-1)[0]
        d_smp = torch.clamp(d_hat[0, ..., 0], 0, 1)
        d_diff = torch.flip((torch.flip(d, dims=[2]) - d_hat)[0, ..., 0], dims=[1])
        return r_smp, d_smp, d_diff

    def naive_adv(self, p, u, r):
        """``:155-165``."""
        c = self.c
        p_adv = self.advect(p, u)
        d_rec = T.p2g_wavg(p_adv[None], self._t(r)[None], c.domain, c.resolution, c.radius, c.nsize, is_2d=False,
                           clip=False, support=4)
        return p_adv.numpy(), d_rec[0, ..., 0].numpy()

    def optimize(self, p, p_id, d, u):
        """``:167-217``."""
        c = self.c
        p = self.advect(p, u)
        var = torch.zeros_like(p)
        opt = TFAdam()
        losses = []
        for _ in range(c.iter):
            v = var.clone().requires_grad_(True)
            loss = self.pressure_loss((p + v)[None])
            g, = torch.autograd.grad(loss, v)
            losses.append(float(loss.detach()))
            var = opt.step(var, g, c.lr)
        x_hat = (p + var)[None]
        _, _, d_diff = self.multiscale(x_hat, d)
        p_new, p_id = sample(d_diff.numpy(), disc=c.disc, threshold=c.threshold, p0=x_hat[0].numpy(), p_id=p_id,
                             src_region=self.src_region)
        r_smp, d_smp, _ = self.multiscale(self._t(p_new)[None], d)
        return {'p': p_new, 'p_id': p_id, 'p_den': r_smp.numpy(), 'l': losses,
                'd_diff': np.mean(d_diff.numpy(), axis=0), 'd_smp': d_smp.numpy()}
