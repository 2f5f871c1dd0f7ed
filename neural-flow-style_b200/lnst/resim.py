"""Resimulation (data-prep) step -- drop-in for ``test_smokegun_resim.SimG2P`` (reference
``test_smokegun_resim.py:17-217``): turns a grid simulation (density d_t, velocity u_t) into the
particle sets (positions + multi-scale densities) that ``styler_3p.Styler.run`` stylises.

Per frame (``optimize``, ``:167-217``)
  1. RK4-advect the particles through u_t                         -> ``lnst_rk4_advect`` (one kernel)
  2. ``iter`` Adam steps on a per-particle displacement against the SPH pressure loss
     mean(where(d_rec > 0, d_rec - rho0, 0)^2)                     -> ``lnst_splat_sph_fwd`` + ``lnst_pressure_loss``
                                                                     + ``lnst_splat_sph_bwd_pos`` + ``lnst_adam_step_dev``
  3. seed new particles in the source window where the advected set does not cover d_t (host, like the reference)
  4. sample the density at the particles, one residual octave at a time
                                                                   -> ``lnst_g2p`` + ``lnst_splat_wavg_*`` + ``lnst_sub_fliph``

Everything stays on the device between the steps; the host sees the loss list, d_diff (for the seeding
``np.where``) and the results.  No CPU fallback: the ops raise when the CUDA library is missing.
"""
import numpy as np
import torch

from . import _lib, ops

f32 = torch.float32

REF_SRC_REGION = ((76, 124), (231, 279), (16, 64))     # test_smokegun_resim.py:117-119 (200x300x200 demo grid)


class SimG2P(object):
    def __init__(self, self_dict, device=None, src_region=None):
        # get arguments (test_smokegun_resim.py:19-21)
        for arg in vars(self_dict):
            setattr(self, arg, getattr(self_dict, arg))
        lib = _lib.get()
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device()) if lib.kind == 'cuda' else torch.device('cpu')
        self.device = torch.device(device)
        self.src_region = src_region if src_region is not None else getattr(self, 'src_region', REF_SRC_REGION)
        self.time_step = 0.5                                           # :53
        self.grid = _lib.make_grid(3, self.resolution, self.domain, self.nsize, False)
        self.mass = 0.8 * (2 * self.radius) ** 3 * self.rest_density   # transform.py:1348-1352
        self.h_sph = self.radius * 4                                   # p2g(..., support=4), :66
        cells = int(np.prod(self.resolution))
        self._d_rec = torch.empty([int(r) for r in self.resolution], dtype=f32, device=self.device)
        self._g_d = torch.empty_like(self._d_rec)
        self._num = torch.empty(1, cells, dtype=f32, device=self.device)
        self._loss = torch.zeros(1, dtype=f32, device=self.device)

    # ---- helpers ----------------------------------------------------------------------------------
    def _t(self, a):
        return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(self.device)

    def _wavg(self, x, r, support):
        """p2g_wavg(x, r, ..., clip=False, support) for moving particles: the weight map is recomputed."""
        hs = [self.radius * support]
        wmap = ops.splat_wavg_wmap(x, self.grid, hs)
        out = torch.empty_like(self._d_rec)
        ops.splat_wavg_fwd(x, r, None, self.grid, hs, wmap, self._num, out)
        return out

    def _multiscale(self, x_hat, d):
        """:82-106 -- returns r_smp [N,octave_n], d_hat [D,H,W] (H stored flipped like every splat), d_diff."""
        r, d_hat = [], None
        for o in range(self.octave_n):
            d_ = ops.sub_fliph(d, d_hat) if o > 0 else d                  # d - d_hi[:,:,::-1]
            r_ = ops.g2p(d_.unsqueeze(-1), x_hat)
            r.append(r_)
            new = self._wavg(x_hat, r_, self.support / self.octave_scale ** o)
            d_hat = new.add_(d_hat) if o > 0 else new
        d_diff = ops.sub_fliph(d, d_hat)                                  # (d[:,:,::-1] - d_hat)[0,:,::-1,:,0]
        return torch.cat(r, dim=-1), d_hat, d_diff

    # ---- reference API ----------------------------------------------------------------------------
    def sample(self, d, disc=1, threshold=0, p0=None, p_id=None):
        """sample particles where d's value is higher than threshold (:110-153), inside the source window"""
        (z0, z1), (y0, y1), (x0, x1) = self.src_region
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
        p = np.stack([p[:, 0] / d.shape[0], p[:, 1] / d.shape[1], p[:, 2] / d.shape[2]], axis=-1)   # [0,1]

        # if there are new particles, add to prev
        if len(p) > 0:
            if p_id is None:
                p_id = np.arange(p.shape[0])
            else:
                p_id0 = p_id[-1] + 1
                p_id = np.concatenate([p_id, np.arange(p_id0, p_id0 + p.shape[0])])
            if p0 is not None:
                p = np.concatenate([p0, p], axis=0)
        return p, p_id

    def advect(self, p, u):
        """x_adv of the graph (:36-55); p, u device tensors."""
        return ops.rk4_advect(u, p, self.time_step)

    def naive_adv(self, p, u, r):
        """reconstruct density field from p_t' with r (:155-165)"""
        p_adv = self.advect(self._t(p), self._t(u))
        d_rec = self._wavg(p_adv, self._t(r), 4)
        return p_adv.cpu().numpy(), d_rec.cpu().numpy()

    def optimize(self, p, p_id, d, u):
        """advect p_t with u_t, redistribute, seed, sample densities (:167-217)"""
        p = self.advect(self._t(p), self._t(u))
        d = self._t(d)

        # optimize for particle redistribution: variable = displacement, fresh Adam slots per frame (:182-184)
        var = torch.zeros_like(p)
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        state = torch.tensor([0.9, 0.999, 0.0], dtype=f32).to(self.device)
        n_it = int(self.iter)
        losses = torch.zeros(max(n_it, 1), dtype=f32, device=self.device)
        for it in range(n_it):
            loss = losses[it:it + 1]
            ops.splat_sph_fwd(p, var, self.grid, self.h_sph, self.mass, out=self._d_rec)
            ops.pressure_loss(self._d_rec, self.rest_density, 1.0, loss, self._g_d)
            grad = ops.splat_sph_bwd_pos(p, var, self.grid, self.h_sph, self.mass, self._g_d)
            ops.adam_step_dev(var, grad, m, v, state, self.lr)
        l = [float(x) for x in losses[:n_it].cpu().numpy()]            # one read for the whole loop

        # seed particles
        x_hat = p + var
        _, _, d_diff = self._multiscale(x_hat, d)
        d_diff = d_diff.cpu().numpy()
        p, p_id = self.sample(d_diff, disc=self.disc, threshold=self.threshold, p0=x_hat.cpu().numpy(), p_id=p_id)

        # sample density at new position
        r_smp, d_hat, _ = self._multiscale(self._t(p), d)
        return {
            'p': p,
            'p_id': p_id,
            'p_den': r_smp.cpu().numpy(),
            'l': l,
            'd_diff': np.mean(d_diff, axis=0),                          # for debug (:212-215)
            'd_smp': torch.clamp(d_hat, 0, 1).cpu().numpy(),
        }
