"""Host-side view sampling (the device-side part of the reference's ``transform.py`` lives in
``csrc/``).  Mirrors ``rot_mat`` / ``rot_mat_uniform`` / ``rot_mat_poisson`` and the
Poisson-disc sampler (``transform.py:14-150, 640-768``): same view lists, same numpy-RNG
consumption order, so a seeded run draws the same views as the reference.
"""
import math

import numpy as np


def rot_z_3d(deg):
    """Rotation in the (D,H) plane (``transform.py:640-648``)."""
    c, s = math.cos(math.radians(deg)), math.sin(math.radians(deg))
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def rot_y_3d(deg):
    """Rotation in the (D,W) plane (``transform.py:650-658``)."""
    c, s = math.cos(math.radians(deg)), math.sin(math.radians(deg))
    return np.array([[c, 0.0, -s], [0.0, 1.0, 0.0], [s, 0.0, c]])


def _axis_samples(lo, hi, unit):
    if unit == 0:
        return [(hi - lo) / 2]
    return list(np.linspace(lo, hi, int(abs(hi - lo) / float(unit) + 1), endpoint=True))


def rot_mat_uniform(phi0, phi1, phi_unit, theta0, theta1, theta_unit):
    """Lattice of views, phi-major (``transform.py:750-768``)."""
    return [{'phi': ph, 'theta': th}
            for ph in _axis_samples(phi0, phi1, phi_unit)
            for th in _axis_samples(theta0, theta1, theta_unit)]


class PoissonDisc:
    """Bridson's algorithm on a background grid of side r/sqrt(2).  Draw order per candidate:
    radius ~ U(r,2r) then angle ~ U(0,2pi); reference point picked with ``rng.choice``."""

    _RING = [(dx, dy) for dy in range(-2, 3) for dx in range(-2, 3) if abs(dx) + abs(dy) < 4]

    def __init__(self, rng, width=50, height=50, r=1, k=30):
        self.rng, self.width, self.height, self.r, self.k = rng, width, height, r, k
        self.a = r / math.sqrt(2)
        self.nx, self.ny = int(width / self.a) + 1, int(height / self.a) + 1
        self.occ = {}
        self.samples = []

    def _cell(self, pt):
        return int(pt[0] // self.a), int(pt[1] // self.a)

    def _far_enough(self, pt):
        cx, cy = self._cell(pt)
        r2 = self.r * self.r
        for dx, dy in self._RING:
            x, y = cx + dx, cy + dy
            if 0 <= x < self.nx and 0 <= y < self.ny and (x, y) in self.occ:
                q = self.samples[self.occ[(x, y)]]
                if (q[0] - pt[0]) ** 2 + (q[1] - pt[1]) ** 2 < r2:
                    return False
        return True

    def _candidate(self, ref):
        tries = 0
        while tries < self.k:
            rho = self.rng.uniform(self.r, 2 * self.r)
            ang = self.rng.uniform(0, 2 * np.pi)
            pt = (ref[0] + rho * np.cos(ang), ref[1] + rho * np.sin(ang))
            if not (0 < pt[0] < self.width and 0 < pt[1] < self.height):
                continue          # outside: redraw without consuming a try
            if self._far_enough(pt):
                return pt
            tries += 1
        return None

    def sample(self):
        first = (self.rng.uniform(0, self.width), self.rng.uniform(0, self.height))
        self.samples = [first]
        self.occ = {self._cell(first): 0}
        active = [0]
        while active:
            i = self.rng.choice(active)
            pt = self._candidate(self.samples[i])
            if pt is None:
                active.remove(i)
                continue
            self.samples.append(pt)
            active.append(len(self.samples) - 1)
            self.occ[self._cell(pt)] = len(self.samples) - 1
        return self.samples


def rot_mat_poisson(phi0, phi1, phi_unit, theta0, theta1, theta_unit, rng):
    """Blue-noise views in the (theta, phi) rectangle (``transform.py:724-748``)."""
    h, w = phi1 - phi0, theta1 - theta0
    if phi_unit == 0:
        h, phi0 = 1, -0.5
    if theta_unit == 0:
        w, theta0 = 1, -0.5
    pts = PoissonDisc(rng, width=w, height=h, r=max(phi_unit, theta_unit) / 2).sample()
    return [{'phi': y + phi0, 'theta': x + theta0} for x, y in pts]


def rot_mat(phi0, phi1, phi_unit, theta0, theta1, theta_unit, sample_type='uniform', rng=None, nv=None):
    """Views and their matrices R = R_y(theta) R_z(phi) (``transform.py:689-722``)."""
    box = (phi0, phi1, phi_unit, theta0, theta1, theta_unit)
    if 'uniform' in sample_type:
        views = rot_mat_uniform(*box)
    else:
        if 'poisson' in sample_type:
            pbox = box
            views = rot_mat_poisson(*pbox, rng) + rot_mat_uniform(phi0, phi1, 0, theta0, theta1, 0)
        else:
            pbox = (phi0, phi1, phi_unit * 2, theta0, theta1, theta_unit * 2)
            views = rot_mat_uniform(*box) + rot_mat_poisson(*pbox, rng)
        if nv is not None and len(views) > nv:
            views = views[len(views) - nv:]
        elif nv is not None and len(views) < nv:
            views = views + rot_mat_poisson(*pbox, rng)[:nv - len(views)]
    mats = [rot_y_3d(v['theta']) @ rot_z_3d(v['phi']) for v in views]
    return mats, views
