#!/usr/bin/env python
"""Golden vectors for the resimulation (data-prep) path FROM THE REFERENCE ITSELF, run here.

    python tests/golden/make_resim_golden.py            (needs /root/reference)

Same arrangement as ``make_reference_golden.py``: ``oracle/tfshim`` stands in for TensorFlow 1.15 and the
UNMODIFIED reference modules are imported from ``/root/reference`` -- ``transform.g2p`` (cubic and linear,
2-D and 3-D) and ``test_smokegun_resim.SimG2P`` (graph construction in ``__init__``, ``optimize`` and
``naive_adv`` are the reference's own code).  Two things are added around it: a stub for the absent
``partio`` module (only the driver's file output uses it), and ``SimG2P.sample``'s source window -- the
reference hard-codes ``d[76:124,231:279,16:64]`` for its 200x300x200 demo grid
(``test_smokegun_resim.py:117-119``); the scaled-down case needs a window that fits its 12x16x10 grid, so
the subclass below repeats ``sample`` with that one line parameterised.

Output: ``tests/golden/ref_resim.npz``; ``tests/test_widen_resim.py`` holds the oracle (CPU) and the CUDA path
(``-m gpu``) to it.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as M  # noqa: E402


def g2p_inputs():
    rng = np.random.RandomState(77)
    d = {}
    d['g3'] = rng.randn(6, 7, 5, 3).astype(np.float32)
    p3 = rng.uniform(-0.08, 1.08, (150, 3)).astype(np.float32)          # some particles outside the grid
    p3[:4] = [[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5], [-1, -1, -1]]       # faces, centre, the drivers' padding row
    d['p3'] = p3
    d['g2'] = rng.randn(9, 6, 2).astype(np.float32)
    p2 = rng.uniform(-0.08, 1.08, (90, 2)).astype(np.float32)
    p2[:3] = [[0, 0], [1, 1], [0.25, 0.75]]
    d['p2'] = p2
    return d


def resim_config():
    """test_smokegun_resim.py:main (:336-372), scaled down: domain == resolution (unit cells), radius .5."""
    import argparse
    c = argparse.Namespace()
    c.resolution = [12, 16, 10]
    c.domain = [12, 16, 10]
    c.scale = 1
    c.disc = 1
    c.radius = 0.5
    c.nsize = 1
    c.support = 4
    c.rest_density = 1000
    c.threshold = 0.01
    c.lr = 0.0005
    c.iter = 5
    c.transmit = 0.01
    c.octave_n = 2
    c.octave_scale = 2
    c.seed = 123
    return c


SRC_REGION = ((2, 6), (9, 14), (3, 8))


def resim_inputs(n_frames=3):
    """Seeded smooth density blob rising through a smooth velocity field: d [T,D,H,W] in [0,1],
    u [T,D,H,W,3] in normalised units per frame, (z,y,x) channel order (:246-250)."""
    c = resim_config()
    D, H, W = c.resolution
    rng = np.random.RandomState(5)
    z, y, x = np.meshgrid((np.arange(D) + .5) / D, (np.arange(H) + .5) / H, (np.arange(W) + .5) / W, indexing='ij')
    ds, us = [], []
    for t in range(n_frames):
        cy = 0.68 - 0.05 * t
        blob = np.exp(-(((z - 0.35) / 0.16) ** 2 + ((y - cy) / 0.17) ** 2 + ((x - 0.55) / 0.2) ** 2))
        ds.append((blob * (blob > 0.05)).astype(np.float32))
        u = np.stack([0.02 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y),
                      -0.09 + 0.03 * np.cos(2 * np.pi * z),
                      0.025 * np.sin(2 * np.pi * y + t)], axis=-1)
        us.append((u + 0.004 * rng.randn(D, H, W, 3)).astype(np.float32))
    return c, np.stack(ds), np.stack(us)


def run():
    M._setup_paths()
    sys.modules.setdefault('partio', types.ModuleType('partio'))       # only run()'s .bgeo output uses it
    import tensorflow as tf
    assert bool(os.environ.get('LNST_REAL_TF')) or 'shim' in tf.__version__
    import transform as RT
    out = {}
    sess = tf.Session()
    I = g2p_inputs()
    for dim, g, p in ((3, I['g3'], I['p3']), (2, I['g2'], I['p2'])):
        gp = tf.placeholder(tf.float32, [1] + [None] * dim + [g.shape[-1]])
        pp = tf.placeholder(tf.float32, [1, None, dim])
        for lin in (False, True):
            y = RT.g2p(gp, pp, is_2d=dim == 2, is_linear=lin)
            out['g2p%d_%s' % (dim, 'linear' if lin else 'cubic')] = sess.run(y, {gp: g[None], pp: p[None]})

    import test_smokegun_resim as RS

    class Sim(RS.SimG2P):
        def sample(self, d, disc=1, threshold=0, p0=None, p_id=None):
            # test_smokegun_resim.py:110-153 verbatim except the source window (:117-119)
            (z0, z1), (y0, y1), (x0, x1) = SRC_REGION
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
            pz, py, px = p[:, 0], p[:, 1], p[:, 2]
            pz /= d.shape[0]
            py /= d.shape[1]
            px /= d.shape[2]
            p = np.stack([pz, py, px], axis=-1)
            if len(p) > 0:
                if p_id is None:
                    p_id = np.arange(p.shape[0])
                else:
                    p_id0 = p_id[-1] + 1
                    p_id = np.concatenate([p_id, np.arange(p_id0, p_id0 + p.shape[0])])
                if p0 is not None:
                    p = np.concatenate([p0, p], axis=0)
            return p, p_id

    c, ds, us = resim_inputs()
    sim = Sim(c)
    p, p_id = sim.sample(ds[0], disc=c.disc, threshold=0)
    out['p0'] = p
    for t in range(ds.shape[0]):                     # the driver's loop, test_smokegun_resim.py:229-268
        res = sim.optimize(p, p_id, ds[t], us[t])
        p, p_id = res['p'], res['p_id']
        for k in ('p', 'p_id', 'p_den', 'l', 'd_diff', 'd_smp'):
            out['f%d_%s' % (t, k)] = np.asarray(res[k])
        print('frame', t, 'particles', p.shape[0], 'loss', res['l'][0], '->', res['l'][-1], flush=True)
    # naive advection branch (:270-283)
    p_adv, d_rec = sim.naive_adv(out['p0'], us[0], np.ones([out['p0'].shape[0], 1]))
    out['naive_p'], out['naive_d'] = p_adv, d_rec
    return out


if __name__ == '__main__':
    o = run()
    np.savez_compressed(os.path.join(HERE, 'ref_resim.npz'), **o)
    print(sorted(o))
