"""Grid simulation -> particle sets (the data the smoke stylisation reads) -- mirrors reference
``test_smokegun_resim.py:219-372``."""
import os

import numpy as np

from .. import partio
from ..config import get_config
from ..resim import SimG2P
from ..util import prepare_dirs_and_logger
from . import save_png


def centred_velocity(v_):
    """mantaflow MAC grid [D,H,W,3] (x,y,z components on the low faces) -> cell-centred, H flipped
    (test_smokegun_resim.py:235-245)"""
    vx = np.dstack((v_, np.zeros((v_.shape[0], v_.shape[1], 1, v_.shape[3]))))
    vx = (vx[:, :, 1:, 0] + vx[:, :, :-1, 0]) * 0.5
    vy = np.hstack((v_, np.zeros((v_.shape[0], 1, v_.shape[2], v_.shape[3]))))
    vy = (vy[:, 1:, :, 1] + vy[:, :-1, :, 1]) * 0.5
    vz = np.vstack((v_, np.zeros((1, v_.shape[1], v_.shape[2], v_.shape[3]))))
    vz = (vz[1:, :, :, 2] + vz[:-1, :, :, 2]) * 0.5
    return np.stack([vx, vy, vz], axis=-1)[:, ::-1]


def normalised_velocity(v_, scale):
    """cells/frame (x,y,z) -> normalised units in (z,y,x) order, y pointing down (:247-250)"""
    vx = v_[..., 0] / v_.shape[2] * scale
    vy = -v_[..., 1] / v_.shape[1] * scale
    vz = v_[..., 2] / v_.shape[0] * scale
    return np.stack([vz, vy, vx], axis=-1)


def save_particles(config, path, p, p_id, p_den):
    """id, position (domain units, x,y,z with y up), density [octave_n], Cd, radius (:288-320)"""
    px, py, pz = p[..., 2], 1 - p[..., 1], p[..., 0]
    p_ = np.stack([px * config.domain[2], py * config.domain[1], pz * config.domain[0]], axis=-1)
    pt = partio.create()
    pt.addAttribute('id', partio.INT, 1)
    pt.addAttribute('position', partio.VECTOR, 3)
    if p_den.shape[1] > 1:
        pt.addAttribute('density', partio.VECTOR, p_den.shape[1])
    else:
        pt.addAttribute('density', partio.FLOAT, 1)
    pt.addAttribute('Cd', partio.FLOAT, 3)
    pt.addAttribute('radius', partio.FLOAT, 1)
    pt.setArray('id', np.asarray(p_id, np.int32))
    pt.setArray('position', p_)
    pt.setArray('density', p_den)
    pt.setArray('Cd', np.repeat(p_den[:, :1], 3, axis=1))
    pt.setArray('radius', np.full([p_.shape[0], 1], config.radius, np.float32))
    partio.write(path, pt)


def render(d_smp, transmit):
    """front view of the sampled density (:322-328)"""
    tr = np.exp(-np.cumsum(d_smp[::-1], axis=0) * transmit)
    d_img = np.sum(d_smp * tr, axis=0)
    d_img /= d_img.max()
    return (d_img[::-1] * 255).astype(np.uint8)


def run(config):
    prepare_dirs_and_logger(config)
    config.rng = np.random.RandomState(config.seed)
    resampler = SimG2P(config)
    p = p_id = p_src = None
    n_prev, l = 0, 0
    for t in range(config.num_frames):
        with np.load(os.path.join(config.data_dir, config.dataset, config.d_path % (config.target_frame + t))) as data:
            d = data['x'][:, ::-1]                             # [D,H,W], [0-1]
        with np.load(os.path.join(config.data_dir, config.dataset, config.v_path % (config.target_frame + t))) as data:
            u = normalised_velocity(centred_velocity(data['x']), config.scale)
        if config.resampling:
            if t == 0:
                p, p_id = resampler.sample(d, disc=config.disc, threshold=0)     # sampling at the beginning wo opt.
            result = resampler.optimize(p, p_id, d, u)
            p, p_id, p_den, d_smp = result['p'], result['p_id'], result['p_den'], result['d_smp']
            l = result['l'][-1] if result['l'] else 0
        else:
            if t == 0:
                p, p_id = resampler.sample(d, disc=config.disc, threshold=0)
                p_src = p
            else:                                              # simply source particles of t=0
                p = np.concatenate([p, p_src], axis=0)
                p_id = np.arange(p.shape[0])
            p_den = np.ones([p.shape[0], 1])
            p, d_smp = resampler.naive_adv(p, u, p_den)
            l = 0
        print(t, 'num particles', p.shape[0], '(+%d)' % (p.shape[0] - n_prev), 'loss', l)
        n_prev = p.shape[0]
        save_particles(config, os.path.join(config.log_dir, '%03d.bgeo' % (config.target_frame + t)), p, p_id,
                       np.asarray(p_den, np.float32))
        save_png(render(d_smp, config.transmit), os.path.join(config.log_dir, '%03d.png' % (config.target_frame + t)))
    with open(os.path.join(config.log_dir, 'stat.txt'), 'w') as f:
        f.write('num particles %d\n' % p.shape[0])
        f.write('loss %.2f' % l)
    return p, p_id


def main(config):
    """scene constants of test_smokegun_resim.py:336-372"""
    config.dataset = 'smokegun'
    config.d_path = 'd_low/%03d.npz'
    config.v_path = 'v_low/%03d.npz'
    if not getattr(config, 'keep_resolution', False):
        config.num_frames = 120
        config.target_frame = 0
        config.scale = 1
        config.domain = [_ * config.scale for _ in [200, 300, 200]]
    config.resolution = [int(_) for _ in config.domain]
    config.disc = 1
    cell_size = 1                                              # == 2*radius*disc
    config.radius = cell_size / config.disc / 2
    config.nsize = 1
    config.support = 4
    config.rest_density = 1000
    config.threshold = 0.01
    config.lr = 0.0005
    config.iter = 20
    config.transmit = 0.01
    config.octave_n = 2
    config.octave_scale = 2 if config.octave_n > 1 else 1
    config.resampling = getattr(config, 'resampling', True)
    if config.resampling:
        config.tag = 'n%d_it%d_o%d' % (config.num_frames, config.iter, config.octave_n)
    else:
        config.tag = 'naive_n%d' % config.num_frames
    return run(config)


if __name__ == '__main__':
    config, unparsed = get_config()
    main(config)
