"""Smoke stylisation driver (density mode) -- mirrors reference ``test_smokegun.py``."""
import os

import numpy as np

from .. import partio
from ..config import get_config
from ..styler_3p import Styler
from ..util import prepare_dirs_and_logger
from . import frame_path, particle_range, save_loss_plot, save_renders


def load_particles(config):
    """positions [nmax,3] normalised (z,y,x), padding rows -1; densities [nmax,num_kernels]
    (test_smokegun.py:29-65: rows are written in file order through the 'id' attribute)."""
    nmin, nmax = particle_range(config)
    print('# range:', nmin, nmax)
    p, r = [], []
    for i in range(config.num_frames):
        pt = partio.read(frame_path(config, i))
        ids = pt.array('id')[:, 0]
        p_ = np.ones([nmax, 3], dtype=np.float32) * -1
        r_ = np.zeros([nmax, config.num_kernels], dtype=np.float32)
        n = pt.numParticles()
        p_[:n] = pt.array('position')[ids]
        r_[:n] = pt.array('density')[ids]
        r.append(r_)
        px, py, pz = p_[..., 0], p_[..., 1], p_[..., 2]       # normalize particle position [0-1]
        px /= config.domain[2]
        py /= config.domain[1]
        pz /= config.domain[0]
        p.append(np.stack([pz, py, px], axis=-1))
    return p, r


def run(config, weights=None):
    prepare_dirs_and_logger(config)
    config.rng = np.random.RandomState(config.seed)

    styler = Styler(config, weights=weights)
    styler.load_img(config.resolution[1:])

    p, r = load_particles(config)
    print('resolution:', config.resolution)
    print('domain:', config.domain)
    print('radius:', config.radius)
    result = styler.run({'p': p, 'r': r})

    save_loss_plot(result['l'], config.log_dir)
    save_renders(config, result)
    for i, d_sty_ in enumerate(result['d']):                   # stylised fields, y flipped back (:94-97)
        np.savez_compressed(os.path.join(config.log_dir, '%03d.npz' % (config.target_frame + i)), x=d_sty_[:, ::-1])
    return result


def main(config, weights=None):
    """scene constants of test_smokegun.py:110-160"""
    config.dataset = 'smokegun'
    config.d_path = 'pt_low_o2/%03d.bgeo'
    config.num_kernels = 2
    config.kernel_scale = 2
    config.support = 4
    config.disc = 1
    cell_size = 1                                              # == 2*radius*disc
    config.radius = cell_size / config.disc / 2
    config.nsize = 1
    config.rest_density = 1000
    if not getattr(config, 'keep_resolution', False):
        config.resolution = [200, 300, 200]
        config.domain = [200, 300, 200]
    config.clip = False
    config.w_density = 0
    config.k = 3
    config.window_sigma = 3
    config.batch_size = 1
    config.frames_per_opt = 1
    config.target_field = 'd'
    config.lr = 0.1
    config.network = 'tensorflow_inception_graph.pb'
    config.style_layer = ['conv2d2', 'mixed3b', 'mixed4b']
    config.w_style_layer = [1, 1, 1]
    config.octave_n = 1
    config.octave_scale = 1.8
    config.transmit = 0.01
    config.iter = 20
    config.resize_scale = 300 / config.resolution[0]
    config.rotate = False
    config.interp = 1
    return run(config, weights)


if __name__ == '__main__':
    config, unparsed = get_config()
    main(config)
