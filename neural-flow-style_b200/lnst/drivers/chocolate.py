"""Liquid stylisation driver (position mode) -- mirrors reference ``test_chocolate.py``."""
import os

import numpy as np

from .. import partio
from ..config import get_config
from ..styler_3p import Styler
from ..util import prepare_dirs_and_logger
from . import frame_path, particle_range, save_loss_plot, save_renders


def load_particles(config):
    """positions [nmax,3] normalised (z,y,x), row = particle id, missing ids = -1 (test_chocolate.py:27-62)"""
    nmin, nmax = particle_range(config)
    print('# range:', nmin, nmax)
    p = []
    for i in range(config.num_frames):
        pt = partio.read(frame_path(config, i))
        ids = pt.array('id')[:, 0]
        p_ = np.ones([nmax, 3], dtype=np.float32) * -1
        p_[ids] = pt.array('position')[ids]
        px, py, pz = p_[..., 0], p_[..., 1], p_[..., 2]
        px /= config.domain[2]
        py /= config.domain[1]
        pz /= config.domain[0]
        p.append(np.stack([pz, py, px], axis=-1))
    return p


def save_particles(config, p_sty):
    """stylised positions back in domain units (x,y,z), padding rows dropped (test_chocolate.py:90-124)"""
    for i in range(config.num_frames):
        px, py, pz = p_sty[i][..., 2], p_sty[i][..., 1], p_sty[i][..., 0]
        p_sty_ = np.stack([px * config.domain[2], py * config.domain[1], pz * config.domain[0]], axis=-1)
        p_sty_ = p_sty_[p_sty_[:, 0] >= 0]
        pt = partio.create()
        pt.addAttribute('position', partio.VECTOR, 3)
        pt.addAttribute('radius', partio.FLOAT, 1)
        pt.setArray('position', p_sty_)
        pt.setArray('radius', np.full([p_sty_.shape[0], 1], config.radius, np.float32))
        partio.write(os.path.join(config.log_dir, '%03d.bgeo' % (config.target_frame + i)), pt)


def run(config, weights=None):
    prepare_dirs_and_logger(config)
    config.rng = np.random.RandomState(config.seed)

    styler = Styler(config, weights=weights)
    styler.load_img(config.resolution[1:])

    p = load_particles(config)
    print('resolution:', config.resolution)
    print('domain:', config.domain)
    print('radius:', config.radius)
    result = styler.run({'p': p})

    save_loss_plot(result['l'], config.log_dir)
    save_particles(config, result['p'])
    save_renders(config, result)
    return result


def main(config, weights=None):
    """scene constants of test_chocolate.py:140-186"""
    config.dataset = 'chocolate'
    config.d_path = 'partio/ParticleData_Fluid_%d.bgeo'
    config.radius = 0.025
    config.support = 4
    config.disc = 2
    config.rest_density = 1000
    base = [128, 128, 128]
    cell_size = 2 * config.radius * config.disc
    config.domain = [float(_ * cell_size) for _ in base]
    config.nsize = max(3 - config.disc, 1)
    if not getattr(config, 'keep_resolution', False):
        config.resolution = [200, 200, 200]                    # upscaling for rendering
    config.lr = 0.002
    config.iter = 20
    config.resize_scale = 1
    config.transmit = 0.2
    config.clip = False
    config.num_kernels = 1
    config.k = 3
    config.network = 'tensorflow_inception_graph.pb'
    config.style_layer = ['conv2d2', 'mixed3b', 'mixed4b']
    config.w_style_layer = [1, 1, 1]
    config.octave_n = 2
    config.octave_scale = 1.8
    config.render_liquid = True
    config.rotate = False
    config.frames_per_opt = 120
    config.batch_size = 1
    config.window_sigma = 9
    config.target_field = 'p'
    if config.w_content == 1:                                   # test_chocolate.py:237-245
        config.tag = 'test_%s_%s_%d' % (config.target_field, config.content_layer, config.content_channel)
    else:
        style = os.path.splitext(os.path.basename(config.style_target))[0]
        config.tag = 'test_%s_%s' % (config.target_field, style)
    config.tag += '_%d_intp%d' % (config.num_frames, config.interp)
    return run(config, weights)


if __name__ == '__main__':
    config, unparsed = get_config()
    main(config)
