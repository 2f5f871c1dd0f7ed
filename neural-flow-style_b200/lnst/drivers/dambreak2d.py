"""2-D colour stylisation driver -- mirrors reference ``test_dambreak2d.py``."""
import os

import numpy as np

from .. import partio
from ..config import get_config
from ..styler_2p import Styler
from ..util import prepare_dirs_and_logger
from . import frame_path, save_loss_plot, save_png


def load_particles(config):
    """positions [N,2] normalised (y,x) and SPH densities [N,1], row = particle id (test_dambreak2d.py:28-54)"""
    p, r = [], []
    for i in range(config.num_frames):
        pt = partio.read(frame_path(config, i))
        ids = pt.array('id')[:, 0]
        n = pt.numParticles()
        p_ = np.zeros([n, 2], dtype=np.float32)
        r_ = np.zeros([n, 1], dtype=np.float32)
        p_[ids] = pt.array('position')[ids, :2]
        r_[ids] = pt.array('density')[ids]
        r.append(r_)
        px, py = p_[..., 0] / config.domain[1], p_[..., 1] / config.domain[0]
        p.append(np.stack([py, px], axis=-1))
    return p, r


def save_particles(config, p, c_sty):
    """positions back in domain units (x,y) with the stylised colours as Cd (test_dambreak2d.py:98-121)"""
    for i in range(config.num_frames):
        px, py = p[i][..., 1] * config.domain[1], p[i][..., 0] * config.domain[0]
        pt = partio.create()
        pt.addAttribute('position', partio.VECTOR, 2)
        pt.addAttribute('Cd', partio.FLOAT, 3)
        pt.addAttribute('radius', partio.FLOAT, 1)
        pt.setArray('position', np.stack([px, py], axis=-1))
        pt.setArray('Cd', c_sty[i])
        pt.setArray('radius', np.full([px.shape[0], 1], config.radius, np.float32))
        partio.write(os.path.join(config.log_dir, '%03d.bgeo' % (config.target_frame + i)), pt)


def run(config, weights=None):
    prepare_dirs_and_logger(config)
    config.rng = np.random.RandomState(config.seed)

    styler = Styler(config, weights=weights)
    styler.load_img(config.resolution)

    p, r = load_particles(config)
    print('resolution:', config.resolution)
    print('domain:', config.domain)
    print('radius:', config.radius)
    print('num particles:', p[0].shape)
    result = styler.run({'p': p, 'r': r})

    save_loss_plot(result['l'], config.log_dir)
    for i, d_sty_ in enumerate(result['d']):                   # [0-255], uint8
        save_png(d_sty_, os.path.join(config.log_dir, '%03d.png' % (config.target_frame + i)))
    for o, d_intm_o in enumerate(result['d_intm']):
        for i, d_intm_ in enumerate(d_intm_o):
            save_png(d_intm_, os.path.join(config.log_dir, 'o%02d_%03d.png' % (o, config.target_frame + i)))
    save_particles(config, p, result['c'])
    return result


def main(config, weights=None):
    """scene constants of test_dambreak2d.py:123-190"""
    config.dataset = 'dambreak2d'
    config.d_path = 'partio/ParticleData_Fluid_%d.bgeo'
    config.radius = 0.025
    config.support = 4
    config.disc = 2
    config.rest_density = 1000
    base = [128, 256] if not getattr(config, 'keep_resolution', False) else list(config.resolution)
    cell_size = 2 * config.radius * config.disc
    config.domain = [float(_ * cell_size) for _ in base]
    config.nsize = max(3 - config.disc, 1)
    config.scale = 4                                           # upscaling for rendering
    config.nsize *= config.scale
    config.resolution = [base[0] * config.scale, base[1] * config.scale]
    config.frames_per_opt = 200
    config.window_sigma = 3
    config.target_field = 'c'
    config.lr = 0.01
    config.iter = 100
    config.octave_n = 3
    config.octave_scale = 1.7
    config.clip = False
    config.network = 'vgg_19.ckpt'
    config.w_style = 1
    config.w_content = 0
    config.style_init = 'noise'
    config.style_layer = ['conv2_1', 'conv3_1']
    config.w_style_layer = [0.5, 0.5]
    config.style_mask = True
    config.style_mask_on_ref = False
    config.style_tiling = 2
    config.w_tv = 0.01
    if config.w_content == 1:                                   # test_dambreak2d.py:192-200
        config.tag = 'test_%s_%s_%d' % (config.target_field, config.content_layer, config.content_channel)
    else:
        style = os.path.splitext(os.path.basename(config.style_target))[0]
        config.tag = 'test_%s_%s' % (config.target_field, style)
    config.tag += '_%d' % config.num_frames
    return run(config, weights)


if __name__ == '__main__':
    config, unparsed = get_config()
    main(config)
