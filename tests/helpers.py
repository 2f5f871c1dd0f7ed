"""Shared test helpers: reference-default configs and small seeded scenes."""
import copy

import numpy as np

from lnst.config import get_config


def make_cfg(**over):
    cfg, _ = get_config([])
    cfg = copy.deepcopy(cfg)
    # what the 3-D smoke driver sets (test_smokegun.py:111-160), scaled down by the caller
    cfg.network = 'vgg_19.ckpt'
    cfg.num_kernels = 2
    cfg.kernel_scale = 2
    cfg.w_content = 0
    cfg.w_style = 1
    cfg.style_layer = ['conv2_1', 'conv3_1']
    cfg.w_style_layer = [0.5, 0.5]
    cfg.octave_n = 1
    cfg.sample_type = 'uniform'
    cfg.rng = np.random.RandomState(cfg.seed)
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


def smoke_cfg(res=16, **over):
    """'d' mode, cubic cells of size 1 (radius .5) like the smokegun driver."""
    base = dict(target_field='d', resolution=[res, res, res], domain=[res, res, res], radius=0.5,
                nsize=1, support=4, k=3, transmit=0.05, lr=0.1, iter=3, rotate=False,
                rest_density=1000, window_sigma=0, frames_per_opt=1)
    base.update(over)
    return make_cfg(**base)


def liquid_cfg(res=16, **over):
    """'p' mode like the chocolate driver, scaled down."""
    cell = 0.1
    base = dict(target_field='p', resolution=[res, res, res], domain=[res * cell] * 3, radius=0.025,
                nsize=1, support=4, k=3, transmit=0.2, lr=0.002, iter=3, rotate=False,
                render_liquid=True, rest_density=1000, window_sigma=0, frames_per_opt=1)
    base.update(over)
    return make_cfg(**base)


def dam_cfg(**over):
    """'c' mode (2-D colour) like the dambreak2d driver (test_dambreak2d.py:128-190), scaled down."""
    res = [24, 32]
    cell = 0.1
    base = dict(target_field='c', resolution=res, domain=[r * cell for r in res], radius=0.025, nsize=2,
                support=4, rest_density=1000, lr=0.01, iter=3, octave_n=1, window_sigma=0, frames_per_opt=1,
                style_layer=['conv1_1'], w_style_layer=[1.0], w_style=1, w_tv=0, conv_math='fp32')
    base.update(over)
    return make_cfg(**base)
