"""Flag surface of the stylisation engine -- a drop-in for the reference's ``config.py:15-113``.

Every flag name, type and default of the reference is kept so that existing driver scripts
(``config, _ = get_config()`` then attribute overrides, ``test_smokegun.py:111-197``) work
unchanged.  The table below is the single source of truth; ``get_config`` materialises it
into an ``argparse`` namespace exactly like the reference does.

Engine-only additions (all optional, reference-preserving defaults) are at the bottom.
"""
import argparse


def str2bool(v):
    """Same truthiness rule as the reference's ``util.str2bool`` (``util.py:401-402``)."""
    if isinstance(v, bool):
        return v
    return v.lower() in ('true', '1')


_NETWORKS = ['tensorflow_inception_graph.pb', 'vgg_19.ckpt', 'vgg_16.ckpt']

# (group, name, type, default, extra kwargs)
_FLAGS = [
    # config.py:15-21
    ('Path', 'data_dir', str, 'data', {}),
    ('Path', 'log_dir', str, 'log', {}),
    ('Path', 'model_dir', str, 'model', {}),
    ('Path', 'd_path', str, 'd/%03d.npz', {}),
    ('Path', 'v_path', str, 'v/%03d.npz', {}),
    ('Path', 'tag', str, 'test', {}),
    # config.py:24-28
    ('Data', 'dataset', str, 'smokegun', {}),
    ('Data', 'target_frame', int, 70, {}),
    ('Data', 'num_frames', int, 1, {}),
    ('Data', 'scale', float, 2.0, {}),
    # config.py:31-36
    ('Network', 'network', str, 'tensorflow_inception_graph.pb', {'choices': _NETWORKS}),
    ('Network', 'pool1', str2bool, False, {}),
    ('Network', 'batch_size', int, 1, {}),
    # config.py:39-41
    ('Grid', 'resolution', int, [384, 288], {'nargs': '+'}),
    ('Grid', 'adv_order', int, 1, {'choices': [1, 2]}),
    # config.py:44-56
    ('Particle', 'domain', int, [12.8, 12.8, 12.8], {'nargs': '+'}),
    ('Particle', 'radius', float, 0.025, {}),
    ('Particle', 'disc', int, 2, {}),
    ('Particle', 'nsize', int, 1, {}),
    ('Particle', 'rest_density', float, 1000, {}),
    ('Particle', 'w_pressure', float, 0, {}),
    ('Particle', 'w_density', float, 0, {}),
    ('Particle', 'window_sigma', float, 2, {}),
    ('Particle', 'interp', int, 1, {}),
    ('Particle', 'support', float, 4, {}),
    ('Particle', 'k', int, 3, {}),
    ('Particle', 'clip', str2bool, False, {}),
    # config.py:59-74
    ('Render', 'resize_scale', float, 1.0, {}),
    ('Render', 'transmit', float, 0.01, {}),
    ('Render', 'rotate', str2bool, False, {}),
    ('Render', 'phi0', int, -5, {}),
    ('Render', 'phi1', int, 5, {}),
    ('Render', 'phi_unit', int, 5, {}),
    ('Render', 'theta0', int, -10, {}),
    ('Render', 'theta1', int, 10, {}),
    ('Render', 'theta_unit', int, 10, {}),
    ('Render', 'v_batch', int, 1, {}),
    ('Render', 'n_views', int, 9, {}),
    ('Render', 'sample_type', str, 'poisson', {'choices': ['uniform', 'poisson', 'both']}),
    ('Render', 'render_liquid', str2bool, False, {}),
    # config.py:77-85
    ('Optimizer', 'target_field', str, 'p', {'choices': ['d', 'p', 'c']}),
    ('Optimizer', 'optimizer', str, 'adam', {}),
    ('Optimizer', 'iter', int, 20, {}),
    ('Optimizer', 'lr', float, 0.0007, {}),
    ('Optimizer', 'lr_scale', float, 1, {}),
    ('Optimizer', 'octave_n', int, 2, {}),
    ('Optimizer', 'octave_scale', float, 1.8, {}),
    ('Optimizer', 'frames_per_opt', int, 10, {}),
    # config.py:88-105
    ('Style', 'content_layer', str, 'mixed4d_3x3_bottleneck_pre_relu', {}),
    ('Style', 'content_channel', int, 139, {}),
    ('Style', 'w_content', float, 1, {}),
    ('Style', 'w_content_amp', float, 100, {}),
    ('Style', 'content_target', str, '', {}),
    ('Style', 'top_k', int, 5, {}),
    ('Style', 'style_layer', str, ['conv3_1'], {'nargs': '+'}),
    ('Style', 'w_style', float, 0, {}),
    ('Style', 'w_style_layer', float, [1], {'nargs': '+'}),
    ('Style', 'hist_layer', str, ['input'], {'nargs': '+'}),
    ('Style', 'w_hist', float, 0, {}),
    ('Style', 'w_hist_layer', float, [1], {'nargs': '+'}),
    ('Style', 'w_tv', float, 0, {}),
    ('Style', 'style_target', str, '', {}),
    ('Style', 'style_mask', str2bool, False, {}),
    ('Style', 'style_mask_on_ref', str2bool, False, {}),
    ('Style', 'style_tiling', int, 1, {}),
    ('Style', 'style_init', str, 'noise', {'choices': ['noise', 'style']}),
    # config.py:108-110
    ('Misc', 'seed', int, 123, {}),
    ('Misc', 'gpu_id', str, '0', {}),
    # ---- engine-only additions (not in the reference) -------------------------------
    # 'sequential' = reference semantics (one Adam step per view, iterates averaged,
    # styler_3p.py:329-352); 'allreduce' = mean view gradient, one Adam step, views shard
    # over ranks (BASELINE.json north_star).
    ('Engine', 'view_mode', str, 'sequential', {'choices': ['sequential', 'allreduce']}),
    # loss-network arithmetic: 'bf16' = tcgen05 tensor cores, 'fp32' = CUDA-core reference.
    ('Engine', 'conv_math', str, 'bf16x3', {'choices': ['bf16', 'bf16x3', 'fp32']}),
    # multi-net loss (BASELINE.json configs[4]: inception semantic + VGG style): when set, the content loss
    # (content_layer / content_channel) is evaluated on this second network, the style loss stays on `network`.
    ('Engine', 'content_network', str, '', {'choices': [''] + _NETWORKS}),
]

arg_lists = []
parser = argparse.ArgumentParser()
_groups = {}
for _g, _name, _type, _default, _kw in _FLAGS:
    if _g not in _groups:
        _groups[_g] = parser.add_argument_group(_g)
        arg_lists.append(_groups[_g])
    _groups[_g].add_argument('--' + _name, type=_type, default=_default, **_kw)


def get_config(argv=None):
    """``config.py:112-114``: returns ``(namespace, unparsed)``.  ``argv=None`` parses
    ``sys.argv`` like the reference; pass ``[]`` for pure defaults."""
    config, unparsed = parser.parse_known_args(argv)
    return config, unparsed
