#!/usr/bin/env python
"""Golden vectors from THE REFERENCE ITSELF, executed in this container.

    python tests/golden/make_reference_golden.py [case ...]      (needs /root/reference)

The reference's only backend, TensorFlow 1.15, cannot be installed here; ``oracle/tfshim`` stands
in for it (a restatement of the published semantics of the ~90 TF ops the path calls -- see that
module's header).  With it on ``sys.path`` this script imports the UNMODIFIED reference modules
from ``/root/reference`` (``styler_3p.Styler``, ``styler_2p.Styler``, ``transform``, ``vgg``,
``styler_base``, ``config``) and calls ``Styler(config).run(params)`` exactly as the reference's
drivers do (``test_smokegun.py:60-75``, ``test_dambreak2d.py:75-90``).  Graph construction, the
session loop, the optimiser calls, the view sampling and the host post-processing that produce
these numbers are therefore the reference's own code, line by line.

Outputs go to ``tests/golden/ref_<case>.npz`` (inputs are re-generated from seeds by the tests):
``tests/test_reference_golden.py`` holds the CPU oracle (``-m "not gpu"``) and the CUDA path through
the C-ABI (``-m gpu``) to them.  Function-level vectors (``ref_ops.npz``) pin the individual
``transform.py`` operators (p2g, p2g_wavg, rotate, advect, rot_mat, W) forward and gradient.

Only two NumPy compat patches (``np.int = int``, removed in NumPy 1.24, ``styler_3p.py:245``; a float
``num`` for ``np.linspace``, ``transform.py:755``, truncated as NumPy < 1.18 did) and the
stand-in modules for matplotlib / skimage / imageio (plotting + same-size target resize) are added
around the reference; the loss-network checkpoint file is replaced by the seeded synthetic weights
(``slim.assign_from_checkpoint_fn`` -> ``tfshim.register_checkpoint``).
"""
import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('LNST_REFERENCE', '/root/reference')
SHIM = os.path.join(ROOT, 'oracle', 'tfshim')


def _setup_paths():
    for p in (os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'neural-flow-style_b200'), ROOT):
        if p not in sys.path:
            sys.path.append(p)
    real_tf = bool(os.environ.get('LNST_REAL_TF'))  # a real TensorFlow 1.15 installation: leave the stand-in out
    for p in ((REF,) if real_tf else (REF, SHIM)):  # the reference's modules and the TF stand-in win
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    if not hasattr(np, 'int'):
        np.int = int                           # styler_3p.py:245 / styler_2p.py:177 (NumPy < 1.24 spelling)
    if not getattr(np.linspace, '_lnst_compat', False):
        _linspace = np.linspace

        def linspace(start, stop, num=50, *a, **k):  # transform.py:755,761 pass a float count; NumPy < 1.18
            return _linspace(start, stop, int(num), *a, **k)   # truncated it with int() (DeprecationWarning)
        linspace._lnst_compat = True
        np.linspace = linspace


def reference_config(**over):
    """The reference's own ``config.get_config()`` + the attributes its drivers add
    (``test_smokegun.py:111-160``), with the same scaled-down overrides as ``tests/helpers.py``."""
    argv, sys.argv = sys.argv, sys.argv[:1]
    try:
        import config as ref_config
        cfg, _ = ref_config.get_config()
    finally:
        sys.argv = argv
    cfg = copy.deepcopy(cfg)
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


def _cfg_from_helper(helper_cfg):
    """Copy every attribute of a tests/helpers config onto the reference's namespace."""
    over = {k: getattr(helper_cfg, k) for k in vars(helper_cfg)}
    return reference_config(**over)


def register_weights(cfg, weights):
    import tensorflow as tf
    path = os.path.join(cfg.data_dir, cfg.model_dir, cfg.network)
    model = os.path.basename(path).split('.')[0]
    ck = {}
    for name, (w, b) in weights.items():
        block = name.split('_')[0]             # conv3_1 -> vgg_19/conv3/conv3_1/{weights,biases}
        ck['%s/%s/%s/weights' % (model, block, name)] = np.asarray(w)
        ck['%s/%s/%s/biases' % (model, block, name)] = np.asarray(b)
    tf.register_checkpoint(path, ck)


# name: (kind, helper, overrides, n particles)
CASES = {
    'density_sequential': ('3d', 'smoke', dict(res=12, iter=3, rotate=True, n_views=9,
                                               style_layer=['conv1_2', 'conv2_1'], w_style_layer=[0.5, 0.5]), 900),
    'density_noview': ('3d', 'smoke', dict(res=14, iter=4, rotate=False,
                                           style_layer=['conv2_1', 'conv3_1'], w_style_layer=[0.5, 0.5]), 1200),
    'density_resize_tv_content': ('3d', 'smoke', dict(res=12, iter=3, rotate=False, resize_scale=1.5, w_tv=1e-3,
                                                      w_content=0.3, content_layer='conv2_1', content_channel=5,
                                                      style_layer=['conv1_2'], w_style_layer=[1.0]), 800),
    # octaves need the target at two sizes (skimage resampling, deviation D4): use the target-free content loss
    'density_octaves_poisson': ('3d', 'smoke', dict(res=16, iter=2, rotate=True, n_views=4, sample_type='poisson',
                                                    octave_n=2, octave_scale=1.8, lr_scale=1.5, w_style=0,
                                                    w_content=1.0, content_layer='conv2_1', content_channel=5), 1000),
    'density_sequence': ('3d', 'smoke', dict(res=12, iter=3, rotate=False, num_frames=4, window_sigma=1.0,
                                             frames_per_opt=2, style_layer=['conv1_2'], w_style_layer=[1.0]), 600),
    # key frames 0,2,4 optimised, frames 1,3 interpolated (styler_3p.py:392-397); 'both' view sampling re-drawn per iteration
    'density_interp_both': ('3d', 'smoke', dict(res=12, iter=2, rotate=True, n_views=3, sample_type='both', num_frames=5,
                                                interp=2, window_sigma=0.8, style_layer=['conv1_2'], w_style_layer=[1.0]), 500),
    # density regulariser (styler_base.py:217-223) and a content TARGET image (:135-141, :233-247)
    'density_reg_content_image': ('3d', 'smoke', dict(res=12, iter=3, rotate=False, w_density=1e-6, w_content=0.5,
                                                      w_content_amp=2.0, content_layer='conv1_2', content_image=True, top_k=0,
                                                      style_layer=['conv2_1'], w_style_layer=[1.0]), 700),
    'position_clip_vgg16': ('3p', 'liquid', dict(res=12, iter=3, clip=True, network='vgg_16.ckpt',
                                                 style_layer=['conv1_2', 'conv2_2'], w_style_layer=[0.7, 0.3]), 600),
    'position_liquid': ('3p', 'liquid', dict(res=12, iter=3, w_pressure=0.5, style_layer=['conv1_2'],
                                             w_style_layer=[1.0]), 700),
    'position_smoke_views': ('3p', 'liquid', dict(res=12, iter=2, rotate=True, n_views=3, render_liquid=False,
                                                  transmit=0.05, style_layer=['conv1_2'], w_style_layer=[1.0]), 500),
    # inception5h path (styler_base.py:17-31,53-57,91-94): the reference parses the GraphDef file (here: the seeded
    # synthetic graph with the inception5h topology, written by lnst.graphdef.serialize and read back by the
    # protobuf library inside oracle/tfshim), imports it and reads layers by tensor name -- semantic transfer on a
    # pre-ReLU bottleneck channel (run.bat:15-20) plus the style layers of test_smokegun.py:141, 3 views
    'density_inception': ('3d', 'smoke', dict(res=20, iter=3, network='tensorflow_inception_graph.pb', rotate=True,
                                              n_views=3, w_content=0.7, content_layer='mixed3b_3x3_bottleneck_pre_relu',
                                              content_channel=5, style_layer=['conv2d2', 'mixed3a', 'mixed3b'],
                                              w_style_layer=[1, 1, 1]), 900),
    'density_inception_pool1': ('3d', 'smoke', dict(res=14, iter=2, network='tensorflow_inception_graph.pb', pool1=True,
                                                    w_style=0, w_content=1.0, content_layer='mixed3a_pool_reduce_pre_relu',
                                                    content_channel=0), 600),
    # style mask in 3-D (styler_base.py:165-169): the mask is the normalised render itself, so the gradient also
    # flows through the mask and through the masked area in the Gram denominator
    'density_style_mask': ('3d', 'smoke', dict(res=12, iter=3, rotate=True, n_views=3, style_mask=True,
                                               style_layer=['conv1_2', 'conv2_1'], w_style_layer=[0.5, 0.5]), 800),
    # v_batch > 1 (config.py:69, styler_3p.py:329-352): a group of views is rendered and normalised by ONE maximum,
    # the Gram loss reads only the group's first image (styler_base.py:98), content / TV average over the group
    'density_vbatch': ('3d', 'smoke', dict(res=12, iter=2, rotate=True, n_views=6, v_batch=3, w_tv=1e-3, w_content=0.4,
                                           content_layer='conv2_1', content_channel=5, style_layer=['conv1_2'],
                                           w_style_layer=[1.0]), 800),
    # batch_size > 1 in 3-D without rotate (styler_3p.py:42,304-363,409-431): two frames per sess.run -- ONE maximum
    # normalises both renders (:158, also in the final renders), Gram terms summed over the batch, content / TV /
    # pressure means over it, one Adam op over both frames' variables
    'density_batch': ('3d', 'smoke', dict(res=12, iter=3, rotate=False, num_frames=4, batch_size=2, frames_per_opt=4,
                                          window_sigma=1.0, w_tv=1e-3, w_content=0.4, content_layer='conv2_1',
                                          content_channel=5, style_layer=['conv1_2'], w_style_layer=[1.0]), 600),
    'position_batch_pressure': ('3p', 'liquid', dict(res=12, iter=3, num_frames=2, batch_size=2, frames_per_opt=2,
                                                     w_pressure=0.5, style_layer=['conv1_2'], w_style_layer=[1.0]), 600),
    'density_style_mask_on_ref': ('3d', 'smoke', dict(res=12, iter=3, rotate=True, n_views=3, style_mask=True,
                                                      style_mask_on_ref=True, style_layer=['conv1_2', 'conv2_1'],
                                                      w_style_layer=[0.5, 0.5]), 800),
    # content TARGET IMAGE on the inception class logits with top_k (styler_base.py:135-141, 233-247): the head of the
    # graph (AvgPool -> Reshape -> MatMul -> BiasAdd = softmax2_pre_activation), the k strongest logits of the target kept
    'density_inception_logits': ('3d', 'smoke', dict(res=20, iter=3, network='tensorflow_inception_graph.pb', rotate=False,
                                                     w_style=0, w_content=1.0, w_content_amp=3.0, content_image=True, top_k=3,
                                                     content_layer='softmax2_pre_activation'), 900),
    'colour_2d': ('2c', 'dam', dict(iter=4, w_tv=0.01, style_layer=['conv1_1', 'conv2_1'], w_style_layer=[0.5, 0.5]), 0),
    'colour_2d_mask': ('2c', 'dam', dict(iter=3, style_mask=True, style_layer=['conv1_1', 'conv2_1'],
                                         w_style_layer=[0.5, 0.5]), 0),
    'colour_2d_frames': ('2c', 'dam', dict(iter=2, num_frames=3, window_sigma=1.0), 0),
    # style_mask_on_ref (styler_base.py:171-173): the style feature is masked and area-normalised like the render's
    'colour_2d_mask_on_ref': ('2c', 'dam', dict(iter=3, style_mask=True, style_mask_on_ref=True,
                                                style_layer=['conv1_1', 'conv2_1'], w_style_layer=[0.5, 0.5]), 0),
    # batch_size > 1 (styler_2p.py:42,236-262): two frames per sess.run -- one joint loss (Gram terms summed over the
    # batch, TV / content averaged), one Adam op over both frames' variables, slots shared by the frames of a group
    'colour_2d_batch': ('2c', 'dam', dict(iter=3, num_frames=4, batch_size=2, frames_per_opt=4, window_sigma=1.0, w_tv=0.01,
                                          w_content=0.3, content_layer='conv1_1', content_channel=3), 0),
}


def inception_nodes(head=False):
    """The synthetic inception5h graph of the 'density_inception*' cases (shared with the tests); ``head``: with the
    classifier head (a 2x2 avgpool0: the 20x20 test renders leave a 3x3 mixed3b map)."""
    from lnst import synth
    return synth.inception5h_nodes(width_div=8, upto='mixed3b', head_pool=2 if head else 0)


def case_inputs(name):
    """(helper config, params dict, style targets per octave) -- shared with the tests."""
    from helpers import smoke_cfg, liquid_cfg
    from lnst import synth
    kind, helper, kw, n = CASES[name]
    if kind == '3d':
        cfg = smoke_cfg(**kw)
        p, r = synth.smoke_particles(n, cfg.num_kernels, pad=4, num_frames=cfg.num_frames)
        params = {'p': p, 'r': r}
    elif kind == '3p':
        cfg = liquid_cfg(**kw)
        params = {'p': synth.liquid_particles(n, num_frames=cfg.num_frames)}
    else:
        from helpers import dam_cfg
        cfg = dam_cfg(**kw)
        p, r = synth.dam_particles_2d(cfg.domain, num_frames=cfg.num_frames)
        params = {'p': p, 'r': r}
    return cfg, params


def run_reference(name):
    _setup_paths()
    import tensorflow as tf
    real_tf = bool(os.environ.get('LNST_REAL_TF'))
    assert real_tf or 'shim' in tf.__version__
    from lnst import synth
    kind = CASES[name][0]
    hcfg, params = case_inputs(name)
    cfg = _cfg_from_helper(hcfg)
    cfg.rng = np.random.RandomState(cfg.seed)
    if 'inception' in cfg.network:             # the GraphDef file the reference opens (styler_base.py:18-23)
        import tempfile
        from lnst import graphdef
        cfg.data_dir = tempfile.mkdtemp(prefix='lnst_ref_')
        os.makedirs(os.path.join(cfg.data_dir, cfg.model_dir))
        with open(os.path.join(cfg.data_dir, cfg.model_dir, cfg.network), 'wb') as f:
            f.write(graphdef.serialize(inception_nodes(head='logits' in name)))
    elif not real_tf:                          # real TF: <data_dir>/<model_dir>/vgg_19.ckpt must hold the seeded weights
        register_weights(cfg, synth.vgg_weights('vgg_16' if '16' in cfg.network else 'vgg_19'))
    if kind == '2c':
        import styler_2p as mod
    else:
        import styler_3p as mod
    styler = mod.Styler(cfg)
    # the drivers call load_img (file -> array); here the target is the seeded synthetic image,
    # generated at the octave size so util.resize is the identity (deviation D4)
    res = cfg.resolution
    hw = [int(int(s) * cfg.resize_scale) for s in res[-2:]] if not np.isclose(cfg.resize_scale, 1) else list(res[-2:])
    styler.content_img = None
    styler.style_img = None
    if cfg.w_style:
        assert cfg.octave_n == 1
        styler.style_img = synth.style_image(hw[0], hw[1])
    if getattr(cfg, 'content_image', False):   # load_img would read config.content_target; same-size seeded image instead
        styler.content_img = synth.style_image(hw[0], hw[1], seed=11)
    out = styler.run(params)
    res_d = {'l': np.asarray(out['l'], np.float64)}
    if kind == '2c':                           # styler_2p.py:289-314: 'd' is the uint8 colour image, 'c' the masked colours
        res_d['d'] = np.asarray(out['d'])
    else:
        res_d['d'] = np.asarray(out['d'], np.float32)
        res_d['r'] = np.asarray(out['r'])
    if out.get('v') is not None:
        res_d['v'] = np.asarray(out['v'], np.float32)
    if out.get('c') is not None:
        res_d['c'] = np.asarray(out['c'], np.float32)
    if len(out.get('d_intm', [])):
        for i, a in enumerate(out['d_intm']):
            res_d['d_intm%d' % i] = np.asarray(a)
    return res_d


def ops_inputs():
    """Seeded inputs of the operator-level vectors (shared with tests/test_reference_golden.py)."""
    rng = np.random.RandomState(2024)
    d = {}
    p3 = rng.uniform(-0.03, 1.03, size=(260, 3)).astype(np.float32)
    p3[:3] = -1.0                                   # the drivers' padding rows
    d['p3'] = p3
    d['g3'] = rng.randn(7, 9, 8).astype(np.float32)        # cotangent for the 3-D splats, res [7,9,8]
    blob = rng.uniform(0.3, 0.7, (700, 3))
    lone = np.array([[0.08, 0.1, 0.12], [0.9, 0.15, 0.5], [-1, -1, -1]])
    d['pw'] = np.concatenate([blob, lone]).astype(np.float32)
    d['xw'] = rng.uniform(-0.2, 1.2, (d['pw'].shape[0], 1)).astype(np.float32)
    d['gw'] = rng.randn(8, 8, 8).astype(np.float32)
    d['p2'] = rng.uniform(0.03, 0.97, (140, 2)).astype(np.float32)
    d['pc'] = rng.uniform(0, 1, (140, 3)).astype(np.float32)
    d['pd'] = (1000 * (1 + 0.02 * rng.randn(140, 1))).astype(np.float32)
    d['g2'] = rng.randn(10, 14, 3).astype(np.float32)
    d['vol'] = rng.rand(9, 8, 10).astype(np.float32)
    d['adv2_d'] = rng.rand(6, 7, 2).astype(np.float32)
    d['adv2_v'] = rng.uniform(-0.5, 0.5, (6, 7, 2)).astype(np.float32)
    d['adv3_d'] = rng.rand(5, 6, 4, 2).astype(np.float32)
    d['adv3_v'] = rng.uniform(-0.5, 0.5, (5, 6, 4, 3)).astype(np.float32)
    d['q'] = np.linspace(0, 1.2, 49).astype(np.float32)
    return d


def run_ops():
    """Operator-level vectors from the reference's ``transform.py`` (forward and ``tf.gradients``)."""
    _setup_paths()
    import tensorflow as tf
    import torch
    import transform as RT
    I = ops_inputs()
    sess = tf.Session()
    out = {}

    def leaf(a):
        return torch.tensor(a, requires_grad=True)

    # W('cubic'), 2-D and 3-D normalisation (transform.py:1233-1245)
    q = tf.placeholder(tf.float32, [None])
    out['W3'] = sess.run(RT.W('cubic')(q, 0.37, is_3d=True), {q: I['q']})
    out['W2'] = sess.run(RT.W('cubic')(q, 0.37, is_3d=False), {q: I['q']})

    # p2g 3-D, clip False / True, gradient w.r.t. the positions (transform.py:1310-1453)
    res, dom = [7, 9, 8], [0.7, 0.9, 0.8]
    for clip in (False, True):
        ph = tf.placeholder(tf.float32, [1, None, 3])
        y = RT.p2g(ph, dom, tf.constant(res), 0.025, 1000.0, 1, is_2d=False, clip=clip, support=4)
        loss = tf.reduce_sum(y[0, ..., 0] * I['g3'])
        pl = leaf(I['p3'][None])
        yv, gv = sess.run([y, tf.gradients(loss, [ph])[0]], {ph: pl})
        out['p2g3_clip%d' % clip], out['p2g3_clip%d_grad' % clip] = yv, gv

    # p2g 2-D colour (pc, pd), nsize 2 -- gradient w.r.t. the colours (styler_2p.py:75-76)
    res2, dom2 = [10, 14], [1.0, 1.4]
    ph, pc, pd = tf.placeholder(tf.float32, [1, None, 2]), tf.placeholder(tf.float32, [1, None, 3]), \
        tf.placeholder(tf.float32, [1, None, 1])
    y = RT.p2g(ph, dom2, tf.constant(res2), 0.025, 1000.0, 2, pc=pc, pd=pd, is_2d=True, clip=False, support=4)
    loss = tf.reduce_sum(y[0] * I['g2'])
    cl = leaf(I['pc'][None])
    out['p2g2_colour'], out['p2g2_colour_grad'] = sess.run([y, tf.gradients(loss, [pc])[0]],
                                                           {ph: I['p2'][None], pc: cl, pd: I['pd'][None]})
    out['p2g2_gray'] = sess.run(RT.p2g(ph, dom2, tf.constant(res2), 0.025, 1000.0, 2, is_2d=True, clip=False),
                                {ph: I['p2'][None]})

    # p2g_wavg 3-D incl. the NaN-gradient rule (transform.py:1577-1704)
    resw, domw = [8, 8, 8], [8, 8, 8]
    for k, support in enumerate((4.0, 2.0)):
        ph, xh = tf.placeholder(tf.float32, [1, None, 3]), tf.placeholder(tf.float32, [1, None, 1])
        y = RT.p2g_wavg(ph, xh, domw, tf.constant(resw), 0.5, 1, kernel='cubic', support=support, clip=False, is_2d=False)
        loss = tf.reduce_sum(y[0, ..., 0] * I['gw'])
        xl = leaf(I['xw'][None])
        out['wavg_s%d' % k], out['wavg_s%d_grad' % k] = sess.run([y, tf.gradients(loss, [xh])[0]],
                                                                 {ph: I['pw'][None], xh: xl})

    # rotate (transform.py:611-628) with the uniform view matrices + one large rotation
    mats, _ = RT.rot_mat(-5, 5, 5, -10, 10, 10, sample_type='uniform')
    mats = mats[:3] + [np.matmul(RT.rot_y_3d(40.0), RT.rot_z_3d(25.0))]
    dph = tf.placeholder(tf.float32, [1, None, None, None, 1])
    d_rot, rm = RT.rotate(dph)
    out['rotate'] = sess.run(d_rot, {dph: I['vol'][None, ..., None], rm: mats})
    out['rotate_mats'] = np.asarray(mats)

    # advect order 1 (transform.py:557-609)
    d2, v2 = tf.placeholder(tf.float32, [1, None, None, 2]), tf.placeholder(tf.float32, [1, None, None, 2])
    out['advect2'] = sess.run(RT.advect(d2, v2, order=1, is_3d=False), {d2: I['adv2_d'][None], v2: I['adv2_v'][None]})
    d3, v3 = tf.placeholder(tf.float32, [1, None, None, None, 2]), tf.placeholder(tf.float32, [1, None, None, None, 3])
    out['advect3'] = sess.run(RT.advect(d3, v3, order=1, is_3d=True), {d3: I['adv3_d'][None], v3: I['adv3_v'][None]})

    # view sampling: every sample_type, RNG order included (transform.py:640-768, 14-150)
    for st in ('uniform', 'poisson', 'both'):
        rng = np.random.RandomState(123)
        for rep in range(2):                         # the loop re-draws every iteration (styler_3p.py:344-349)
            m, _ = RT.rot_mat(-5, 5, 5, -10, 10, 10, sample_type=st, rng=rng, nv=9)
            out['views_%s_%d' % (st, rep)] = np.asarray(m)
    return out


def main(argv):
    if argv == ['ops'] or not argv:
        out = run_ops()
        np.savez_compressed(os.path.join(HERE, 'ref_ops.npz'), **out)
        print('== reference operators:', sorted(out))
        if argv:
            return
    names = argv or list(CASES)
    for name in names:
        print('== reference run:', name, flush=True)
        out = run_reference(name)
        np.savez_compressed(os.path.join(HERE, 'ref_%s.npz' % name), **out)
        print('   l =', out['l'].tolist())


if __name__ == '__main__':
    main(sys.argv[1:])
