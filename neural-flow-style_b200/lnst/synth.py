"""Seeded synthetic inputs for tests and ``bench.py``.

The reference's datasets come from external simulators (mantaflow / SPlisHSPlasH) and its
loss-network weights from downloads; none of that exists offline, so every measurement uses
the generators below.  All use ``numpy.random.RandomState`` / a CPU ``torch.Generator`` so the
same arrays are produced on every machine.
"""
from collections import OrderedDict

import numpy as np
import torch

_VGG_CFG = {
    'vgg_19': [(2, 64), (2, 128), (4, 256), (4, 512), (4, 512)],
    'vgg_16': [(2, 64), (2, 128), (3, 256), (3, 512), (3, 512)],
}


def smoke_particles(n, num_kernels=2, seed=123, centre=(0.5, 0.45, 0.5), radii=(0.30, 0.35, 0.30),
                    num_frames=1, pad=0):
    """Particles uniform inside an ellipsoid, normalised (z,y,x); r[:,0]~U(.2,1) and the
    finer kernels ~U(-.1,.1).  ``pad`` extra rows are the drivers' p=-1 padding
    (``test_smokegun.py:48``).  Later frames move by a small rigid swirl so ids persist."""
    rng = np.random.RandomState(seed)
    pts = np.zeros((0, 3))
    while pts.shape[0] < n:
        c = rng.uniform(-1, 1, size=(2 * n + 16, 3))
        pts = np.concatenate([pts, c[(c ** 2).sum(-1) < 1.0]], 0)
    pts = pts[:n] * np.array(radii) + np.array(centre)
    r0 = np.concatenate([rng.uniform(0.2, 1.0, size=(n, 1)),
                         rng.uniform(-0.1, 0.1, size=(n, max(num_kernels - 1, 0)))], -1)
    p_list, r_list = [], []
    for t in range(num_frames):
        ang = 0.02 * t
        c, s = np.cos(ang), np.sin(ang)
        q = pts - np.array(centre)
        q = np.stack([q[:, 0] * c - q[:, 2] * s, q[:, 1] + 0.002 * t, q[:, 0] * s + q[:, 2] * c], -1)
        pt = (q + np.array(centre)).astype(np.float32)
        rt = r0.astype(np.float32)
        if pad:
            pt = np.concatenate([pt, -np.ones((pad, 3), np.float32)], 0)
            rt = np.concatenate([rt, np.zeros((pad, num_kernels), np.float32)], 0)
        p_list.append(pt)
        r_list.append(rt)
    return p_list, r_list


def liquid_particles(n, seed=123, num_frames=1):
    """A liquid blob for the position mode ('p'): jittered lattice in a box, moved by an
    analytic divergence-free swirl between frames (persistent ids)."""
    rng = np.random.RandomState(seed)
    m = int(np.ceil(n ** (1 / 3.0)))
    g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing='ij'), -1).reshape(-1, 3)[:n]
    base = 0.3 + 0.4 * (g + 0.5) / m + rng.uniform(-0.2, 0.2, size=(n, 3)) * 0.4 / m
    out = []
    for t in range(num_frames):
        ang = 0.03 * t
        c, s = np.cos(ang), np.sin(ang)
        q = base - 0.5
        q = np.stack([q[:, 0], q[:, 1] * c - q[:, 2] * s, q[:, 1] * s + q[:, 2] * c], -1) + 0.5
        out.append(q.astype(np.float32))
    return out


def dam_particles_2d(domain, spacing=0.05, seed=123, num_frames=1, rest_density=1000.0):
    """2-D dam block (``test_dambreak2d.py`` style): lattice over x<.35 Dx, y<.6 Dy with
    jitter; r = SPH density ~ rho0 (1 + .02 N)."""
    rng = np.random.RandomState(seed)
    dy, dx = float(domain[0]), float(domain[1])
    ys = np.arange(spacing * 0.5, 0.6 * dy, spacing)
    xs = np.arange(spacing * 0.5, 0.35 * dx, spacing)
    g = np.stack(np.meshgrid(ys, xs, indexing='ij'), -1).reshape(-1, 2)
    n = g.shape[0]
    p_list, r_list = [], []
    for t in range(num_frames):
        pt = g + rng.uniform(-0.1 * spacing, 0.1 * spacing, size=g.shape) + np.array([0.0, 0.01 * t])
        pt = (pt / np.array([dy, dx])).astype(np.float32)
        p_list.append(pt)
        r_list.append((rest_density * (1 + 0.02 * rng.randn(n, 1))).astype(np.float32))
    return p_list, r_list


def style_image(h, w, seed=7):
    """Low-pass seeded noise, [h,w,3] float32 in 0..255 -- generated directly at the octave
    size so no image resampling is involved in parity runs."""
    from scipy.ndimage import gaussian_filter
    rng = np.random.RandomState(seed)
    img = rng.uniform(0, 1, size=(h, w, 3))
    img = gaussian_filter(img, sigma=(2.0, 2.0, 0))
    img = (img - img.min()) / (img.max() - img.min() + 1e-12)
    return (img * 255).astype(np.float32)


def vgg_weights(model='vgg_19', seed=19):
    """Seeded He-normal weights in slim layout: {'conv1_1': (w[3,3,Cin,Cout], b[Cout]), ...}
    (float32 CPU tensors).  Same recipe as ``oracle.vgg.synthetic_weights`` -- kept separate so
    the product never imports the oracle; ``tests/test_synth.py`` checks they agree."""
    g = torch.Generator().manual_seed(seed)
    w = OrderedDict()
    cin = 3
    for b, (rep, cout) in enumerate(_VGG_CFG[model], start=1):
        for i in range(1, rep + 1):
            std = float(np.sqrt(2.0 / (9 * cin)))
            wt = torch.randn(3, 3, cin, cout, generator=g, dtype=torch.float32) * std
            bs = torch.randn(cout, generator=g, dtype=torch.float32)
            w['conv%d_%d' % (b, i)] = (wt, bs)
            cin = cout
    return w


# ---- synthetic inception5h GraphDef (the real tensorflow_inception_graph.pb is not available offline) -------
_INCEPTION_MODULES = [       # name, 1x1, 3x3 bottleneck, 3x3, 5x5 bottleneck, 5x5, pool_reduce  (GoogLeNet / inception5h)
    ('mixed3a', 64, 96, 128, 16, 32, 32), ('mixed3b', 128, 128, 192, 32, 96, 64), 'maxpool4',
    ('mixed4a', 192, 96, 204, 16, 48, 64), ('mixed4b', 160, 112, 224, 24, 64, 64),
    ('mixed4c', 128, 128, 256, 24, 64, 64), ('mixed4d', 112, 144, 288, 32, 64, 64),
    ('mixed4e', 256, 160, 320, 32, 128, 128), 'maxpool10',
    ('mixed5a', 256, 160, 320, 48, 128, 128), ('mixed5b', 384, 192, 384, 48, 128, 128),
]


def inception5h_nodes(seed=5, width_div=1, upto=None, head_pool=0, n_classes=1008):
    """The inception5h topology with the file's node names (``conv2d0_pre_relu/conv`` -> ``conv2d0_pre_relu`` ->
    ``conv2d0``, ``mixed3a_3x3_bottleneck_pre_relu``, ``mixed3a`` ...; run.bat:15-20 and test_smokegun.py:141 name
    such tensors) as a list of ``lnst.graphdef.Node`` with seeded He-normal weights.  ``width_div`` divides every
    channel count (tests); ``upto`` stops after that module.  LRN attributes are those of the Caffe GoogLeNet
    conversion (radius 2, bias 1, alpha 2e-5, beta .75) -- the real file's own attributes are used when it is
    loaded.  Serialise with ``lnst.graphdef.serialize`` to get a ``.pb``.

    ``head_pool`` > 0 appends the classifier head after the last module built: ``avgpool0`` (AvgPool head_pool x head_pool,
    stride 1, VALID; 7 in the real file, for 224 x 224 inputs) -> ``avgpool0/reshape`` [-1, C] ->
    ``softmax2_pre_activation/matmul`` -> ``softmax2_pre_activation`` (BiasAdd): the class logits the reference's
    content-target mode with ``top_k`` reads (styler_base.py:240-246)."""
    from .graphdef import Node
    rng = np.random.RandomState(seed)
    nodes = [Node('input', 'Placeholder', [], {})]

    def ch(c):
        return max(int(c) // width_div, 2)

    def conv(name, src, k, cin, cout, stride=1):
        w = (rng.randn(k, k, cin, cout) * np.sqrt(2.0 / (k * k * cin))).astype(np.float32)
        b = (0.05 * rng.randn(cout)).astype(np.float32)
        nodes.append(Node(name + '_w', 'Const', [], {'value': w}))
        nodes.append(Node(name + '_b', 'Const', [], {'value': b}))
        nodes.append(Node(name + '_pre_relu/conv', 'Conv2D', [src, name + '_w'],
                          {'strides': [1, stride, stride, 1], 'padding': b'SAME'}))
        nodes.append(Node(name + '_pre_relu', 'BiasAdd', [name + '_pre_relu/conv', name + '_b'], {}))
        nodes.append(Node(name, 'Relu', [name + '_pre_relu'], {}))
        return name, cout

    def pool(name, src, stride):
        nodes.append(Node(name, 'MaxPool', [src], {'ksize': [1, 3, 3, 1], 'strides': [1, stride, stride, 1],
                                                    'padding': b'SAME'}))
        return name

    def lrn(name, src):
        nodes.append(Node(name, 'LRN', [src], {'depth_radius': 2, 'bias': 1.0, 'alpha': 2e-5, 'beta': 0.75}))
        return name

    cur, c = conv('conv2d0', 'input', 7, 3, ch(64), stride=2)
    cur = lrn('localresponsenorm0', pool('maxpool0', cur, 2))
    cur, c = conv('conv2d1', cur, 1, c, ch(64))
    cur, c = conv('conv2d2', cur, 3, c, ch(192))
    cur = pool('maxpool1', lrn('localresponsenorm1', cur), 2)
    for m in _INCEPTION_MODULES:
        if isinstance(m, str):
            cur = pool(m, cur, 2)
            continue
        name, c1, c3b, c3, c5b, c5, cp = m
        a, ca = conv(name + '_1x1', cur, 1, c, ch(c1))
        b, cb = conv(name + '_3x3_bottleneck', cur, 1, c, ch(c3b))
        b, cb = conv(name + '_3x3', b, 3, cb, ch(c3))
        d, cd = conv(name + '_5x5_bottleneck', cur, 1, c, ch(c5b))
        d, cd = conv(name + '_5x5', d, 5, cd, ch(c5))
        e = pool(name + '_pool', cur, 1)
        e, ce = conv(name + '_pool_reduce', e, 1, c, ch(cp))
        nodes.append(Node(name + '/concat_dim', 'Const', [], {'value': np.asarray(3, np.int32)}))
        nodes.append(Node(name, 'Concat', [name + '/concat_dim', a, b, d, e], {'N': 4}))
        cur, c = name, ca + cb + cd + ce
        if upto == name:
            break
    if head_pool:
        nc = max(int(n_classes) // width_div, 4)
        nodes.append(Node('avgpool0', 'AvgPool', [cur], {'ksize': [1, head_pool, head_pool, 1], 'strides': [1, 1, 1, 1],
                                                          'padding': b'VALID'}))
        nodes.append(Node('avgpool0/reshape/shape', 'Const', [], {'value': np.asarray([-1, c], np.int32)}))
        nodes.append(Node('avgpool0/reshape', 'Reshape', ['avgpool0', 'avgpool0/reshape/shape'], {}))
        nodes.append(Node('softmax2_w', 'Const', [], {'value': (rng.randn(c, nc) * np.sqrt(1.0 / c)).astype(np.float32)}))
        nodes.append(Node('softmax2_b', 'Const', [], {'value': (0.05 * rng.randn(nc)).astype(np.float32)}))
        nodes.append(Node('softmax2_pre_activation/matmul', 'MatMul', ['avgpool0/reshape', 'softmax2_w'],
                          {'transpose_a': False, 'transpose_b': False}))
        nodes.append(Node('softmax2_pre_activation', 'BiasAdd', ['softmax2_pre_activation/matmul', 'softmax2_b'], {}))
    return nodes
