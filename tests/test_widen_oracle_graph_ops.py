"""The oracle's GraphDef operators (oracle/graphnet.py -- the part of the inception path no reference run can pin)
against brute-force NumPy loops written straight from TensorFlow's published definitions:

  Conv2D / MaxPool 'SAME':  out = ceil(in / stride), pad_total = max((out-1)*stride + k - in, 0), pad_before = pad_total // 2
  MaxPool: padding cells are ignored;  LRN: x / (bias + alpha * sum_{|j-c|<=r} x_j^2)^beta  (no 1/n)
  Concat(concat_dim, values...) / ConcatV2(values..., axis);  BiasAdd;  Relu
"""
import numpy as np
import pytest
import torch

from lnst.graphdef import Node
from oracle import graphnet as OG
from oracle import vgg as OV


def _same(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2


def np_conv(x, w, stride):
    n, H, W, ci = x.shape
    kh, kw, _, co = w.shape
    (OH, pt), (OW, pl) = _same(H, kh, stride), _same(W, kw, stride)
    y = np.zeros((n, OH, OW, co))
    for b in range(n):
        for oy in range(OH):
            for ox in range(OW):
                for ky in range(kh):
                    for kx in range(kw):
                        yy, xx = oy * stride + ky - pt, ox * stride + kx - pl
                        if 0 <= yy < H and 0 <= xx < W:
                            y[b, oy, ox] += x[b, yy, xx] @ w[ky, kx]
    return y


def np_maxpool(x, k, stride):
    n, H, W, c = x.shape
    (OH, pt), (OW, pl) = _same(H, k, stride), _same(W, k, stride)
    y = np.full((n, OH, OW, c), -np.inf)
    for oy in range(OH):
        for ox in range(OW):
            for ky in range(k):
                for kx in range(k):
                    yy, xx = oy * stride + ky - pt, ox * stride + kx - pl
                    if 0 <= yy < H and 0 <= xx < W:
                        y[:, oy, ox] = np.maximum(y[:, oy, ox], x[:, yy, xx])
    return y


def np_lrn(x, r, bias, alpha, beta):
    y = np.zeros_like(x)
    C = x.shape[-1]
    for c in range(C):
        lo, hi = max(c - r, 0), min(c + r, C - 1)
        y[..., c] = x[..., c] / (bias + alpha * (x[..., lo:hi + 1] ** 2).sum(-1)) ** beta
    return y


@pytest.mark.parametrize('H,W,k,stride', [(7, 8, 7, 2), (6, 5, 3, 1), (9, 9, 5, 1), (5, 7, 1, 1), (8, 6, 3, 2)])
def test_graph_ops_match_brute_force(H, W, k, stride):
    rng = np.random.RandomState(H * 10 + k)
    img = rng.uniform(0, 255, (2, H, W, 3))
    f32 = lambda a: a.astype(np.float32).astype(np.float64)   # the graph stores float32 constants
    w = f32(rng.randn(k, k, 3, 4) / k)
    b = f32(rng.randn(4))
    w2 = f32(rng.randn(1, 1, 4, 3))
    nodes = [Node('input', 'Placeholder'),
             Node('w', 'Const', [], {'value': w.astype(np.float32)}), Node('b', 'Const', [], {'value': b.astype(np.float32)}),
             Node('w2', 'Const', [], {'value': w2.astype(np.float32)}),
             Node('c/conv', 'Conv2D', ['input', 'w'], {'strides': [1, stride, stride, 1], 'padding': b'SAME'}),
             Node('c_pre', 'BiasAdd', ['c/conv', 'b']), Node('c', 'Relu', ['c_pre']),
             Node('pool', 'MaxPool', ['c'], {'ksize': [1, 3, 3, 1], 'strides': [1, 2, 2, 1], 'padding': b'SAME'}),
             Node('pool1', 'MaxPool', ['c'], {'ksize': [1, 3, 3, 1], 'strides': [1, 1, 1, 1], 'padding': b'SAME'}),
             Node('lrn', 'LRN', ['c'], {'depth_radius': 1, 'bias': 2.0, 'alpha': 1e-3, 'beta': 0.75}),
             Node('c2/conv', 'Conv2D', ['lrn', 'w2'], {'strides': [1, 1, 1, 1], 'padding': b'SAME'}),
             Node('axis', 'Const', [], {'value': np.asarray(3, np.int32)}),
             Node('cat', 'Concat', ['axis', 'c', 'pool1', 'c2/conv']),
             Node('cat2', 'ConcatV2', ['c2/conv', 'lrn', 'axis'])]
    got = OG.forward(torch.tensor(img), nodes, ['pool', 'cat', 'cat2'])
    x = img - np.asarray(torch.tensor(OV.MEAN_RGB, dtype=torch.float64))                         # vgg.preprocess feeds the graph (styler_base.py:56)
    c_pre = np_conv(x, w, stride) + b
    c = np.maximum(c_pre, 0)
    lrn = np_lrn(c, 1, 2.0, 1e-3, 0.75)
    c2 = np_conv(lrn, w2, 1)
    want = {'c_pre': c_pre, 'c': c, 'pool': np_maxpool(c, 3, 2), 'pool1': np_maxpool(c, 3, 1), 'lrn': lrn,
            'cat': np.concatenate([c, np_maxpool(c, 3, 1), c2], -1), 'cat2': np.concatenate([c2, lrn], -1)}
    for name, v in want.items():
        np.testing.assert_allclose(got[name].numpy(), v, rtol=1e-9, atol=1e-9, err_msg=name)
    assert got['input'] is not None and got['input'].shape == (2, H, W, 3)      # layer 'input' is d_img (styler_base.py:92)


def test_head_ops_match_brute_force():
    """AvgPool (VALID and SAME: padding cells are not counted), Reshape, MatMul"""
    rng = np.random.RandomState(3)
    img = rng.uniform(0, 255, (2, 5, 6, 3))
    w = rng.randn(3, 4).astype(np.float32)
    nodes = [Node('input', 'Placeholder'),
             Node('ap', 'AvgPool', ['input'], {'ksize': [1, 2, 2, 1], 'strides': [1, 1, 1, 1], 'padding': b'VALID'}),
             Node('aps', 'AvgPool', ['input'], {'ksize': [1, 3, 3, 1], 'strides': [1, 2, 2, 1], 'padding': b'SAME'}),
             Node('shape', 'Const', [], {'value': np.asarray([-1, 3], np.int32)}),
             Node('flat', 'Reshape', ['ap', 'shape']), Node('w', 'Const', [], {'value': w}),
             Node('mm', 'MatMul', ['flat', 'w'], {'transpose_a': False, 'transpose_b': False})]
    got = OG.forward(torch.tensor(img), nodes, ['mm', 'aps'])
    x = img - np.asarray(torch.tensor(OV.MEAN_RGB, dtype=torch.float64))
    ap = np.zeros((2, 4, 5, 3))
    for oy in range(4):
        for ox in range(5):
            ap[:, oy, ox] = x[:, oy:oy + 2, ox:ox + 2].mean(axis=(1, 2))
    np.testing.assert_allclose(got['ap'].numpy(), ap, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(got['mm'].numpy(), ap.reshape(-1, 3) @ w.astype(np.float64), rtol=1e-9, atol=1e-9)
    (OH, pt), (OW, pl) = _same(5, 3, 2), _same(6, 3, 2)
    aps = np.zeros((2, OH, OW, 3))
    for oy in range(OH):
        for ox in range(OW):
            y0, x0 = oy * 2 - pt, ox * 2 - pl
            aps[:, oy, ox] = x[:, max(y0, 0):min(y0 + 3, 5), max(x0, 0):min(x0 + 3, 6)].mean(axis=(1, 2))
    np.testing.assert_allclose(got['aps'].numpy(), aps, rtol=1e-9, atol=1e-9)


def test_pool1_only_touches_the_first_convolution():
    rng = np.random.RandomState(0)
    img = torch.tensor(rng.uniform(0, 255, (1, 9, 9, 3)))
    w = rng.randn(7, 7, 3, 2).astype(np.float32)
    nodes = [Node('input', 'Placeholder'), Node('w', 'Const', [], {'value': w}),
             Node('conv2d0_pre_relu/conv', 'Conv2D', ['input', 'w'], {'strides': [1, 2, 2, 1], 'padding': b'SAME'}),
             Node('other/conv', 'Conv2D', ['input', 'w'], {'strides': [1, 2, 2, 1], 'padding': b'SAME'})]
    a = OG.forward(img, nodes, ['conv2d0_pre_relu/conv', 'other/conv'], pool1=True)
    assert a['conv2d0_pre_relu/conv'].shape[1:3] == (9, 9) and a['other/conv'].shape[1:3] == (5, 5)    # styler_base.py:26-31
