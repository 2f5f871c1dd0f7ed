"""Oracle interpretation of a frozen TensorFlow GraphDef loss network (inception5h), reference
``styler_base.py:17-31,53-57,91-94``: ``tf.import_graph_def(graph_def, {'input': vgg.preprocess(d)})`` and layers
read by tensor name.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  PARITY: **unpinned** for the network itself -- neither
``tensorflow_inception_graph.pb`` nor TensorFlow is available here, so the op semantics below are TF's published
ones (SAME padding with the odd cell after, MaxPoolGrad to the first maximum = torch's rule, LRN with alpha not
divided by the window) and the tests run on a seeded synthetic graph of the inception5h topology
(``lnst.synth.inception5h_nodes``).  Every op is a plain torch op, gradients come from autograd.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import vgg as V


def _clean(name):
    name = name[7:] if name.startswith('import/') else name
    name = name[1:] if name.startswith('^') else name
    return name.split(':')[0]


def _same_pad(x, k, stride, value=0.0):
    H, W = x.shape[-2:]
    pads = []
    for size in (W, H):                                        # F.pad takes the last axis first
        out = -(-size // stride)
        total = max((out - 1) * stride + k - size, 0)
        pads += [total // 2, total - total // 2]
    return F.pad(x, pads, value=value)


def forward(d_img, nodes, wanted, pool1=False):
    """d_img [B,H,W,3] in 0..255 -> {name: tensor NHWC} for every node needed by ``wanted``."""
    by_name = {n.name: n for n in nodes}
    const = {n.name: n.attr['value'] for n in nodes if n.op == 'Const'}
    dt = d_img.dtype
    vals = {'input': V.preprocess(d_img)}                      # styler_base.py:56

    def get(name):
        name = _clean(name)
        if name in vals:
            return vals[name]
        if name in const:
            return torch.as_tensor(np.asarray(const[name])).to(dt) if np.asarray(const[name]).dtype.kind == 'f' \
                else const[name]
        n = by_name[name]
        a = n.attr
        if n.op == 'Conv2D':
            strides = list(a.get('strides', [1, 1, 1, 1]))
            if pool1 and 'conv2d0_pre_relu/conv' in n.name:    # styler_base.py:26-31
                strides[1:3] = [1, 1]
            x, w = get(n.inputs[0]).permute(0, 3, 1, 2), get(n.inputs[1])
            if (a.get('padding') or b'SAME') == b'SAME':
                x = _same_pad(x, w.shape[0], strides[1])
            y = F.conv2d(x, w.permute(3, 2, 0, 1), None, stride=strides[1]).permute(0, 2, 3, 1)
        elif n.op == 'BiasAdd':
            y = get(n.inputs[0]) + get(n.inputs[1])
        elif n.op == 'Relu':
            y = F.relu(get(n.inputs[0]))
        elif n.op == 'MaxPool':
            k, s = list(a['ksize'])[1], list(a['strides'])[1]
            x = get(n.inputs[0]).permute(0, 3, 1, 2)
            if (a.get('padding') or b'SAME') == b'SAME':
                x = _same_pad(x, k, s, value=float('-inf'))
            y = F.max_pool2d(x, k, s).permute(0, 2, 3, 1)
        elif n.op == 'LRN':
            r, bias = int(a.get('depth_radius', 5)), float(a.get('bias', 1.0))
            alpha, beta = float(a.get('alpha', 1.0)), float(a.get('beta', 0.5))
            x = get(n.inputs[0])
            sq = F.pad(x * x, (r, r))
            s = sum(sq[..., i:i + x.shape[-1]] for i in range(2 * r + 1))
            y = x * (bias + alpha * s) ** (-beta)
        elif n.op in ('Concat', 'ConcatV2'):
            parts = n.inputs[1:] if n.op == 'Concat' else n.inputs[:-1]
            axis = int(np.asarray(const[_clean(n.inputs[0] if n.op == 'Concat' else n.inputs[-1])]).reshape(-1)[0])
            y = torch.cat([get(p) for p in parts], dim=axis)
        elif n.op == 'Identity':
            y = get(n.inputs[0])
        elif n.op == 'AvgPool':                                # SAME padding cells are not counted
            k, s = list(a['ksize'])[1], list(a['strides'])[1]
            x = get(n.inputs[0]).permute(0, 3, 1, 2)
            if (a.get('padding') or b'VALID') == b'SAME':
                ones = _same_pad(torch.ones(1, 1, x.shape[2], x.shape[3], dtype=x.dtype), k, s)
                y = F.avg_pool2d(_same_pad(x, k, s), k, s, divisor_override=1) / F.avg_pool2d(ones, k, s, divisor_override=1)
            else:
                y = F.avg_pool2d(x, k, s)
            y = y.permute(0, 2, 3, 1)
        elif n.op == 'Reshape':
            y = get(n.inputs[0]).reshape([int(v) for v in np.asarray(const[_clean(n.inputs[1])]).reshape(-1)])
        elif n.op == 'MatMul':
            y = get(n.inputs[0]) @ get(n.inputs[1])
        else:
            raise NotImplementedError(n.op)
        vals[name] = y
        return y

    for w in wanted:
        if 'input' not in w:
            get(w)
    vals['input'] = d_img                                      # styler_base.py:92: layer 'input' is d_img itself
    return vals
