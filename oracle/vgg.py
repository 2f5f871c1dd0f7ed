"""Oracle restatement of the reference's slim VGG-16/19 loss network (``vgg.py``).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Weights are in slim layout (HWIO) in a
dict ``{'conv1_1': (w[3,3,Cin,Cout], b[Cout]), ...}``; the pretrained checkpoint is not
available offline, so tests/bench use seeded synthetic weights (``synthetic_weights``).
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# vgg.py:16-18 -- mean only, no std (``vgg.py:50-53``)
MEAN_RGB = (0.485 * 255, 0.456 * 255, 0.406 * 255)

_CFG = {
    'vgg_19': [(2, 64), (2, 128), (4, 256), (4, 512), (4, 512)],   # vgg.py:89-113
    'vgg_16': [(2, 64), (2, 128), (3, 256), (3, 512), (3, 512)],   # vgg.py:68-87
}


def layer_specs(model='vgg_19'):
    """[(name, cin, cout)] for the conv layers in network order."""
    out, cin = [], 3
    for b, (rep, cout) in enumerate(_CFG[model], start=1):
        for i in range(1, rep + 1):
            out.append(('conv%d_%d' % (b, i), cin, cout))
            cin = cout
    return out


def synthetic_weights(model='vgg_19', seed=19, dtype=torch.float32):
    """Seeded He-normal HWIO weights + N(0,1) biases (CPU generator => machine independent)."""
    g = torch.Generator().manual_seed(seed)
    w = OrderedDict()
    for name, cin, cout in layer_specs(model):
        std = float(np.sqrt(2.0 / (9 * cin)))
        wt = torch.randn(3, 3, cin, cout, generator=g, dtype=torch.float32) * std
        bs = torch.randn(cout, generator=g, dtype=torch.float32)
        w[name] = (wt.to(dtype), bs.to(dtype))
    return w


def preprocess(images):
    """``vgg.py:50-53``."""
    return images - torch.tensor(MEAN_RGB, dtype=images.dtype)


def forward(d_img, weights, model='vgg_19', upto=None):
    """``vgg.py:68-113`` + ``load_vgg`` (``:115-120``): 3x3 SAME conv + bias + ReLU, 2x2/2 VALID
    **average** pool between blocks.  End points: ``conv{b}_{i}`` (post-ReLU) and ``pool{b}``,
    plus 'input' = d_img (``styler_base.py:92``).  d_img [B,H,W,3] in 0..255.

    ``upto``: stop after this end point (an optimisation only; the reference builds all).
    """
    ep = OrderedDict()
    ep['input'] = d_img
    x = preprocess(d_img).permute(0, 3, 1, 2)
    for b, (rep, _) in enumerate(_CFG[model], start=1):
        for i in range(1, rep + 1):
            name = 'conv%d_%d' % (b, i)
            w, bias = weights[name]
            x = F.relu(F.conv2d(x, w.to(x.dtype).permute(3, 2, 0, 1), bias.to(x.dtype), padding=1))
            ep[name] = x.permute(0, 2, 3, 1)
            if upto == name:
                return ep
        x = F.avg_pool2d(x, 2, 2)
        ep['pool%d' % b] = x.permute(0, 2, 3, 1)
        if upto == 'pool%d' % b:
            return ep
    return ep
