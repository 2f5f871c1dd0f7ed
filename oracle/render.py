"""Oracle restatement of the field post-processing and volume rendering that the reference
builds inline in ``styler_3p.py:112-164`` and ``styler_base.py:33-45``.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.
"""
import numpy as np
import torch
import torch.nn.functional as F


def smooth3(d, k):
    """``styler_3p.py:112-121``: conv3d with K = k1 x k1 x k1 / sum, k1=[1,k,1], stride 1,
    SAME zero padding.  d [B,D,H,W,1]."""
    if k <= 0:
        return d
    k1 = torch.tensor([1.0, float(k), 1.0], dtype=d.dtype)
    K = k1[:, None, None] * k1[None, :, None] * k1[None, None, :]
    K = (K / K.sum()).reshape(1, 1, 3, 3, 3)
    x = d.permute(0, 4, 1, 2, 3)
    y = F.conv3d(x, K, padding=1)
    return y.permute(0, 2, 3, 4, 1)


def field_post(d, k):
    """smooth + ``tf.maximum(d, 0)`` (``styler_3p.py:112-125``) = ``self.d_out``."""
    return torch.clamp(smooth3(d, k), min=0.0)


def render(d, transmit, render_liquid):
    """``styler_3p.py:148-158``.  d [B',D,H,W,C] -> [B',H,W,C].

    smoke : T_i = exp(-tau * sum_{j>=i} d_j) (inclusive reverse cumsum), I = sum_i d_i T_i,
            then I /= max(I) over the WHOLE tensor (``:158``; reduce_max splits its gradient
            equally among ties, as torch.max() over all elements... see below).
    liquid: I = 1 - exp(-tau * sum_i d_i), no normalisation.
    """
    if render_liquid:
        return 1.0 - torch.exp(-d.sum(dim=1) * transmit)
    cs = torch.flip(torch.cumsum(torch.flip(d, dims=[1]), dim=1), dims=[1])
    img = (d * torch.exp(-cs * transmit)).sum(dim=1)
    # torch.amax distributes the gradient evenly among ties, like TF's reduce_max gradient.
    return img / torch.amax(img)


def resize_bilinear_legacy(x, out_h, out_w):
    """tf.compat.v1.image.resize(BILINEAR), align_corners=False, no half-pixel centres
    (``styler_base.py:38``): src = dst * (in/out); lower=floor(src); upper=min(lower+1,in-1).
    x [B,H,W,C]."""
    B, H, W, C = x.shape
    dt = x.dtype

    def axis(n_in, n_out):
        scale = torch.tensor(float(n_in) / float(n_out), dtype=dt)
        src = torch.arange(n_out, dtype=dt) * scale
        lo = torch.floor(src).to(torch.int64).clamp(max=n_in - 1)
        hi = torch.clamp(lo + 1, max=n_in - 1)
        return lo, hi, src - lo.to(dt)

    ylo, yhi, yf = axis(H, out_h)
    xlo, xhi, xf = axis(W, out_w)
    top = x[:, ylo][:, :, xlo] + (x[:, ylo][:, :, xhi] - x[:, ylo][:, :, xlo]) * xf[None, None, :, None]
    bot = x[:, yhi][:, :, xlo] + (x[:, yhi][:, :, xhi] - x[:, yhi][:, :, xlo]) * xf[None, None, :, None]
    return top + (bot - top) * yf[None, :, None, None]


def resized_hw(h, w, resize_scale):
    """``styler_base.py:36-37``: int32(float32(scale) * float32(H))."""
    s = np.float32(resize_scale)
    return int(np.float32(h) * s), int(np.float32(w) * s)


def to_loss_net_input(d, resize_scale, target_field):
    """``styler_base.py:33-45``: optional bilinear up-scale, x255, gray -> RGB. Returns
    ``d_img`` [B,H',W',3] in 0..255."""
    if not np.isclose(resize_scale, 1):
        h, w = resized_hw(d.shape[1], d.shape[2], resize_scale)
        d = resize_bilinear_legacy(d, h, w)
    d = d * 255
    if 'c' not in target_field:
        d = torch.cat([d] * 3, dim=-1)
    return d
