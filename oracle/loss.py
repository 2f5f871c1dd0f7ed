"""Oracle restatement of ``StylerBase._loss`` and its helpers (``styler_base.py:96-231``).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  The histogram loss (``:187-209``) is out
of scope (w_hist = 0 by default, buggy at HEAD -- SURVEY.md section 2, row 3).
"""
import torch


def gram_matrix(x, batch_size):
    """``styler_base.py:96-102``: per-image F^T F with F = reshape(x[i], [h*w, C]); loops over
    ``range(batch_size)`` only -- extra view slices are ignored."""
    out = []
    for i in range(batch_size):
        f = x[i].reshape(-1, x.shape[-1])
        out.append(f.t() @ f)
    return torch.stack(out, 0)


def bicubic_legacy(x, out_h, out_w):
    """tf.compat.v1.image.resize(BICUBIC), align_corners=False, legacy coordinates
    (``styler_base.py:166``).  Keys kernel a=-0.75, src = dst*in/out, indices clamped.
    (TF quantises the fraction to a 1024-entry table; not restated -- parity unpinned.)"""
    B, H, W, C = x.shape
    dt = x.dtype
    a = -0.75

    def kern(t):
        t = t.abs()
        w1 = ((a + 2) * t - (a + 3)) * t * t + 1
        w2 = ((a * t - 5 * a) * t + 8 * a) * t - 4 * a
        return torch.where(t <= 1, w1, torch.where(t < 2, w2, torch.zeros_like(t)))

    def axis(n_in, n_out):
        src = torch.arange(n_out, dtype=dt) * (float(n_in) / float(n_out))
        base = torch.floor(src)
        idx, wts = [], []
        for o in (-1, 0, 1, 2):
            idx.append((base.to(torch.int64) + o).clamp(0, n_in - 1))
            wts.append(kern(src - (base + o)))
        return idx, wts

    yi, yw = axis(H, out_h)
    xi, xw = axis(W, out_w)
    rows = sum(x[:, yi[k]] * yw[k][None, :, None, None] for k in range(4))
    return sum(rows[:, :, xi[k]] * xw[k][None, None, :, None] for k in range(4))


def style_loss(features, style_features, w_style_layer, batch_size, d_gray=None,
               style_mask=False, style_mask_on_ref=False):
    """``styler_base.py:152-185``.  features/style_features: lists of [B,h,w,C]."""
    total = 0
    per_layer = []
    for f, fs, wl in zip(features, style_features, w_style_layer):
        denom = float(2 * f.shape[1] * f.shape[2] * f.shape[3])
        sdenom = float(2 * fs.shape[1] * fs.shape[2] * fs.shape[3])
        if style_mask:
            m = bicubic_legacy(d_gray, f.shape[1], f.shape[2])
            f = f * m
            area = m[..., 0].sum(dim=(1, 2), keepdim=True)
            denom = 2 * area * float(f.shape[3])
            if style_mask_on_ref:
                fs = fs * m
                sdenom = 2 * area * float(f.shape[3])
        g = gram_matrix(f, batch_size) / denom
        gs = gram_matrix(fs, batch_size) / sdenom
        l = ((g - gs) ** 2).sum()
        per_layer.append(l)
        total = total + wl * l
    return total, per_layer


def content_loss(feature, content_channel, content_feature=None, w_content_amp=100):
    """``styler_base.py:135-148``."""
    if content_feature is not None:
        return ((feature - content_feature * w_content_amp) ** 2).mean()
    if content_channel:
        c = content_channel
        return (-feature[..., c].mean() + feature[..., :c].abs().mean()
                + feature[..., c + 1:].abs().mean())
    return -feature.mean()


def tv_loss(d_img):
    """``styler_base.py:211-213``: mean over images of tf.image.total_variation (anisotropic
    L1, summed per image)."""
    dh = (d_img[:, 1:] - d_img[:, :-1]).abs().sum(dim=(1, 2, 3))
    dw = (d_img[:, :, 1:] - d_img[:, :, :-1]).abs().sum(dim=(1, 2, 3))
    return (dh + dw).mean()


def density_loss(d_vars):
    """``styler_base.py:217-223`` on the clipped variables ``self.d`` (``styler_3p.py:75``)."""
    d_loss, d_pres = 0, 0
    for d in d_vars:
        d_loss = d_loss + d.sum() ** 2
        d_pres = d_pres + (-torch.log(d.abs() + 1e-6)).sum()
    return d_loss + d_pres * 1e3


def pressure_loss(d_field):
    """``styler_3p.py:96-98`` + ``styler_base.py:228-230``: mean(where(d>0, d-1, 0)^2) on the
    pre-smoothing normalised SPH density."""
    pr = torch.where(d_field > 0, d_field - 1, torch.zeros_like(d_field))
    return (pr ** 2).mean()
