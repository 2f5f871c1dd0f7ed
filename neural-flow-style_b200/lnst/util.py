"""Host-side helpers on the stylisation path (the subset of the reference's ``util.py`` the
Styler loop touches: ``crop_ratio`` :176-185, ``resize`` :187-207, ``denoise`` :169-170,
``prepare_dirs_and_logger`` / ``save_config`` :404-425).  Visualisation helpers are out of
scope."""
import json
import os
import sys
from datetime import datetime

import numpy as np
from scipy import ndimage

from .config import str2bool  # noqa: F401  (re-exported like the reference's util.str2bool)


def octave_sizes(resolution, octave_n, octave_scale):
    """Coarse-to-fine grid sizes (``styler_3p.py:241-246``): res, res//s, ... reversed."""
    sizes, cur = [], np.array(resolution)
    for _ in range(octave_n):
        sizes.append([int(v) for v in cur])
        cur = (cur // octave_scale).astype(int)
    return sizes[::-1]


def crop_ratio(img, ratio):
    """Centre-crop ``img`` [h,w,...] to aspect ratio w/h = ``ratio``."""
    h, w = img.shape[:2]
    if w / float(h) > ratio:
        nh, nw = h, int(h * ratio)
    else:
        nh, nw = int(w / ratio), w
    assert nh <= h and nw <= w
    oy, ox = int((h - nh) * 0.5), int((w - nw) * 0.5)
    return img[oy:oy + nh, ox:ox + nw]


def resize(img, size, order=3):
    """Resample an image [h,w] or [h,w,c] to ``size`` = (H,W) for style/content targets.

    The reference calls skimage.transform.resize(order=3, mode='constant',
    anti_aliasing=True) after normalising to [0,1] (``util.py:187-207``).  skimage is not
    available offline; this does the same steps with scipy: Gaussian pre-filter with
    sigma = max(0,(scale-1)/2) per axis, then cubic-spline resampling on the half-pixel
    grid.  (skimage's own bicubic warp differs in the interpolant; parity unpinned --
    only the style TARGET depends on it, never the optimisation arithmetic.)
    """
    img = np.asarray(img, dtype=np.float32)
    if tuple(img.shape[:2]) == tuple(size[:2]):
        return img.copy()
    vmin, vmax = float(img.min()), float(img.max())
    norm = vmin < -1 or vmax > 1
    x = (img - vmin) / (vmax - vmin) if norm else img
    chans = [x] if x.ndim == 2 else [x[..., c] for c in range(x.shape[-1])]
    out = []
    for ch in chans:
        fac = [ch.shape[0] / float(size[0]), ch.shape[1] / float(size[1])]
        sig = [max(0.0, (f - 1) / 2) for f in fac]
        if any(s > 0 for s in sig):
            ch = ndimage.gaussian_filter(ch, sig, mode='constant', cval=0)
        yy = (np.arange(size[0]) + 0.5) * fac[0] - 0.5
        xx = (np.arange(size[1]) + 0.5) * fac[1] - 0.5
        grid = np.meshgrid(yy, xx, indexing='ij')
        out.append(ndimage.map_coordinates(ch, grid, order=order, mode='constant', cval=0).astype(np.float32))
    y = out[0] if x.ndim == 2 else np.stack(out, -1)
    return y * (vmax - vmin) + vmin if norm else y


def denoise(x, sigma):
    """``util.py:169-170`` (host version; the loop uses the lnst_temporal_gauss kernel)."""
    return ndimage.gaussian_filter(x, sigma=sigma)


def get_time():
    return datetime.now().strftime('%m%d_%H%M%S')


def save_config(config):
    path = os.path.join(config.log_dir, 'params.json')
    blob = {k: v for k, v in config.__dict__.items() if isinstance(v, (int, float, str, bool, list, type(None)))}
    with open(path, 'w') as fp:
        json.dump(blob, fp, indent=4, sort_keys=True)
    return path


def prepare_dirs_and_logger(config):
    """log/<dataset>/<MMDD_HHMMSS>_<tag>/params.json like the reference (no chdir)."""
    config.command = str(sys.argv)
    config.log_dir = os.path.join(config.log_dir, config.dataset, '%s_%s' % (get_time(), config.tag))
    os.makedirs(config.log_dir, exist_ok=True)
    save_config(config)
