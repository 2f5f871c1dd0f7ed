"""``skimage.transform.resize`` for the one case the golden runs use: target already at the
requested size.  skimage then applies no anti-aliasing blur (sigma = max(0, (factor-1)/2) = 0) and a
cubic warp on the identity grid, i.e. it returns the image (as float64).  Any other size raises:
the resampling filter itself is deviation D4 in DESIGN.md and is not pinned."""
import numpy as np


def resize(image, output_shape, order=1, mode='constant', anti_aliasing=True, **kw):
    output_shape = tuple(int(s) for s in output_shape)
    if tuple(image.shape[:len(output_shape)]) != output_shape:
        raise NotImplementedError('skimage stand-in: only same-size resize (got %s -> %s)'
                                  % (image.shape, output_shape))
    return np.asarray(image, dtype=np.float64)


def rescale(*a, **k):
    raise NotImplementedError('skimage stand-in: rescale')
