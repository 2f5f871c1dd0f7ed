"""Demo drivers -- the callers either side of the stylisation path, mirroring the reference scripts
``test_smokegun.py`` / ``test_chocolate.py`` / ``test_dambreak2d.py`` / ``test_smokegun_resim.py``:
``main(config)`` fills in the scene constants, ``run(config)`` loads the particle files, calls
``Styler(config).run(params)`` (or ``SimG2P``) and writes the results in the reference's formats
(``.bgeo`` particle sets, ``.npz`` fields with the y-flip, ``.png`` renders).

    python -m lnst.drivers.smokegun --data_dir data --num_frames 1 ...

File conventions shared by the drivers live here.
"""
import os

import numpy as np

from .. import partio


def frame_path(config, i):
    return os.path.join(config.data_dir, config.dataset, config.d_path % (config.target_frame + i))


def particle_range(config):
    """min / max particle count over the frames (test_smokegun.py:29-36)."""
    nmin, nmax = np.iinfo(np.int32).max, 0
    for i in range(config.num_frames):
        n = partio.read(frame_path(config, i)).numParticles()
        nmin, nmax = min(nmin, n), max(nmax, n)
    return nmin, nmax


def save_png(arr, path):
    from PIL import Image
    Image.fromarray(arr).save(path)


def save_loss_plot(loss_per_octave, log_dir):
    """loss_plot.png when matplotlib is present (test_smokegun.py:77-86); the raw curves always."""
    np.savez(os.path.join(log_dir, 'loss.npz'), **{'oct%d' % o: np.asarray(l) for o, l in enumerate(loss_per_octave)})
    try:
        import matplotlib
        matplotlib.use('Agg')
        import matplotlib.pyplot as plt
    except ImportError:
        return
    lb = []
    for o, l_ in enumerate(loss_per_octave):
        lb_, = plt.plot(range(len(l_)), l_, label='oct %d' % o)
        lb.append(lb_)
    plt.legend(handles=lb)
    plt.savefig(os.path.join(log_dir, 'loss_plot.png'))
    plt.close()


def save_renders(config, result):
    """final renders and the intermediate octave renders (test_smokegun.py:88-107)"""
    for i, r in enumerate(result.get('r') if result.get('r') is not None else []):
        save_png(r, os.path.join(config.log_dir, '%03d.png' % (config.target_frame + i)))
    for o, d_intm_o in enumerate(result['d_intm']):
        for i, d_intm_ in enumerate(d_intm_o):
            if d_intm_ is None:
                continue
            save_png(d_intm_, os.path.join(config.log_dir, 'o%02d_%03d.png' % (o, config.target_frame + i)))
