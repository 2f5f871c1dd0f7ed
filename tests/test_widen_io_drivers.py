"""Particle file I/O (``lnst/partio.py``: the slice of the partio API the reference drivers call, BGEO v5)
and the demo drivers either side of the hot path (``lnst/drivers``: reference ``test_smokegun.py``,
``test_chocolate.py``, ``test_dambreak2d.py``, ``test_smokegun_resim.py``)."""
import gzip
import os
import struct

import numpy as np
import pytest
import torch

from helpers import smoke_cfg, liquid_cfg, dam_cfg
from lnst import partio, synth


def _demo_set(n=37, seed=0):
    rng = np.random.RandomState(seed)
    pt = partio.create()
    pid = pt.addAttribute('id', partio.INT, 1)
    pos = pt.addAttribute('position', partio.VECTOR, 3)
    den = pt.addAttribute('density', partio.VECTOR, 2)
    rad = pt.addAttribute('radius', partio.FLOAT, 1)
    for i in range(n):                                          # the drivers' per-particle loop (test_smokegun_resim.py:304-316)
        j = pt.addParticle()
        pt.set(pid, j, (int(n - 1 - i),))
        pt.set(pos, j, tuple(rng.rand(3).astype(float)))
        pt.set(den, j, tuple(rng.rand(2).astype(float)))
        pt.set(rad, j, (0.5,))
    return pt


@pytest.mark.parametrize('compressed', [True, False])
def test_bgeo_round_trip(tmp_path, compressed):
    pt = _demo_set()
    path = str(tmp_path / 'a.bgeo')
    partio.write(path, pt, compressed=compressed)
    back = partio.read(path)
    assert back.numParticles() == pt.numParticles() == 37
    assert back.numAttributes() == 4
    for name in ('id', 'position', 'density', 'radius'):
        a, b = pt.attributeInfo(name), back.attributeInfo(name)
        assert (a.type, a.count) == (b.type, b.count), name
        np.testing.assert_array_equal(pt.array(name), back.array(name))
    assert back.get(back.attributeInfo('id'), 0) == (36,)
    assert isinstance(back.get(back.attributeInfo('position'), 3)[0], float)


def test_bgeo_layout_is_houdini_v5(tmp_path):
    """Header fields, attribute table and the first point record, byte for byte (BGEO.cpp layout)."""
    pt = partio.create()
    pt.addAttribute('position', partio.VECTOR, 3)
    pt.addAttribute('id', partio.INT, 1)
    pt.setArray('position', np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], np.float32))
    pt.setArray('id', np.array([7, 9], np.int32))
    path = str(tmp_path / 'b.bgeo')
    partio.write(path, pt)
    blob = gzip.open(path, 'rb').read()
    assert blob[:4] == b'Bgeo' and blob[4:5] == b'V'
    version, n, nprims, npg, nprg, npa, nva, npra, na = struct.unpack_from('>9i', blob, 5)
    assert (version, n, nprims, npg, nprg, npa, nva, npra, na) == (5, 2, 1, 0, 0, 1, 0, 1, 0)
    off = 5 + 36
    assert struct.unpack_from('>h', blob, off)[0] == 2 and blob[off + 2:off + 4] == b'id'
    size, htype, default = struct.unpack_from('>Hii', blob, off + 4)
    assert (size, htype, default) == (1, 1, 0)
    off += 4 + 10
    assert struct.unpack_from('>4fi', blob, off) == (1.0, 2.0, 3.0, 1.0, 7)
    assert struct.unpack_from('>4fi', blob, off + 20) == (4.0, 5.0, 6.0, 1.0, 9)
    assert blob[-2:] == b'\x00\xff'


def test_bgeo_empty_and_errors(tmp_path):
    pt = partio.create()
    pt.addAttribute('position', partio.VECTOR, 3)
    path = str(tmp_path / 'e.bgeo')
    partio.write(path, pt)
    assert partio.read(path).numParticles() == 0
    with pytest.raises(ValueError):
        partio.write(str(tmp_path / 'x.ptc'), pt)
    bad = tmp_path / 'bad.bgeo'
    bad.write_bytes(b'not a bgeo file at all')
    with pytest.raises(ValueError):
        partio.read(str(bad))


# ---- drivers (engine on the CPU interpreter here, on the B200 with -m gpu) ---------------------------
def _write_frames(root, dataset, d_path, frames, domain3, dens=None, dim=3):
    os.makedirs(os.path.dirname(os.path.join(root, dataset, d_path % 0)), exist_ok=True)
    for t, p in enumerate(frames):
        ok = p[:, 0] >= 0
        pp = p[ok]
        pt = partio.create()
        pt.addAttribute('id', partio.INT, 1)
        pt.addAttribute('position', partio.VECTOR, 3)
        n = pp.shape[0]
        pt.setArray('id', np.arange(n, dtype=np.int32))
        if dim == 3:                                            # files hold (x,y,z) in domain units
            xyz = np.stack([pp[:, 2] * domain3[2], pp[:, 1] * domain3[1], pp[:, 0] * domain3[0]], -1)
        else:
            xyz = np.stack([pp[:, 1] * domain3[1], pp[:, 0] * domain3[0], np.zeros(n)], -1)
        pt.setArray('position', xyz)
        if dens is not None:
            k = dens[t].shape[1]
            pt.addAttribute('density', partio.VECTOR if k > 1 else partio.FLOAT, k)
            pt.setArray('density', dens[t][ok])
        partio.write(os.path.join(root, dataset, d_path % t), pt)


def _common(cfg, tmp_path, dataset, d_path):
    from PIL import Image
    cfg.data_dir, cfg.log_dir = str(tmp_path / 'data'), str(tmp_path / 'log')
    cfg.dataset, cfg.d_path, cfg.target_frame, cfg.tag = dataset, d_path, 0, 'drv'
    os.makedirs(os.path.join(cfg.data_dir, 'image'), exist_ok=True)
    hw = cfg.resolution[-2:]
    cfg.style_target = os.path.join(cfg.data_dir, 'image', 'style.png')
    Image.fromarray(synth.style_image(hw[0], hw[1]).astype(np.uint8)).save(cfg.style_target)
    return cfg


def test_smokegun_driver(dev, tmp_path):
    from lnst.drivers import smokegun
    from lnst.styler_3p import Styler
    cfg = _common(smoke_cfg(res=12, iter=2, conv_math='fp32', style_layer=['conv1_2'], w_style_layer=[1.0]),
                  tmp_path, 'smokegun', 'pt/%03d.bgeo')
    p, r = synth.smoke_particles(500, 2, pad=6)
    _write_frames(cfg.data_dir, cfg.dataset, cfg.d_path, p, cfg.domain, dens=r)
    p_l, r_l = smokegun.load_particles(cfg)
    assert p_l[0].shape == (500, 3)                      # the 6 padding rows are not written to the file
    np.testing.assert_allclose(p_l[0], p[0][p[0][:, 0] >= 0], atol=1e-6)
    np.testing.assert_array_equal(r_l[0], r[0][p[0][:, 0] >= 0])
    out = smokegun.run(cfg, weights=synth.vgg_weights())
    files = sorted(os.listdir(cfg.log_dir))
    assert '000.png' in files and '000.npz' in files and 'params.json' in files and 'loss.npz' in files
    with np.load(os.path.join(cfg.log_dir, '000.npz')) as f:     # field stored with the y-flip (test_smokegun.py:94-97)
        np.testing.assert_array_equal(f['x'], out['d'][0][:, ::-1])
    # same numbers as calling the Styler directly on the same arrays
    from PIL import Image
    cfg2 = smoke_cfg(res=12, iter=2, conv_math='fp32', style_layer=['conv1_2'], w_style_layer=[1.0])
    direct = Styler(cfg2, weights=synth.vgg_weights())
    direct.style_img = np.float32(Image.open(cfg.style_target))
    ref = direct.run({'p': p_l, 'r': r_l})
    np.testing.assert_allclose(out['l'], ref['l'], rtol=1e-4)      # atomics reorder sums between two runs


def test_chocolate_driver_writes_particles(dev, tmp_path):
    from lnst.drivers import chocolate
    cfg = _common(liquid_cfg(res=12, iter=2, conv_math='fp32', style_layer=['conv1_2'], w_style_layer=[1.0]),
                  tmp_path, 'chocolate', 'partio/ParticleData_Fluid_%d.bgeo')
    p = synth.liquid_particles(300, num_frames=2)
    cfg.num_frames = 2
    _write_frames(cfg.data_dir, cfg.dataset, cfg.d_path, p, cfg.domain)
    out = chocolate.run(cfg, weights=synth.vgg_weights())
    for t in range(2):
        back = partio.read(os.path.join(cfg.log_dir, '%03d.bgeo' % t))
        want = out['p'][t]
        want = want[want[:, 2] >= 0]
        assert back.numParticles() == want.shape[0]
        xyz = back.array('position')
        np.testing.assert_allclose(xyz[:, 0], want[:, 2] * cfg.domain[2], rtol=1e-6)
        np.testing.assert_allclose(xyz[:, 2], want[:, 0] * cfg.domain[0], rtol=1e-6)
        np.testing.assert_allclose(back.array('radius'), cfg.radius)


def test_dambreak2d_driver(dev, tmp_path):
    from lnst.drivers import dambreak2d
    cfg = _common(dam_cfg(iter=2), tmp_path, 'dambreak2d', 'partio/ParticleData_Fluid_%d.bgeo')
    p, r = synth.dam_particles_2d(cfg.domain)
    _write_frames(cfg.data_dir, cfg.dataset, cfg.d_path, p, cfg.domain, dens=r, dim=2)
    out = dambreak2d.run(cfg, weights=synth.vgg_weights())
    back = partio.read(os.path.join(cfg.log_dir, '000.bgeo'))
    np.testing.assert_allclose(back.array('Cd'), out['c'][0], atol=1e-7)
    np.testing.assert_allclose(back.array('position')[:, 0], p[0][:, 1] * cfg.domain[1], rtol=1e-5)
    assert os.path.exists(os.path.join(cfg.log_dir, '000.png'))


def test_resim_driver_feeds_the_smoke_driver(dev, tmp_path):
    """grid simulation -> resim driver -> .bgeo particle sets -> smoke driver's loader: the two sides of the path."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    import make_resim_golden as G
    from lnst.drivers import smokegun_resim, smokegun
    c, ds, us = G.resim_inputs(n_frames=2)
    c.data_dir, c.log_dir, c.dataset, c.tag = str(tmp_path / 'data'), str(tmp_path / 'log'), 'smokegun', 'rs'
    c.d_path, c.v_path, c.target_frame, c.num_frames, c.resampling = 'd_low/%03d.npz', 'v_low/%03d.npz', 0, 2, True
    c.src_region = G.SRC_REGION
    for sub in ('d_low', 'v_low'):
        os.makedirs(os.path.join(c.data_dir, c.dataset, sub))
    for t in range(2):
        np.savez(os.path.join(c.data_dir, c.dataset, c.d_path % t), x=ds[t][:, ::-1])
        np.savez(os.path.join(c.data_dir, c.dataset, c.v_path % t), x=np.zeros(ds[t].shape + (3,), np.float32))
    p, p_id = smokegun_resim.run(c)
    pt = partio.read(os.path.join(c.log_dir, '001.bgeo'))
    assert pt.numParticles() == p.shape[0] and pt.attributeInfo('density').count == c.octave_n
    np.testing.assert_array_equal(pt.array('id')[:, 0], p_id)
    np.testing.assert_allclose(pt.array('position')[:, 1], (1 - p[:, 1]) * c.domain[1], rtol=1e-5, atol=1e-5)
    assert os.path.exists(os.path.join(c.log_dir, '001.png')) and os.path.exists(os.path.join(c.log_dir, 'stat.txt'))
    # the smoke driver reads what the resim driver wrote
    cfg = smoke_cfg(res=12)
    cfg.resolution, cfg.domain = c.resolution, c.domain
    cfg.data_dir, cfg.dataset, cfg.d_path, cfg.target_frame, cfg.num_frames = os.path.dirname(c.log_dir), \
        os.path.basename(c.log_dir), '%03d.bgeo', 0, 2
    cfg.num_kernels = c.octave_n
    p_l, r_l = smokegun.load_particles(cfg)
    assert p_l[1].shape == (p.shape[0], 3) and r_l[1].shape == (p.shape[0], 2)


def test_mac_velocity_conversion():
    from lnst.drivers.smokegun_resim import centred_velocity, normalised_velocity
    rng = np.random.RandomState(0)
    v = rng.randn(3, 4, 5, 3)
    c = centred_velocity(v)
    assert c.shape == v.shape
    # x component: mean of a cell's two x-faces, last face closed (zero); H axis flipped
    np.testing.assert_allclose(c[1, 3 - 2, 1, 0], 0.5 * (v[1, 2, 1, 0] + v[1, 2, 2, 0]))
    np.testing.assert_allclose(c[1, 3 - 2, 4, 0], 0.5 * v[1, 2, 4, 0])
    u = normalised_velocity(c, 2.0)
    np.testing.assert_allclose(u[..., 0], c[..., 2] / 3 * 2.0)
    np.testing.assert_allclose(u[..., 1], -c[..., 1] / 4 * 2.0)


# ---- scene constants: lnst.drivers.*.main against the reference scripts' own main() (captured by
# tests/golden/make_driver_config_golden.py from the unmodified scripts) --------------------------------------
ENGINE_ONLY = {'view_mode', 'conv_math', 'content_network', 'command', 'log_dir'}


@pytest.mark.parametrize('ref_name,mod_name', [('test_smokegun', 'smokegun'), ('test_chocolate', 'chocolate'),
                                               ('test_dambreak2d', 'dambreak2d'), ('test_smokegun_resim', 'smokegun_resim')])
def test_driver_main_sets_the_reference_scene_constants(ref_name, mod_name, monkeypatch):
    import importlib
    import json
    from lnst.config import get_config
    want = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_driver_configs.json')))[ref_name]
    mod = importlib.import_module('lnst.drivers.' + mod_name)
    got = {}
    monkeypatch.setattr(mod, 'run', lambda cfg, *a, **k: got.update(vars(cfg)))
    cfg, _ = get_config([])
    mod.main(cfg)
    diffs = []
    for k, v in want.items():
        if k in ENGINE_ONLY:
            continue
        g = got.get(k, '<missing>')
        g = list(g) if isinstance(g, tuple) else g
        same = (g == v) or (isinstance(v, float) and isinstance(g, (int, float)) and abs(g - v) <= 1e-12 * max(1, abs(v))) \
            or (isinstance(v, list) and isinstance(g, list) and len(g) == len(v) and
                all(a == b or (isinstance(a, (int, float)) and isinstance(b, (int, float)) and abs(a - b) < 1e-9)
                    for a, b in zip(g, v)))
        if not same:
            diffs.append((k, g, v))
    assert not diffs, diffs


def test_dambreak2d_main_runs_with_default_flags(dev, tmp_path, monkeypatch):
    """main() sets style_mask=True on vgg_19 and leaves conv_math at the engine default; the Styler must
    pick the fp32 loss-net path (with a warning) instead of raising, and the whole driver must run end to end."""
    import warnings
    from lnst.config import get_config
    from lnst.drivers import dambreak2d
    cfg, _ = get_config([])
    cfg.keep_resolution, cfg.resolution = True, [6, 8]             # base grid; main() scales it by 4
    cfg.data_dir, cfg.log_dir = str(tmp_path / 'data'), str(tmp_path / 'log')
    cfg.num_frames, cfg.target_frame = 1, 0
    from PIL import Image
    os.makedirs(os.path.join(cfg.data_dir, 'image'), exist_ok=True)
    cfg.style_target = os.path.join(cfg.data_dir, 'image', 'style.png')
    Image.fromarray(np.uint8(synth.style_image(32, 32))).save(cfg.style_target)
    real_run = dambreak2d.run

    def short_run(c, weights=None):
        c.iter, c.octave_n = 2, 1                                  # scene constants stay main()'s, the budget shrinks
        p, r = synth.dam_particles_2d(c.domain)
        _write_frames(c.data_dir, c.dataset, c.d_path, p, c.domain, dens=r, dim=2)
        return real_run(c, weights=synth.vgg_weights())
    monkeypatch.setattr(dambreak2d, 'run', short_run)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        out = dambreak2d.main(cfg)
    assert any("'fp32'" in str(x.message) for x in w)
    assert cfg.style_mask and len(out['l'][0]) == 2 and np.isfinite(out['l'][0]).all()
