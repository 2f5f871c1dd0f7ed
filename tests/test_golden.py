"""Committed golden vectors (tests/golden/, made by tests/golden/make_golden.py).

CPU (-m "not gpu"): the oracle still reproduces them -- a drift guard on the checker itself -- and the
reference's own known-answer (the 5x5 warp tables of transform.py:1865-1884) holds for the oracle's
interpolation.  GPU (-m gpu): ``lnst.styler_3p.Styler.run`` through the C-ABI hits the same numbers
within the stated tolerances (fp32 loss net: loss 2e-4, field 2e-4 max, variables 2e-3 rel-L2;
tensor-core bf16 loss net: loss 2e-2, field 5e-2 max) on a box where neither the reference nor the
oracle inputs' generator need to exist.
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, 'golden')
sys.path.insert(0, GOLD)
import make_golden  # noqa: E402


def _load(name):
    return dict(np.load(os.path.join(GOLD, name + '.npz')))


def test_reference_warp_tables_hold_for_the_oracle():
    import oracle.transform as T
    kat = json.load(open(os.path.join(GOLD, 'warp_kat.json')))
    img = torch.tensor(kat['image'], dtype=torch.float32).reshape(1, 5, 5, 1)
    g = T.mgrid(5, 5).reshape(1, 2, -1)
    ident = T.interpolate(img, [g[:, 0], g[:, 1]]).reshape(5, 5)
    np.testing.assert_array_equal(ident.numpy(), np.asarray(kat['identity'], np.float32))
    z = 0.5 * g
    zoom = T.interpolate(img, [z[:, 0], z[:, 1]]).reshape(5, 5)
    np.testing.assert_allclose(zoom.numpy(), np.asarray(kat['zoom_in'], np.float32), rtol=0, atol=1e-5)


@pytest.mark.parametrize('name', ['density_allreduce', 'position_liquid'])
def test_oracle_reproduces_golden(name):
    want = _load(name)
    got = make_golden.run_case(name)
    np.testing.assert_allclose(got['l'], want['l'], rtol=1e-5)
    np.testing.assert_allclose(got['d'], want['d'], rtol=0, atol=1e-5 * np.abs(want['d']).max())
    assert np.linalg.norm(got['g_opt'] - want['g_opt']) <= 1e-4 * np.linalg.norm(want['g_opt'])


def _engine(name, conv_math):
    from helpers import smoke_cfg, liquid_cfg
    from lnst import synth
    from lnst.styler_3p import Styler
    kind, kw, n = make_golden.CASES[name]
    kw = dict(kw, conv_math=conv_math)
    sty = synth.style_image(kw['res'], kw['res'])
    if kind == '3d':
        p, r = synth.smoke_particles(n, 2, pad=4)
        st = Styler(smoke_cfg(**kw), weights=synth.vgg_weights())
        params = {'p': p, 'r': r}
    else:
        st = Styler(liquid_cfg(**kw), weights=synth.vgg_weights())
        params = {'p': synth.liquid_particles(n)}
    st.style_img = sty
    return st.run(params)


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['density_sequential', 'density_allreduce', 'density_tc_shapes', 'position_liquid'])
def test_cuda_fp32_path_hits_golden(name):
    want = _load(name)
    out = _engine(name, 'fp32')
    np.testing.assert_allclose(np.asarray(out['l']), want['l'], rtol=2e-4)
    assert np.abs(out['d'] - want['d']).max() <= 2e-4 * np.abs(want['d']).max()
    tol = 5e-3 if name == 'position_liquid' else 2e-3
    assert np.linalg.norm(out['g_opt'][0] - want['g_opt']) <= tol * np.linalg.norm(want['g_opt'])
    assert np.abs(out['r'].astype(int) - want['r'].astype(int)).max() <= 1


@pytest.mark.gpu
def test_cuda_tensor_core_path_hits_golden():
    want = _load('density_tc_shapes')
    out = _engine('density_tc_shapes', 'bf16')
    np.testing.assert_allclose(np.asarray(out['l']), want['l'], rtol=2e-2)
    assert np.abs(out['d'] - want['d']).max() <= 5e-2 * np.abs(want['d']).max()
