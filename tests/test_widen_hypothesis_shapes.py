"""Hypothesis-driven shapes for the widening-row kernels on the CPU interpreter of the kernel sources: odd sizes,
one-pixel images, more taps than pixels, empty particle sets, channel counts around the register window / tile edges
(SURVEY.md section 4, item 3).  CPU only -- the GPU runs the fixed-shape versions of these tests."""
import numpy as np
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from lnst import _lib, ops
from oracle import transform as T
from test_widen_graphnet_kernels import ref_conv, ref_maxpool, ref_lrn

CFG = dict(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


def _use_emu(emu_lib):
    prev = _lib._lib
    _lib.set_for_testing(emu_lib)
    return prev


def _close(got, want, tol):
    got, want = got.detach().double(), want.detach().double()
    assert got.shape == want.shape
    scale = float(want.abs().max()) if want.numel() else 0.0
    assert float((got - want).abs().max()) <= tol * max(scale, 1e-30) if want.numel() else True


@settings(**CFG)
@given(n=st.integers(1, 2), H=st.integers(1, 9), W=st.integers(1, 9), cin=st.integers(1, 5), cout=st.integers(1, 5),
       k=st.sampled_from([1, 3, 5, 7]), stride=st.sampled_from([1, 2]), seed=st.integers(0, 1000))
def test_conv2d_any_shape(emu_lib, n, H, W, cin, cout, k, stride, seed):
    prev = _use_emu(emu_lib)
    try:
        rng = np.random.RandomState(seed)
        x = torch.tensor(rng.randn(n, H, W, cin).astype(np.float32), requires_grad=True)
        w = torch.tensor(rng.randn(k, k, cin, cout).astype(np.float32))
        b = torch.tensor(rng.randn(cout).astype(np.float32))
        want = ref_conv(x, w, b, stride, 'SAME', relu=True)
        got = ops.conv2d_f32(x.detach(), w, b, stride, 'SAME', relu=True)
        _close(got, want, 5e-6)
        g = torch.tensor(rng.randn(*want.shape).astype(np.float32))
        (want * g).sum().backward()
        gx = torch.empty(x.shape)
        ops.conv2d_bwd_data_f32(g, w, x.shape, stride, 'SAME', gx, accumulate=False, relu_y=got)
        _close(gx, x.grad, 5e-6)
    finally:
        _lib.set_for_testing(prev)


@settings(**CFG)
@given(H=st.integers(1, 8), W=st.integers(1, 8), C=st.integers(1, 6), stride=st.sampled_from([1, 2]), r=st.integers(0, 3),
       seed=st.integers(0, 1000))
def test_pool_and_lrn_any_shape(emu_lib, H, W, C, stride, r, seed):
    prev = _use_emu(emu_lib)
    try:
        rng = np.random.RandomState(seed)
        xv = rng.randn(1, H, W, C).astype(np.float32)
        xv[xv < 0] = 0
        x = torch.tensor(xv, requires_grad=True)
        want = ref_maxpool(x, 3, stride)
        np.testing.assert_array_equal(ops.maxpool_fwd(x.detach(), 3, stride).numpy(), want.detach().numpy())
        g = torch.tensor(rng.randn(*want.shape).astype(np.float32))
        (want * g).sum().backward()
        gx = torch.empty(x.shape)
        ops.maxpool_bwd(g, x.detach(), 3, stride, 'SAME', gx, accumulate=False)
        _close(gx, x.grad, 3e-6)
        y = ops.lrn_fwd(x.detach(), r, 1.0, 0.01, 0.75)
        _close(y, ref_lrn(x.detach(), r, 1.0, 0.01, 0.75), 3e-6)
    finally:
        _lib.set_for_testing(prev)


@settings(**CFG)
@given(dims=st.lists(st.integers(1, 6), min_size=2, max_size=3), C=st.integers(1, 9), n=st.integers(0, 40),
       linear=st.booleans(), seed=st.integers(0, 1000))
def test_g2p_any_shape(emu_lib, dims, C, n, linear, seed):
    prev = _use_emu(emu_lib)
    try:
        rng = np.random.RandomState(seed)
        dim = len(dims)
        g = torch.tensor(rng.randn(*dims, C).astype(np.float32))
        p = torch.tensor(rng.uniform(-0.3, 1.3, (n, dim)).astype(np.float32))      # well outside the grid too
        got = ops.g2p(g, p, linear=linear)
        assert got.shape == (n, C)
        if n:
            want = T.g2p(g[None], p[None], is_2d=dim == 2, is_linear=linear)[0]
            _close(got, want, 5e-6)
            x_adv = ops.rk4_advect(torch.tensor(rng.uniform(-0.1, 0.1, dims + [dim]).astype(np.float32)), p, 0.5,
                                   linear=linear)
            assert x_adv.shape == p.shape and bool(torch.isfinite(x_adv).all())
    finally:
        _lib.set_for_testing(prev)
