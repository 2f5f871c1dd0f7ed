"""Hypothesis-driven shapes for the widening-row kernels on the CPU interpreter of the kernel sources: odd sizes,
one-pixel images, more taps than pixels, empty particle sets, channel counts around the register window / tile edges
(SURVEY.md section 4, item 3).  CPU only -- the GPU runs the fixed-shape versions of these tests."""
import numpy as np
import torch
from hypothesis import HealthCheck, example, given, settings, strategies as st

from lnst import _lib, ops
from oracle import transform as T
from test_widen_graphnet_kernels import ref_conv, ref_maxpool, ref_lrn

CFG = dict(max_examples=25, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


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


# ---- the hot-path operators under random shapes (odd sizes, N = 0, particles on faces / outside, nsize 1..4) ----------
from oracle import render as R  # noqa: E402


@settings(**CFG)
@given(res=st.lists(st.integers(1, 7), min_size=3, max_size=3), n=st.integers(0, 60), nsize=st.sampled_from([1, 2, 4]),
       clip=st.booleans(), seed=st.integers(0, 1000))
def test_splat_sph_any_shape(emu_lib, res, n, nsize, clip, seed):
    prev = _use_emu(emu_lib)
    try:
        rng = np.random.RandomState(seed)
        cell = 0.1
        domain = [r * cell for r in res]
        p = rng.uniform(-0.1, 1.1, (n, 3)).astype(np.float32)
        if n > 3:
            p[0], p[1], p[2] = 0.0, 1.0, -1.0                     # faces and the drivers' padding row
        p = torch.tensor(p)
        grid = _lib.make_grid(3, res, domain, nsize, clip)
        radius, support, rho = 0.025, 4, 1000.0
        scale = 0.8 * (2 * radius) ** 3
        out = ops.splat_sph_fwd(p, None, grid, radius * support, scale)
        pv = p.clone().requires_grad_(True)
        want = T.p2g(pv[None], domain, res, radius, rho, nsize, is_2d=False, clip=clip, support=support)[0, ..., 0] / rho
        _close(out, want, 2e-5)
        if n:
            g = torch.tensor(rng.randn(*res).astype(np.float32))
            (want * g).sum().backward()
            gp = ops.splat_sph_bwd_pos(p, None, grid, radius * support, scale, g)
            _close(gp, pv.grad, 5e-5)
    finally:
        _lib.set_for_testing(prev)


@settings(**CFG)
@given(res=st.lists(st.integers(1, 8), min_size=3, max_size=3), k=st.sampled_from([1, 3, 5]), seed=st.integers(0, 1000))
def test_smooth_and_raymarch_any_shape(emu_lib, res, k, seed):
    prev = _use_emu(emu_lib)
    try:
        rng = np.random.RandomState(seed)
        D, H, W = res
        d = torch.tensor((rng.rand(D, H, W) - 0.3).astype(np.float32), requires_grad=True)
        want = R.field_post(d[None, ..., None], k)[0, ..., 0]
        out = ops.smooth3_relu_fwd(d.detach(), torch.empty(D, H, W), k)
        _close(out, want, 2e-5)
        vol = torch.relu(d.detach())
        mats = [np.identity(3), np.matmul(T.rot_y_3d(17.0), T.rot_z_3d(-9.0))]
        rot = torch.tensor(np.asarray(mats), dtype=torch.float32).reshape(-1, 9)
        img, stot = torch.empty(2, H, W), torch.empty(2, H, W)
        ops.raymarch_fwd(vol, rot, 0.3, False, img, stot)
        dr = T.rotate(vol[None, ..., None], mats)
        cs = torch.flip(torch.cumsum(torch.flip(dr, [1]), 1), [1])
        _close(img, (dr * torch.exp(-cs * 0.3)).sum(1)[..., 0], 3e-5)
    finally:
        _lib.set_for_testing(prev)


def test_empty_particle_set_runs_like_the_oracle(emu_lib):
    """N = 0 (a frame before the first particles are seeded): every particle-sized buffer is empty, the C-ABI takes
    NULL for them, and the loop still renders / evaluates the (constant) loss like the oracle (liquid render: the
    smoke render's 0/0 normalisation is NaN in the reference as well)."""
    import warnings
    from helpers import liquid_cfg
    from lnst import synth
    from lnst.styler_3p import Styler
    from oracle.styler import Oracle3P
    import oracle.vgg
    prev = _use_emu(emu_lib)
    try:
        kw = dict(res=8, iter=2, conv_math='fp32', style_layer=['conv1_2'], w_style_layer=[1.0])
        params = {'p': [np.zeros((0, 3), np.float32)]}
        sty = synth.style_image(8, 8)
        st = Styler(liquid_cfg(**kw), weights=synth.vgg_weights(), device=torch.device('cpu'))
        st.style_img = sty
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            out = st.run(params)
            ref = Oracle3P(liquid_cfg(**kw), oracle.vgg.synthetic_weights()).run(params, style_targets=[sty])
        np.testing.assert_allclose(out['l'][0], ref['l'][0], rtol=2e-4)
        assert out['p'][0].shape == (0, 3) and float(np.abs(out['d']).max()) == 0.0
    finally:
        _lib.set_for_testing(prev)


@settings(**CFG)
@given(res=st.lists(st.integers(1, 6), min_size=3, max_size=3), n=st.integers(0, 50), nk=st.integers(1, 3),
       nsize=st.sampled_from([1, 2]), seed=st.integers(0, 1000))
def test_splat_wavg_any_shape(emu_lib, res, n, nk, nsize, seed):
    prev = _use_emu(emu_lib)
    try:
        rng = np.random.RandomState(seed)
        p = torch.tensor(rng.uniform(-0.05, 1.05, (n, 3)).astype(np.float32))
        r = torch.tensor(rng.uniform(0.2, 1.0, (n, nk)).astype(np.float32))
        var = torch.tensor(rng.uniform(-1.3, 1.3, (n, nk)).astype(np.float32))       # some outside the clip range
        grid = _lib.make_grid(3, res, res, nsize, False)
        hs = [0.5 * 4 / 2 ** k for k in range(nk)]
        cells = int(np.prod(res))
        wmap = ops.splat_wavg_wmap(p, grid, hs)
        out = ops.splat_wavg_fwd(p, r, var, grid, hs, wmap, torch.empty(nk, cells), torch.empty(*res))
        vv = var.clone().requires_grad_(True)
        x = r + torch.clamp(vv, -1, 1)
        want = sum(T.p2g_wavg(p[None], x[None, :, k:k + 1], res, res, 0.5, nsize, is_2d=False, clip=False,
                              support=4 / 2 ** k)[0, ..., 0] for k in range(nk)) if n else torch.zeros(*res)
        _close(out, want, 2e-5)
        if n:
            g = torch.tensor(rng.randn(*res).astype(np.float32))
            (want * g).sum().backward()
            gv = ops.splat_wavg_bwd(p, var, grid, hs, wmap, g, torch.empty(n, nk))
            nan_w, nan_g = torch.isnan(vv.grad), torch.isnan(gv)
            assert torch.equal(nan_w, nan_g)                                         # TF's where/div NaN rule
            ok = ~nan_w
            if ok.any():
                assert float((gv[ok] - vv.grad[ok]).abs().max()) <= 5e-5 * max(float(vv.grad[ok].abs().max()), 1e-30)
    finally:
        _lib.set_for_testing(prev)


@settings(**CFG)
@given(n=st.integers(0, 70), width=st.integers(1, 3), T_=st.integers(1, 6), sigma=st.floats(0.3, 3.0), seed=st.integers(0, 1000))
def test_adam_and_loop_glue_any_shape(emu_lib, n, width, T_, sigma, seed):
    from scipy.ndimage import gaussian_filter
    from oracle.adam import TFAdam
    prev = _use_emu(emu_lib)
    try:
        rng = np.random.RandomState(seed)
        var = torch.tensor(rng.randn(n, width).astype(np.float32))
        g = torch.tensor(rng.randn(n, width).astype(np.float32))
        m, v = torch.zeros_like(var), torch.zeros_like(var)
        state = torch.tensor([0.9, 0.999, 0.0])
        ref = TFAdam()
        want = var.clone()
        v0 = float(var.abs().max()) if n else 0.0
        for _ in range(2):
            ops.adam_step_dev(var, g, m, v, state, 0.05)
            want = ref.step(want, g, 0.05)
        if n:                                                   # fp32 round-off of the operands (|var0|, 2 lr), not of the result
            assert float((var - want).abs().max()) <= 1e-6 * (v0 + 0.1)
        assert abs(float(state[0]) - 0.9 ** 3) < 1e-6                                # the step counter ticks for n = 0 too
        x = torch.tensor(rng.randn(T_, max(n, 1), width).astype(np.float32))
        y = ops.temporal_gauss(x, sigma)
        _close(y, torch.tensor(gaussian_filter(x.numpy(), sigma=(sigma, 0, 0))), 3e-6)
    finally:
        _lib.set_for_testing(prev)


@settings(**CFG)
@given(H=st.integers(1, 9), W=st.integers(1, 9), oh=st.integers(1, 12), ow=st.integers(1, 12), nv=st.integers(1, 3),
       seed=st.integers(0, 1000))
def test_image_glue_any_shape(emu_lib, H, W, oh, ow, nv, seed):
    prev = _use_emu(emu_lib)
    try:
        rng = np.random.RandomState(seed)
        img = torch.tensor(rng.rand(nv, H, W).astype(np.float32) + 0.01)
        stats = ops.image_max(img, torch.empty(2 * nv))
        gray = ops.normalize_fwd(img, stats, torch.empty_like(img))
        _close(gray, img / img.reshape(nv, -1).max(1).values.reshape(nv, 1, 1), 2e-6)
        x = torch.tensor(rng.rand(nv, H, W, 1).astype(np.float32), requires_grad=True)
        want = R.resize_bilinear_legacy(x, oh, ow)
        _close(ops.resize_bilinear_fwd(x.detach(), oh, ow), want, 3e-6)
        g = torch.tensor(rng.randn(*want.shape).astype(np.float32))
        (want * g).sum().backward()
        _close(ops.resize_bilinear_bwd(g, H, W), x.grad, 3e-6)
    finally:
        _lib.set_for_testing(prev)


from oracle import loss as L  # noqa: E402


@settings(**CFG)
@given(H=st.integers(1, 9), W=st.integers(1, 9), cin=st.integers(1, 70), cout=st.integers(1, 70), n=st.integers(1, 2),
       seed=st.integers(0, 1000))
def test_fp32_loss_net_layers_any_shape(emu_lib, H, W, cin, cout, n, seed):
    prev = _use_emu(emu_lib)
    try:
        rng = np.random.RandomState(seed)
        x = torch.tensor(rng.randn(n, H, W, cin).astype(np.float32), requires_grad=True)
        w = torch.tensor((rng.randn(3, 3, cin, cout) / np.sqrt(9 * cin)).astype(np.float32))
        b = torch.tensor(rng.randn(cout).astype(np.float32))
        y = ops.conv3x3_f32(x.detach(), w, b, relu=True)
        want = torch.relu(torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, padding=1))
        want = want.permute(0, 2, 3, 1)
        _close(y, want, 2e-5)
        g = torch.tensor(rng.randn(*want.shape).astype(np.float32))
        (want * g).sum().backward()
        gx = ops.conv3x3_f32((g * (want > 0)).detach().contiguous(), w.flip(0, 1).permute(0, 1, 3, 2).contiguous(), None,
                             relu=False)
        _close(gx, x.grad, 3e-5)
        if H >= 2 and W >= 2:
            xp = x.detach().clone().requires_grad_(True)
            pw = torch.nn.functional.avg_pool2d(xp.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
            _close(ops.avgpool2_fwd(xp.detach()), pw, 2e-6)
            gp = torch.tensor(rng.randn(*pw.shape).astype(np.float32))
            (pw * gp).sum().backward()
            _close(ops.avgpool2_bwd(gp, None, xp.shape), xp.grad, 2e-6)
    finally:
        _lib.set_for_testing(prev)


@settings(**CFG)
@example(h=1, w=3, ch=1, hs=6, channel=0, seed=107)
@given(h=st.integers(1, 8), w=st.integers(1, 8), ch=st.integers(1, 80), hs=st.integers(1, 6), channel=st.integers(0, 79),
       seed=st.integers(0, 1000))
def test_losses_any_shape(emu_lib, h, w, ch, hs, channel, seed):
    prev = _use_emu(emu_lib)
    try:
        rng = np.random.RandomState(seed)
        P = h * w
        channel = channel % ch
        F = torch.tensor(np.maximum(rng.randn(P, ch), 0).astype(np.float32), requires_grad=True)
        Fs = torch.tensor(np.maximum(rng.randn(hs * 3, ch), 0).astype(np.float32))
        Gs, loss = torch.empty(ch, ch), torch.zeros(1)
        ops.gram_diff(Fs, 2.0 * hs * 3 * ch, None, 0.0, Gs, None)
        G = torch.empty_like(Gs)
        ops.gram_diff(F.detach(), 2.0 * P * ch, Gs, 0.7, G, loss)
        want, _ = L.style_loss([F.reshape(1, h, w, ch)], [Fs.reshape(1, hs, 3, ch)], [0.7], 1)
        # the loss is the square of a (possibly almost cancelled) difference G - Gs: its round-off scales with |G|^2,
        # not with the loss itself (h=1, w=3, ch=1, hs=6, seed=107 gives a loss of 1e-8 from G ~ 0.1)
        floor = 0.7 * 32 * np.finfo(np.float32).eps * float(max(G.abs().max(), Gs.abs().max())) ** 2 * ch * ch
        assert abs(float(loss) - float(want)) <= 3e-5 * abs(float(want)) + floor, (float(loss), float(want), floor)
        want.backward()
        gF = torch.empty(P, ch)
        ops.gram_bwd(F.detach(), G, 0.7 * 4.0 / (2.0 * P * ch), 0.0, 0, gF)
        _close(gF, F.grad, 3e-5)
        Fc = F.detach().clone().requires_grad_(True)
        wc = L.content_loss(Fc.reshape(1, h, w, ch), channel) * 1.5
        wc.backward()
        loss.zero_()
        gC = torch.empty(P, ch)
        ops.content_loss(Fc.detach(), channel, 1.5, loss, gC, 0.0, 0)
        if channel and channel == ch - 1:                       # mean over the empty slice f[..., c+1:]: NaN in TF, oracle, kernel
            assert np.isnan(float(wc.detach())) and np.isnan(float(loss))
        else:
            assert abs(float(loss) - float(wc.detach())) <= 3e-6 * max(abs(float(wc.detach())), float(Fc.detach().abs().max()), 1e-30)
        _close(gC, Fc.grad, 3e-6)
        img = torch.tensor((rng.rand(1, h, w, 3) * 255).astype(np.float32), requires_grad=True)
        tv = L.tv_loss(img) * 0.01
        tv.backward()
        loss.zero_()
        gi = torch.empty(h, w, 3)
        ops.tv_loss(img.detach()[0], 0.01, loss, gi)
        assert abs(float(loss) - float(tv.detach())) <= 3e-6 * max(abs(float(tv.detach())), 1.0)
        _close(gi, img.grad[0], 3e-6) if float(img.grad.abs().max()) > 0 else None
    finally:
        _lib.set_for_testing(prev)


@settings(**CFG)
@given(res=st.lists(st.integers(1, 7), min_size=3, max_size=3), liquid=st.booleans(), big=st.booleans(), seed=st.integers(0, 1000))
def test_raymarch_backward_any_shape(emu_lib, res, liquid, big, seed):
    prev = _use_emu(emu_lib)
    try:
        rng = np.random.RandomState(seed)
        D, H, W = res
        vol = torch.tensor((rng.rand(D, H, W) * (rng.rand(D, H, W) > 0.3)).astype(np.float32), requires_grad=True)
        mats = [np.matmul(T.rot_y_3d(70.0 if big else 8.0), T.rot_z_3d(-55.0 if big else 4.0)), np.identity(3)]
        rot = torch.tensor(np.asarray(mats), dtype=torch.float32).reshape(-1, 9)
        img, stot = torch.empty(2, H, W), torch.empty(2, H, W)
        ops.raymarch_fwd(vol.detach(), rot, 0.2, liquid, img, stot)
        dr = T.rotate(vol[None, ..., None], mats)
        if liquid:
            want = (1.0 - torch.exp(-dr.sum(1) * 0.2))[..., 0]
        else:
            cs = torch.flip(torch.cumsum(torch.flip(dr, [1]), 1), [1])
            want = (dr * torch.exp(-cs * 0.2)).sum(1)[..., 0]
        _close(img, want, 3e-5)
        g = torch.tensor(rng.randn(2, H, W).astype(np.float32))
        g[0, 0, 0] = 0.0                                          # a ray without incoming gradient
        (want * g).sum().backward()
        gv = torch.zeros(D, H, W)
        ops.raymarch_bwd(vol.detach(), rot, 0.2, liquid, stot, g, gv)
        _close(gv, vol.grad, 5e-5)
    finally:
        _lib.set_for_testing(prev)


@settings(**CFG)
@given(res=st.lists(st.integers(1, 9), min_size=2, max_size=2), n=st.integers(0, 50), nsize=st.sampled_from([1, 2, 4]),
       seed=st.integers(0, 1000))
def test_colour_splat_2d_any_shape(emu_lib, res, n, nsize, seed):
    prev = _use_emu(emu_lib)
    try:
        rng = np.random.RandomState(seed)
        dom = [r * 0.1 for r in res]
        p = torch.tensor(rng.uniform(-0.05, 1.05, (n, 2)).astype(np.float32))
        pc = torch.tensor(rng.rand(n, 3).astype(np.float32), requires_grad=True)
        pd = torch.tensor((1000 * (1 + 0.02 * rng.randn(n, 1))).astype(np.float32))
        grid = _lib.make_grid(2, res, dom, nsize, False)
        h, scale = 0.025 * 4, 0.8 * (2 * 0.025) ** 2 * 1000.0
        out = ops.splat_sph_fwd(p, None, grid, h, scale, pc=pc.detach(), pd=pd)
        want = T.p2g(p[None], dom, res, 0.025, 1000.0, nsize, pc=pc[None], pd=pd[None], is_2d=True, clip=False)[0]
        _close(out, want, 2e-5)
        if n:
            g = torch.tensor(rng.randn(*res, 3).astype(np.float32))
            (want * g).sum().backward()
            _close(ops.splat_sph_bwd_color(p, grid, h, scale, pd, 3, 1000.0, g), pc.grad, 3e-5)
    finally:
        _lib.set_for_testing(prev)


@settings(**CFG)
@given(dims=st.lists(st.integers(1, 7), min_size=2, max_size=3), C=st.integers(1, 3), seed=st.integers(0, 1000))
def test_warps_any_shape(emu_lib, dims, C, seed):
    prev = _use_emu(emu_lib)
    try:
        rng = np.random.RandomState(seed)
        dim = len(dims)
        d = torch.tensor(rng.rand(*dims, C).astype(np.float32))
        vel = torch.tensor(rng.uniform(-0.6, 0.6, dims + [dim]).astype(np.float32))
        _close(ops.advect(d, vel), T.advect(d[None], vel[None], is_3d=dim == 3)[0], 3e-6)
        if dim == 3:
            vol = d[..., 0].contiguous()
            mats = [np.matmul(T.rot_y_3d(25.0), T.rot_z_3d(-40.0)), np.identity(3)]
            rot = torch.tensor(np.asarray(mats), dtype=torch.float32).reshape(-1, 9)
            _close(ops.rotate_fwd(vol, rot), T.rotate(vol[None, ..., None], mats)[..., 0], 3e-6)
    finally:
        _lib.set_for_testing(prev)


@settings(**CFG)
@given(nv=st.integers(1, 3), H=st.integers(1, 7), W=st.integers(1, 7), ties=st.integers(1, 4), oh=st.integers(1, 9),
       ow=st.integers(1, 9), seed=st.integers(0, 1000))
def test_normalise_ties_and_bicubic_any_shape(emu_lib, nv, H, W, ties, oh, ow, seed):
    prev = _use_emu(emu_lib)
    try:
        rng = np.random.RandomState(seed)
        img = rng.rand(nv, H, W).astype(np.float32) * 0.9
        for v in range(nv):                                       # `ties` pixels share the maximum: reduce_max splits its gradient
            idx = rng.choice(H * W, size=min(ties, H * W), replace=False)
            img[v].reshape(-1)[idx] = 0.95
        img = torch.tensor(img, requires_grad=True)
        stats = ops.image_max(img.detach(), torch.empty(2 * nv))
        gray = ops.normalize_fwd(img.detach(), stats, torch.empty(nv, H, W))
        want = torch.stack([img[v] / torch.amax(img[v]) for v in range(nv)])
        _close(gray, want, 2e-6)
        g = torch.tensor(rng.randn(nv, H, W).astype(np.float32))
        (want * g).sum().backward()
        gi = ops.normalize_bwd(img.detach(), stats, g, torch.empty(nv), torch.empty(nv, H, W))
        # relative to the cotangent: the gradient itself can cancel to exactly 0 (a one-pixel image is constant 1)
        assert float((gi - img.grad).abs().max()) <= 3e-5 * max(float(g.abs().max()) / 0.95, 1e-30)
        x = torch.tensor(rng.rand(1, H, W, 1).astype(np.float32))
        _close(ops.resize_bicubic_fwd(x, oh, ow), L.bicubic_legacy(x, oh, ow), 3e-6)
    finally:
        _lib.set_for_testing(prev)

