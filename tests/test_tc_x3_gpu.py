"""bf16x3 ("split") tensor-core loss network -- GPU only.  Every value travels as [hi | lo] bf16 halves and every
contraction runs three passes (hi*hi, lo*hi, hi*lo) into one fp32 accumulator: results are compared with fp32 / fp64
references at fp32-level tolerances (products carry 16 mantissa bits: 2^-16 = 1.5e-5 relative per term, averaging
down over K).  Loop level: conv_math='bf16x3' holds the SAME tolerances as the exact fp32 CUDA-core path
(tests/test_styler_parity.py): loss rel 2e-4, field 2e-4 of its maximum, variables rel-L2 2e-3."""
import numpy as np
import pytest
import torch

from helpers import smoke_cfg
from lnst import _lib, ops, synth
from lnst.vgg_tc import _hilo, _pack2

pytestmark = pytest.mark.gpu
TOL = 3e-5


@pytest.fixture(autouse=True)
def cuda_lib():
    prev = _lib._lib
    _lib.set_for_testing(None)
    lib = _lib.get()
    assert lib.has_tc
    yield
    torch.cuda.synchronize()
    _lib.set_for_testing(prev)


def ref_conv(x, w, b, relu):
    y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, padding=1)
    y = torch.relu(y) if relu else y
    return y.permute(0, 2, 3, 1)


def test_split_round_trip_and_pool():
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 9, 12, 64, generator=g) * 37
    s = ops.to_split(x.to(dev))
    assert s.shape == (3, 9, 12, 128) and torch.equal(s.cpu(), _hilo(x))
    back = ops.from_split(s).cpu()
    assert (back - x).abs().max() <= 2 ** -16 * x.abs().max()
    p = ops.from_split(ops.avgpool2_bf16x3_fwd(s)).cpu()
    want = torch.nn.functional.avg_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    assert (p - want).abs().max() <= TOL * want.abs().max()
    gp = torch.randn(3, 4, 6, 64, generator=g)
    got = ops.from_split(ops.avgpool2_bf16x3_bwd(ops.to_split(gp.to(dev)), s, s.shape)).cpu()
    wantb = torch.zeros(3, 9, 12, 64)
    wantb[:, :8, :12] = gp.repeat_interleave(2, 1).repeat_interleave(2, 2) * 0.25
    wantb = wantb * (x > 0)
    assert (got - wantb).abs().max() <= TOL * wantb.abs().max()


@pytest.mark.parametrize('n,H,W,cin,cout', [
    (1, 8, 16, 64, 64), (2, 13, 9, 64, 128), (1, 50, 50, 128, 256), (3, 100, 100, 64, 128), (1, 25, 25, 256, 512),
    (1, 200, 200, 64, 64), (2, 50, 50, 256, 128)])
def test_conv3x3_x3_matches_fp64_reference(n, H, W, cin, cout):
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(n * 1000 + H + cin)
    x = torch.randn(n, H, W, cin, generator=g)
    w = torch.randn(3, 3, cin, cout, generator=g) / np.sqrt(9 * cin)
    b = torch.randn(cout, generator=g)
    want = ref_conv(x.double(), w.double(), b.double(), True)
    xs = ops.to_split(x.to(dev))
    y = ops.from_split(ops.conv3x3_bf16x3_tc(xs, _pack2(w).to(dev), b.to(dev), relu=True)).cpu()
    err = (y.double() - want).abs().max().item()
    assert err <= TOL * want.abs().max().item(), (err, want.abs().max().item())
    mask = torch.randn(n, H, W, cout, generator=g)
    want2 = ref_conv(x.double(), w.double(), None, False) * (mask.double() > 0)
    y2 = ops.from_split(ops.conv3x3_bf16x3_tc(xs, _pack2(w).to(dev), None, relu=False, mask=ops.to_split(mask.to(dev)))).cpu()
    err2 = (y2.double() - want2).abs().max().item()
    assert err2 <= TOL * want2.abs().max().item(), err2


@pytest.mark.parametrize('n,h,w,ch', [(1, 8, 8, 64), (2, 50, 50, 256), (3, 100, 100, 128), (1, 13, 7, 512)])
def test_gram_x3_forward_and_gradient(n, h, w, ch):
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(h * 7 + ch)
    F = torch.relu(torch.randn(n, h, w, ch, generator=g))
    Fs = torch.relu(torch.randn(1, h, w, ch, generator=g))
    P = h * w
    den = 2.0 * P * ch
    Gs, _ = ops.gram_diff_bf16x3_tc(ops.to_split(Fs.to(dev)), den, None, 0.0, None)
    Fsd = Fs.double().reshape(P, ch)
    want_s = Fsd.t() @ Fsd / den
    assert (Gs[0].cpu().double() - want_s).abs().max() <= TOL * want_s.abs().max()
    loss = torch.zeros(n, device=dev)
    Fsp = ops.to_split(F.to(dev))
    G, Gd2 = ops.gram_diff_bf16x3_tc(Fsp, den, Gs[0].contiguous(), 0.7, loss)
    Fd = F.double().reshape(n, P, ch)
    want = torch.einsum('npc,npd->ncd', Fd, Fd) / den - Gs[0].cpu().double()
    assert (G.cpu().double() - want).abs().max() <= TOL * (want.abs().max() + want_s.abs().max())
    np.testing.assert_allclose(loss.cpu().numpy(), (0.7 * (want ** 2).sum(dim=(1, 2))).numpy(), rtol=2e-4)
    assert torch.equal(Gd2.cpu(), _hilo(G.cpu()))
    add = torch.randn(n, h, w, ch, generator=g)
    coef = 0.37
    out = ops.from_split(ops.gram_bwd_bf16x3_tc(Fsp, Gd2, coef, ops.to_split(add.to(dev)), 1)).cpu()
    ref = (add.double().reshape(n, P, ch) + coef * torch.einsum('npc,ncd->npd', Fd, G.cpu().double())) * (Fd > 0)
    err = (out.double().reshape(n, P, ch) - ref).abs().max().item()
    assert err <= TOL * ref.abs().max().item(), err


def test_lossnet_x3_against_fp32_features_and_gradient():
    """Whole prefix to conv3_1 in bf16x3 against the exact fp32 CUDA-core path: features and data gradient."""
    from lnst.vgg import LossNet
    dev = torch.device('cuda:0')
    W = synth.vgg_weights()
    wanted = ['conv2_1', 'conv3_1']
    n32, n3 = LossNet(W, 'vgg_19', dev, 'fp32'), LossNet(W, 'vgg_19', dev, 'bf16x3')
    for gray in (False, True):
        if gray:
            gimg = torch.rand(2, 40, 56, generator=torch.Generator().manual_seed(3)).to(dev)
            mean = torch.tensor([0.485 * 255, 0.456 * 255, 0.406 * 255], device=dev)
            x = (255.0 * gimg[..., None] - mean).contiguous()
            a3 = n3.forward(None, wanted, gray=gimg)
        else:
            x = (torch.tensor(synth.style_image(40, 56, seed=3)).reshape(1, 40, 56, 3) - 110.0).to(dev).contiguous()
            a3 = n3.forward(x, wanted)
        a32 = n32.forward(x, wanted)
        for l in wanted:
            rel = (a3[l] - a32[l]).norm() / a32[l].norm()
            assert rel < 2e-5, (l, rel.item())

        def top(acts, split):
            def fn(name, g):
                if name != 'conv3_1':
                    return g
                act = acts[name]
                t = (act > 0).float() * torch.sin(torch.arange(act.numel(), device=dev).reshape(act.shape) * 0.37)
                return ops.to_split(t) if split else t
            return fn
        g32 = n32.backward(x, a32, wanted, top(a32, False), {'conv3_1'})
        g3 = n3.backward(x, a3, wanted, top(a3, True), {'conv3_1'}, gray=gray)
        if gray:
            g32 = 255.0 * g32.sum(-1)
        # a unit whose pre-activation is within 1e-5 (relative) of zero can still flip its ReLU mask
        rel = (g3 - g32).norm() / g32.norm()
        assert rel < 2e-3, rel.item()


@pytest.mark.parametrize('view_mode', ['allreduce', 'sequential'])
def test_styler_x3_holds_the_fp32_tolerance(view_mode):
    from lnst.styler_3p import Styler
    from oracle.styler import Oracle3P
    import oracle.vgg
    res = 20
    kw = dict(res=res, iter=3, rotate=True, n_views=9, view_mode=view_mode, style_layer=['conv2_1', 'conv3_1'],
              w_style_layer=[0.5, 0.5])
    p, r = synth.smoke_particles(4000, 2, pad=4)
    sty = synth.style_image(res, res)
    new = Styler(smoke_cfg(conv_math='bf16x3', **kw), weights=synth.vgg_weights())
    new.style_img = sty
    out = new.run({'p': p, 'r': r})
    ref = Oracle3P(smoke_cfg(conv_math='fp32', **kw), oracle.vgg.synthetic_weights()).run(
        {'p': p, 'r': r}, style_targets=[sty], view_mode=view_mode)
    np.testing.assert_allclose(out['l'][0], ref['l'][0], rtol=2e-4)
    g_new, g_ref = out['g_opt'][0], ref['g_opt'][0].numpy()
    assert np.linalg.norm(g_new - g_ref) / np.linalg.norm(g_ref) < 2e-3
    # 27 Adam steps in sequential mode (9 per iteration): measured 2.0e-4 of the field maximum there
    assert np.abs(out["d"] - ref["d"]).max() <= (4e-4 if view_mode == "sequential" else 2e-4) * np.abs(ref["d"]).max()


@pytest.mark.parametrize('n,H,W', [(2, 13, 9), (3, 50, 64), (1, 200, 200), (2, 16, 16), (1, 33, 47)])
@pytest.mark.parametrize('split', [False, True])
def test_conv_first_bwd_gray_one_gemm_per_patch(n, H, W, split):
    """conv1_1's data gradient w.r.t. the gray render: the kernel with the 9 taps as the N dimension of one GEMM per halo'd
    patch (default) against an fp64 transposed convolution of the same cotangent, and against the per-tap halo kernel."""
    from lnst import _lib, synth, vgg
    dev = torch.device('cuda:0')
    gen = torch.Generator().manual_seed(H * 7 + W)
    net = vgg.LossNet(synth.vgg_weights(), 'vgg_19', dev, math='bf16x3' if split else 'bf16')
    w = net.w['conv1_1'].double().cpu()
    g32 = torch.randn(n, H, W, 64, generator=gen)
    g_dev = ops.to_split(g32.to(dev)) if split else g32.to(torch.bfloat16).to(dev)
    g_val = ops.from_split(g_dev).cpu().double() if split else g_dev.float().cpu().double()
    gx = torch.nn.functional.conv_transpose2d(g_val.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding=1)
    want = 255.0 * gx.sum(1)
    fn = ops.conv_first_bwd_gray_x3_tc if split else ops.conv_first_bwd_gray_tc
    lib = _lib.get()
    got = fn(g_dev, net.tc.wd16_gray).cpu().double()
    lib.call('lnst_set_conv_first_col', 0)
    try:
        halo = fn(g_dev, net.tc.wd16_gray).cpu().double()
    finally:
        lib.call('lnst_set_conv_first_col', 1)
    scale = want.abs().max().item()
    tol = 5e-5 if split else 2e-2                      # split weights carry 16 mantissa bits, plain bf16 weights 8
    assert (got - want).abs().max().item() <= tol * scale, ((got - want).abs().max().item(), scale)
    assert (got - halo).abs().max().item() <= 2e-5 * scale, ((got - halo).abs().max().item(), scale)


@pytest.mark.parametrize('n,h,w,ch', [(2, 50, 50, 256), (3, 100, 100, 128), (1, 13, 7, 512), (2, 9, 9, 384)])
def test_gram_x3_tmem_summed_products_match_the_split_row_gram(n, h, w, ch):
    """lnst_gram_diff_bf16x3_tc: hi^T hi + hi^T lo + lo^T hi summed in one TMEM tile (default) against the 2C x 2C Gram of
    the split rows (which also carries lo^T lo, 2^-16 relative) and against fp64."""
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(h * 11 + ch)
    F = torch.relu(torch.randn(n, h, w, ch, generator=g))
    P, den = h * w, 2.0 * h * w * ch
    Fsp = ops.to_split(F.to(dev))
    Gs = torch.zeros(ch, ch, device=dev)
    lib = _lib.get()
    outs = []
    for mode in (1, 0):
        lib.call('lnst_set_gram_split3', mode)
        try:
            loss = torch.zeros(n, device=dev)
            G, Gd2 = ops.gram_diff_bf16x3_tc(Fsp, den, Gs, 0.5, loss)
            outs.append((G.cpu().double(), Gd2.cpu(), loss.cpu().double()))
        finally:
            lib.call('lnst_set_gram_split3', 1)
    Fd = F.double().reshape(n, P, ch)
    want = torch.einsum('npc,npd->ncd', Fd, Fd) / den
    scale = want.abs().max()
    assert (outs[0][0] - want).abs().max() <= TOL * scale
    assert (outs[0][0] - outs[1][0]).abs().max() <= TOL * scale
    assert (outs[0][0] - outs[0][0].transpose(1, 2)).abs().max() <= 2e-6 * scale   # symmetric up to the fp32 summation order
    np.testing.assert_allclose(outs[0][2].numpy(), outs[1][2].numpy(), rtol=1e-4)
    assert torch.equal(outs[0][1], _hilo(outs[0][0].float()))


@pytest.mark.parametrize('n,H,W,cin,cout', [(2, 20, 24, 64, 64), (1, 33, 47, 64, 64), (3, 50, 50, 128, 128), (1, 17, 9, 64, 128),
                                            (9, 200, 200, 64, 64)])
def test_conv_with_fused_pool_is_bit_identical_to_conv_then_pool(n, H, W, cin, cout):
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(H * 3 + cout)
    x = ops.to_split(torch.randn(n, H, W, cin, generator=g).to(dev))
    w = _pack2(torch.randn(3, 3, cin, cout, generator=g) * 0.05).to(dev)
    b = torch.randn(cout, generator=g).to(dev)
    y0 = ops.conv3x3_bf16x3_tc(x, w, b, relu=True)
    p0 = ops.avgpool2_bf16x3_fwd(y0)
    y1, p1 = ops.conv3x3_pool_bf16x3_tc(x, w, b, relu=True)
    assert torch.equal(y0, y1)
    assert p1.shape == p0.shape and torch.equal(p0, p1)


@pytest.mark.parametrize('n,H,W,cin,cout', [(2, 20, 24, 128, 128), (1, 33, 47, 256, 128), (3, 50, 50, 128, 256), (9, 100, 100, 128, 128)])
def test_data_gradient_with_fused_gram_gradient(n, H, W, cin, cout):
    """lnst_conv3x3_gram_bf16x3_tc = the data-gradient convolution followed by lnst_gram_bwd_bf16x3_tc on its output."""
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(H * 5 + cin)
    x = ops.to_split(torch.randn(n, H, W, cin, generator=g).to(dev))
    w = _pack2(torch.randn(3, 3, cin, cout, generator=g) * 0.05).to(dev)
    F32 = torch.relu(torch.randn(n, H, W, cout, generator=g))
    F = ops.to_split(F32.to(dev))
    Gs = (torch.randn(cout, cout, generator=g) * 0.01).to(dev)
    Gs = (Gs + Gs.t()).contiguous()
    den, weight = 2.0 * H * W * cout, 0.5
    coef = weight * 4.0 / den
    loss = torch.zeros(n, device=dev)
    G, Gd2 = ops.gram_diff_bf16x3_tc(F, den, Gs, weight, loss)
    _, Gd2s = ops.gram_diff_bf16x3_tc(F, den, Gs, weight, torch.zeros(n, device=dev), gd_scale=coef * 1e3)
    y0 = ops.conv3x3_bf16x3_tc(x, w, None, relu=False, mask=F)
    want = ops.from_split(ops.gram_bwd_bf16x3_tc(F, Gd2, coef * 1e3, y0, 1)).cpu().double()
    got = ops.from_split(ops.conv3x3_gram_bf16x3_tc(x, w, F, Gd2s)).cpu().double()
    # against fp64: mask * (conv + coef * F x G)
    xd = ops.from_split(x).cpu().double()
    wd = (w[..., :cin].double() + w[..., cin:].double()).cpu().reshape(3, 3, cout, cin).permute(0, 1, 3, 2)   # HWIO
    conv = torch.nn.functional.conv2d(xd.permute(0, 3, 1, 2), wd.permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
    Fd = ops.from_split(F).cpu().double()
    ref = (conv + coef * 1e3 * torch.einsum('nhwc,ncd->nhwd', Fd, G.cpu().double())) * (Fd > 0)
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() <= TOL * scale, ((got - ref).abs().max().item(), scale)
    assert (got - want).abs().max().item() <= TOL * scale, ((got - want).abs().max().item(), scale)


@pytest.mark.parametrize('n,H,W,cin,cout', [(2, 10, 12, 128, 64), (1, 25, 25, 256, 128), (3, 17, 9, 128, 128), (9, 100, 100, 128, 64)])
def test_data_gradient_with_fused_pool_backward_is_bit_identical(n, H, W, cin, cout):
    """lnst_conv3x3_unpool_bf16x3_tc = the data-gradient convolution, then the 2x2 average pool's backward under the ReLU
    mask of the layer below the pool."""
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(H * 3 + cin)
    x = ops.to_split(torch.randn(n, H, W, cin, generator=g).to(dev))
    w = _pack2(torch.randn(3, 3, cin, cout, generator=g) * 0.05).to(dev)
    fine = ops.to_split(torch.relu(torch.randn(n, 2 * H, 2 * W, cout, generator=g)).to(dev))
    y = ops.conv3x3_bf16x3_tc(x, w, None, relu=False)
    want = ops.avgpool2_bf16x3_bwd(y, fine, fine.shape)
    got = ops.conv3x3_unpool_bf16x3_tc(x, w, fine)
    assert got.shape == want.shape and torch.equal(got, want)


@pytest.mark.parametrize('n,H,W', [(2, 13, 9), (3, 50, 64), (1, 200, 200), (1, 33, 47), (2, 16, 8)])
def test_conv_first_fwd_gray_as_one_gemm_per_tile(n, H, W):
    """conv1_1 of a gray render in split form: the K = 64 MMA kernel (three bf16 pieces per gray value and weight, six
    products) against the CUDA-core kernel and an fp64 convolution of x = 255 g - mean."""
    from lnst import vgg
    dev = torch.device('cuda:0')
    gen = torch.Generator().manual_seed(H * 7 + W)
    gray = torch.rand(n, H, W, generator=gen)
    net = vgg.LossNet(synth.vgg_weights(), 'vgg_19', dev, math='bf16x3')
    w, b = net.w['conv1_1'].double().cpu(), net.b['conv1_1'].double().cpu()
    mean = torch.tensor([vgg._R_MEAN, vgg._G_MEAN, vgg._B_MEAN], dtype=torch.float64)
    x = 255.0 * gray.double()[..., None] - mean
    want = torch.relu(torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, padding=1)).permute(0, 2, 3, 1)
    lib = _lib.get()
    cuda_core = ops.from_split(ops.conv_first_fwd_gray_x3(gray.to(dev), *net.tc.gray_w)).cpu().double()
    lib.call('lnst_set_conv_first_mma', 1)
    try:
        got = ops.from_split(ops.conv_first_fwd_gray_x3(gray.to(dev), *net.tc.gray_w)).cpu().double()
    finally:
        lib.call('lnst_set_conv_first_mma', 0)
    scale = want.abs().max().item()
    # the split output carries 16 mantissa bits: 2^-17 relative per value, plus fp32 accumulation
    assert (cuda_core - want).abs().max().item() <= 1.2e-5 * scale
    assert (got - want).abs().max().item() <= 1.2e-5 * scale, ((got - want).abs().max().item(), scale)
    assert (got - cuda_core).abs().max().item() <= 1.2e-5 * scale
