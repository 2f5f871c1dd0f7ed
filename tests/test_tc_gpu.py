"""Tensor-core (tcgen05 + TMA) loss-network path -- GPU only.

The convolution is checked against an fp32 PyTorch reference evaluated on the SAME
bf16-rounded operands, so the only differences are fp32 accumulation order and the final bf16
rounding of the output: tolerance 2^-7 * max|ref| (one bf16 ulp at the top of the range).
Loop-level: conv_math='bf16' against the fp32 oracle, loss rel <= 2e-2 (stated in DESIGN.md).
"""
import numpy as np
import pytest
import torch

from helpers import smoke_cfg
from lnst import _lib, ops, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def cuda_lib():
    prev = _lib._lib
    _lib.set_for_testing(None)
    lib = _lib.get()
    assert lib.has_tc, 'tensor-core entry points missing from the CUDA library'
    yield
    torch.cuda.synchronize()
    _lib.set_for_testing(prev)


def pack(w):
    return w.permute(0, 1, 3, 2).reshape(9, w.shape[3], w.shape[2]).to(torch.bfloat16).contiguous()


def ref_conv(x, w, b, relu):
    y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, padding=1)
    y = torch.relu(y) if relu else y
    return y.permute(0, 2, 3, 1)


@pytest.mark.parametrize('n,H,W,cin,cout', [
    (1, 8, 16, 64, 64),          # exactly one tile
    (2, 13, 9, 64, 128),         # ragged tile edges, batch
    (1, 50, 50, 128, 256),       # conv3_1 shape of the headline workload
    (3, 100, 100, 64, 128),      # conv2_1
    (1, 25, 25, 256, 512),       # deeper layer, long K loop
    (1, 200, 200, 64, 64),       # conv1_2
])
def test_conv3x3_tc_matches_fp32_reference(n, H, W, cin, cout):
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(n * 1000 + H + cin)
    x = torch.randn(n, H, W, cin, generator=g).to(torch.bfloat16)
    w = (torch.randn(3, 3, cin, cout, generator=g) / np.sqrt(9 * cin)).to(torch.bfloat16)
    b = torch.randn(cout, generator=g)
    want = ref_conv(x.float(), w.float(), b, True)
    y = ops.conv3x3_bf16_tc(x.to(dev), pack(w.float()).to(dev), b.to(dev), relu=True)
    err = (y.float().cpu() - want).abs().max().item()
    assert err <= 2 ** -7 * want.abs().max().item(), (err, want.abs().max().item())
    # data-gradient form: no bias, no ReLU, ReLU mask of the layer below
    mask = torch.randn(n, H, W, cout, generator=g).to(torch.bfloat16)
    want2 = ref_conv(x.float(), w.float(), None, False) * (mask.float() > 0)
    y2 = ops.conv3x3_bf16_tc(x.to(dev), pack(w.float()).to(dev), None, relu=False, mask=mask.to(dev))
    err2 = (y2.float().cpu() - want2).abs().max().item()
    assert err2 <= 2 ** -7 * want2.abs().max().item(), err2


@pytest.mark.parametrize('n,h,w,cin,cout', [(2, 37, 29, 64, 64), (1, 50, 50, 128, 256), (3, 16, 8, 64, 128),
                                            (1, 100, 100, 128, 128), (2, 25, 25, 256, 128)])
def test_halo_kernel_equals_per_tap_kernel(n, h, w, cin, cout):
    """3x3 convolution through one halo'd patch per tile (taps via descriptor offsets; weights resident or
    streamed depending on the layer) against the one-TMA-tile-per-tap kernel: same bf16 output."""
    from lnst import _lib
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(n * 1000 + h + cin)
    x = torch.randn(n, h, w, cin, generator=g).to(torch.bfloat16).to(dev)
    wp = (torch.randn(9, cout, cin, generator=g) / (3 * cin ** 0.5)).to(torch.bfloat16).to(dev)
    b = torch.randn(cout, generator=g).to(dev)
    mask = torch.randn(n, h, w, cout, generator=g).to(torch.bfloat16).to(dev)
    lib = _lib.get()
    outs = []
    for halo in (0, 2):                      # 2: the halo kernel for every layer (resident or streamed weights)
        lib.call('lnst_set_conv_halo', halo)
        try:
            outs.append((ops.conv3x3_bf16_tc(x, wp, b, relu=True).float().cpu(),
                         ops.conv3x3_bf16_tc(x, wp, None, relu=False, mask=mask).float().cpu()))
        finally:
            lib.call('lnst_set_conv_halo', 2)              # the default
    for a, c in zip(outs[0], outs[1]):
        assert (a - c).abs().max() <= 2 ** -7 * a.abs().max()
        assert (a - c).abs().mean() <= 1e-3 * a.abs().mean()


def test_edge_layers_pool_and_conversions():
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 11, 15, 3, generator=g) * 50
    w = torch.randn(3, 3, 3, 64, generator=g) / 5
    b = torch.randn(64, generator=g)
    y = ops.conv3x3_mixed(x.to(dev), w.to(dev), b.to(dev), relu=True, out_bf16=True)
    want = ref_conv(x, w, b, True)
    assert (y.float().cpu() - want).abs().max() <= 2 ** -7 * want.abs().max()
    gy = torch.randn(2, 11, 15, 64, generator=g).to(torch.bfloat16)
    wd = w.flip(0, 1).permute(0, 1, 3, 2).contiguous()
    gx = ops.conv3x3_mixed(gy.to(dev), wd.to(dev), None, relu=False, out_bf16=False)
    want = ref_conv(gy.float(), wd, None, False)
    assert gx.dtype == torch.float32 and (gx.cpu() - want).abs().max() <= 1e-4 * want.abs().max()
    # dedicated conv1_1 kernels
    y1 = ops.conv_first_fwd(x.to(dev), w.to(dev), b.to(dev))
    want = ref_conv(x, w, b, True)
    assert (y1.float().cpu() - want).abs().max() <= 2 ** -7 * want.abs().max()
    gx1 = ops.conv_first_bwd(gy.to(dev), wd.to(dev))
    want = ref_conv(gy.float(), wd, None, False)
    assert (gx1.cpu() - want).abs().max() <= 1e-4 * want.abs().max()
    # the same data gradient as a 64 -> 16 tensor-core convolution (halo'd-patch kernel, fp32 [.,3] epilogue)
    wd16 = torch.zeros(9, 16, 64, dtype=torch.bfloat16)
    wd16[:, :3] = wd.permute(0, 1, 3, 2).reshape(9, 3, 64).to(torch.bfloat16)
    gx2 = ops.conv_first_bwd_tc(gy.to(dev), wd16.to(dev))
    want16 = ref_conv(gy.float(), wd.to(torch.bfloat16).float(), None, False)
    assert (gx2.cpu() - want16).abs().max() <= 1e-3 * want16.abs().max()
    a = torch.randn(2, 9, 12, 64, generator=g).to(torch.bfloat16)
    p = ops.avgpool2_bf16_fwd(a.to(dev))
    wantp = torch.nn.functional.avg_pool2d(a.float().permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    assert (p.float().cpu() - wantp).abs().max() <= 2 ** -7 * wantp.abs().max()
    gp = torch.randn(2, 4, 6, 64, generator=g).to(torch.bfloat16)
    back = ops.avgpool2_bf16_bwd(gp.to(dev), a.to(dev), a.shape)
    wantb = torch.zeros(2, 9, 12, 64)
    wantb[:, :8, :12] = gp.float().repeat_interleave(2, 1).repeat_interleave(2, 2) * 0.25
    wantb = wantb * (a.float() > 0)
    assert (back.float().cpu() - wantb).abs().max() <= 2 ** -7 * wantb.abs().max()
    z = torch.randn(1001, generator=g)
    assert torch.equal(ops.to_bf16(z.to(dev)).cpu(), z.to(torch.bfloat16))
    assert torch.equal(ops.to_f32(z.to(torch.bfloat16).to(dev)).cpu(), z.to(torch.bfloat16).float())


@pytest.mark.parametrize('n,h,w,ch', [(1, 8, 8, 64), (2, 50, 50, 256), (3, 100, 100, 128), (1, 13, 7, 512)])
def test_gram_tc_forward_and_gradient(n, h, w, ch):
    """F^T F on tensor cores (MN-major operands, split-K) and its gradient F x G."""
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(h * 7 + ch)
    F = torch.relu(torch.randn(n, h, w, ch, generator=g)).to(torch.bfloat16)
    Fs = torch.relu(torch.randn(1, h, w, ch, generator=g)).to(torch.bfloat16)
    P = h * w
    den = 2.0 * P * ch
    Gs, _ = ops.gram_diff_bf16_tc(Fs.to(dev), den, None, 0.0, None, want_bf16=False)
    want_s = torch.einsum('pc,pd->cd', Fs.float().reshape(P, ch), Fs.float().reshape(P, ch)) / den
    assert (Gs[0].cpu() - want_s).abs().max() <= 1e-4 * want_s.abs().max()
    loss = torch.zeros(n, device=dev)
    G, Gd = ops.gram_diff_bf16_tc(F.to(dev), den, Gs[0].contiguous(), 0.7, loss)
    Ff = F.float().reshape(n, P, ch)
    want = torch.einsum('npc,npd->ncd', Ff, Ff) / den - want_s
    assert (G.cpu() - want).abs().max() <= 2e-4 * (want.abs().max() + want_s.abs().max())
    want_loss = 0.7 * (want ** 2).sum(dim=(1, 2))
    np.testing.assert_allclose(loss.cpu().numpy(), want_loss.numpy(), rtol=2e-3)
    # gradient: (addend + coef * F G) * (F > 0), with G rounded to bf16 like the kernel's operand
    add = torch.randn(n, h, w, ch, generator=g).to(torch.bfloat16)
    coef = 0.37
    out = ops.gram_bwd_bf16_tc(F.to(dev), Gd, coef, add.to(dev), 1)
    ref = (add.float().reshape(n, P, ch) + coef * torch.einsum('npc,ncd->npd', Ff, Gd.float().cpu())) * (Ff > 0)
    err = (out.float().cpu().reshape(n, P, ch) - ref).abs().max().item()
    assert err <= 2 ** -7 * ref.abs().max().item(), err


def test_lossnet_bf16_against_fp32_features_and_gradient():
    """Whole prefix to conv3_1 on tensor cores vs the fp32 CUDA-core path on the same input."""
    from lnst.vgg import LossNet
    dev = torch.device('cuda:0')
    W = synth.vgg_weights()
    x = torch.tensor(synth.style_image(40, 56, seed=3)).reshape(1, 40, 56, 3) - 110.0
    x = x.to(dev).contiguous()
    wanted = ['conv2_1', 'conv3_1']
    n32, n16 = LossNet(W, 'vgg_19', dev, 'fp32'), LossNet(W, 'vgg_19', dev, 'bf16')
    a32, a16 = n32.forward(x, wanted), n16.forward(x, wanted)
    for l in wanted:
        rel = (a16[l] - a32[l]).norm() / a32[l].norm()
        assert rel < 2e-2, (l, rel.item())

    def top_grad(acts, as_bf16):
        def fn(name, g):
            if name != 'conv3_1':
                return g
            act = acts[name]                                        # fp32 view in both back ends
            t = (act > 0).float() * torch.sin(torch.arange(act.numel(), device=dev).reshape(act.shape) * 0.37)
            return t.to(torch.bfloat16) if as_bf16 else t
        return fn

    g32 = n32.backward(x, a32, wanted, top_grad(a32, False), {'conv3_1'})
    g16 = n16.backward(x, a16, wanted, top_grad(a16, True), {'conv3_1'})
    # bf16 perturbs activations by ~1e-2 relative; the ~1% of units whose pre-activation sits that
    # close to zero flip their ReLU mask, which alone is ~sqrt(0.01) = 10% L2 noise on a noise-like
    # top gradient (the gradient of a ReLU net is discontinuous); measured 0.09.
    rel = (g16 - g32).norm() / g32.norm()
    assert rel < 0.2, rel.item()


@pytest.mark.parametrize('view_mode', ['allreduce', 'sequential'])
def test_styler_bf16_vs_oracle(view_mode):
    from lnst.styler_3p import Styler
    from oracle.styler import Oracle3P
    import oracle.vgg
    res = 20
    kw = dict(res=res, iter=3, rotate=True, n_views=9, view_mode=view_mode, style_layer=['conv2_1', 'conv3_1'],
              w_style_layer=[0.5, 0.5])
    p, r = synth.smoke_particles(4000, 2, pad=4)
    sty = synth.style_image(res, res)
    new = Styler(smoke_cfg(conv_math='bf16', **kw), weights=synth.vgg_weights())
    new.style_img = sty
    out = new.run({'p': p, 'r': r})
    ref = Oracle3P(smoke_cfg(conv_math='fp32', **kw), oracle.vgg.synthetic_weights()).run(
        {'p': p, 'r': r}, style_targets=[sty], view_mode=view_mode)
    np.testing.assert_allclose(out['l'][0], ref['l'][0], rtol=2e-2)
    g_new, g_ref = out['g_opt'][0], ref['g_opt'][0].numpy()
    assert np.linalg.norm(g_new - g_ref) / np.linalg.norm(g_ref) < 0.15
    assert np.abs(out['d'] - ref['d']).max() <= 0.1 * np.abs(ref['d']).max()


@pytest.mark.parametrize('n,H,W', [(2, 13, 9), (3, 50, 64), (1, 200, 200)])
def test_gray_conv1_1_equals_the_rgb_form(n, H, W):
    """conv1_1 taken straight from a gray render (x255, RGB replication and mean subtraction folded into the
    weights, lnst_conv_first_fwd_gray / lnst_conv_first_bwd_gray_tc) against the explicit form: net input
    x = 255*g - mean built by lnst_to_net_input_fwd, fp32 convolution, and the gradient summed over channels."""
    from lnst import vgg
    dev = torch.device('cuda:0')
    gen = torch.Generator().manual_seed(H * 7 + W)
    gray = torch.rand(n, H, W, generator=gen)
    net = vgg.LossNet(synth.vgg_weights(), 'vgg_19', dev, math='bf16')
    assert net.gray_path()
    w, b = net.w['conv1_1'].float().cpu(), net.b['conv1_1'].float().cpu()
    mean = torch.tensor([vgg._R_MEAN, vgg._G_MEAN, vgg._B_MEAN])
    x = 255.0 * gray[..., None] - mean                                    # [n,H,W,3]
    want = ref_conv(x, w, b, True)
    y = ops.conv_first_fwd_gray(gray.to(dev), *net.tc.gray_w)
    err = (y.float().cpu() - want).abs().max().item()
    assert err <= 2 ** -7 * want.abs().max().item(), (err, want.abs().max().item())
    # data gradient w.r.t. the gray image: 255 * sum_c conv_transpose(g, w)_c
    g = torch.randn(n, H, W, 64, generator=gen).to(torch.bfloat16)
    gx = torch.nn.functional.conv_transpose2d(g.float().permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding=1)
    want_g = 255.0 * gx.sum(1)
    got_g = ops.conv_first_bwd_gray_tc(g.to(dev), net.tc.wd16_gray)
    err = (got_g.cpu() - want_g).abs().max().item()
    assert err <= 2e-2 * want_g.abs().max().item(), (err, want_g.abs().max().item())     # bf16 weights


def test_gray_path_runs_the_loop_like_the_rgb_path():
    """Styler.run with the gray fast path on and off (conv_math='bf16'): same optimisation within bf16 noise."""
    from lnst.styler_3p import Styler
    outs = []
    for gray in (True, False):
        cfg = smoke_cfg(res=20, iter=3, rotate=True, n_views=3, view_mode='allreduce', conv_math='bf16',
                        style_layer=['conv2_1'], w_style_layer=[1.0])
        p, r = synth.smoke_particles(3000, 2, pad=4)
        st = Styler(cfg, weights=synth.vgg_weights())
        st.gray_conv = gray
        st.style_img = synth.style_image(20, 20)
        assert st._gray_path() == gray
        outs.append(st.run({'p': p, 'r': r}))
    np.testing.assert_allclose(outs[0]['l'][0], outs[1]['l'][0], rtol=2e-2)
    assert np.abs(outs[0]['d'] - outs[1]['d']).max() <= 5e-2 * np.abs(outs[1]['d']).max()
