"""fp32 operators of the GraphDef loss networks (csrc/graphnet.cu) against plain PyTorch fp32 references with TF's
conventions (SAME padding with the odd cell after, MaxPoolGrad to the first maximum, LRN without the 1/n).
Twice: CPU interpreter of the kernel sources here (``dev=emu``) and the B200 through the C-ABI (``-m gpu``)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from lnst import ops
from test_kernel_parity import close


def tf_pad(x_nchw, k, stride, value=0.0):
    """TF 'SAME' padding of an NCHW tensor (explicit, so torch's symmetric padding rule is not involved)."""
    H, W = x_nchw.shape[-2:]
    (OH, pt), (OW, pl) = ops.same_pad(H, k, stride), ops.same_pad(W, k, stride)
    pb = max((OH - 1) * stride + k - H, 0) - pt
    pr = max((OW - 1) * stride + k - W, 0) - pl
    return F.pad(x_nchw, (pl, pr, pt, pb), value=value)


def ref_conv(x, w, b, stride, padding, relu):
    xp = x.permute(0, 3, 1, 2)
    if padding == 'SAME':
        xp = tf_pad(xp, w.shape[0], stride)
    y = F.conv2d(xp, w.permute(3, 2, 0, 1), b, stride=stride)
    return (F.relu(y) if relu else y).permute(0, 2, 3, 1)


def ref_maxpool(x, k, stride):
    return F.max_pool2d(tf_pad(x.permute(0, 3, 1, 2), k, stride, value=float('-inf')), k, stride).permute(0, 2, 3, 1)


def ref_lrn(x, r, bias, alpha, beta):
    sq = F.pad(x * x, (r, r))
    s = sum(sq[..., i:i + x.shape[-1]] for i in range(2 * r + 1))
    return x * (bias + alpha * s) ** (-beta)


CONVS = [  # n, H, W, Cin, Cout, k, stride, padding  -- the inception5h shapes in small
    (2, 13, 11, 3, 8, 7, 2, 'SAME'), (1, 9, 10, 6, 5, 1, 1, 'SAME'), (2, 8, 7, 5, 9, 3, 1, 'SAME'),
    (1, 7, 9, 4, 6, 5, 1, 'SAME'), (1, 12, 12, 3, 4, 7, 1, 'SAME'), (1, 9, 8, 3, 4, 3, 2, 'VALID'),
    (1, 70, 5, 70, 66, 3, 1, 'SAME'),          # more than one 64-wide tile in M, N and K
]


@pytest.mark.parametrize('n,H,W,cin,cout,k,stride,padding', CONVS)
def test_conv2d_fwd_bwd(dev, n, H, W, cin, cout, k, stride, padding):
    rng = np.random.RandomState(k * 10 + stride)
    x = torch.tensor(rng.randn(n, H, W, cin).astype(np.float32), requires_grad=True)
    w = torch.tensor((rng.randn(k, k, cin, cout) / np.sqrt(k * k * cin)).astype(np.float32))
    b = torch.tensor(rng.randn(cout).astype(np.float32))
    want = ref_conv(x, w, b, stride, padding, relu=False)
    got = ops.conv2d_f32(x.detach().to(dev), w.to(dev), b.to(dev), stride, padding)
    close(got, want, tol=3e-6, what='conv2d')
    close(ops.conv2d_f32(x.detach().to(dev), w.to(dev), b.to(dev), stride, padding, relu=True), F.relu(want), tol=3e-6)
    g = torch.tensor(rng.randn(*want.shape).astype(np.float32))
    (want * g).sum().backward()
    gx = torch.full(x.shape, 7.0).to(dev)
    ops.conv2d_bwd_data_f32(g.to(dev), w.to(dev), x.shape, stride, padding, gx, accumulate=False)
    close(gx, x.grad, tol=3e-6, what='conv2d dgrad')
    ops.conv2d_bwd_data_f32(g.to(dev), w.to(dev), x.shape, stride, padding, gx, accumulate=True)
    close(gx, 2 * x.grad, tol=3e-6, what='conv2d dgrad accumulate')


@pytest.mark.parametrize('cin', [5, 3])                      # 3: the thin-input kernel (one thread per input pixel)
def test_conv2d_dgrad_with_relu_mask(dev, cin):
    rng = np.random.RandomState(3)
    x = torch.tensor(rng.randn(2, 7, 6, cin).astype(np.float32), requires_grad=True)
    w = torch.tensor(rng.randn(3, 3, cin, 4).astype(np.float32))
    b = torch.tensor(rng.randn(4).astype(np.float32))
    y = ref_conv(x, w, b, 1, 'SAME', relu=True)
    g = torch.tensor(rng.randn(*y.shape).astype(np.float32))
    (y * g).sum().backward()
    yk = ops.conv2d_f32(x.detach().to(dev), w.to(dev), b.to(dev), 1, 'SAME', relu=True)
    gx = torch.empty(x.shape).to(dev)
    ops.conv2d_bwd_data_f32(g.to(dev), w.to(dev), x.shape, 1, 'SAME', gx, accumulate=False, relu_y=yk)
    close(gx, x.grad, tol=3e-6, what='dgrad through conv+relu')


def test_conv2d_into_concat_slice(dev):
    rng = np.random.RandomState(0)
    x = torch.tensor(rng.randn(1, 6, 5, 4).astype(np.float32))
    w = torch.tensor(rng.randn(3, 3, 4, 5).astype(np.float32))
    out = torch.full((1, 6, 5, 12), -1.0).to(dev)
    ops.conv2d_f32(x.to(dev), w.to(dev), None, 1, 'SAME', out=out, ch_off=4)
    want = ref_conv(x, w, None, 1, 'SAME', False)
    close(out[..., 4:9].contiguous(), want, tol=3e-6)
    assert bool((out[..., :4] == -1).all()) and bool((out[..., 9:] == -1).all())
    # the data gradient reads its cotangent from the same slice
    g = torch.tensor(rng.randn(1, 6, 5, 12).astype(np.float32))
    xg = x.clone().requires_grad_(True)
    (ref_conv(xg, w, None, 1, 'SAME', False) * g[..., 4:9]).sum().backward()
    gx = torch.empty(x.shape).to(dev)
    ops.conv2d_bwd_data_f32(g.to(dev), w.to(dev), x.shape, 1, 'SAME', gx, accumulate=False, ch_off=4)
    close(gx, xg.grad, tol=3e-6)


@pytest.mark.parametrize('H,W,stride', [(9, 8, 2), (7, 7, 1), (12, 5, 2), (2, 3, 2)])
def test_maxpool_fwd_bwd(dev, H, W, stride):
    rng = np.random.RandomState(H)
    xv = rng.randn(2, H, W, 5).astype(np.float32)
    xv[xv < 0.3] = 0.0                                           # post-ReLU-like input: many exact ties at 0
    x = torch.tensor(xv, requires_grad=True)
    want = ref_maxpool(x, 3, stride)
    got = ops.maxpool_fwd(x.detach().to(dev), 3, stride)
    np.testing.assert_array_equal(got.cpu().numpy(), want.detach().numpy())
    g = torch.tensor(rng.randn(*want.shape).astype(np.float32))
    (want * g).sum().backward()                                  # torch also routes to the first maximum
    gx = torch.empty(x.shape).to(dev)
    ops.maxpool_bwd(g.to(dev), x.detach().to(dev), 3, stride, 'SAME', gx, accumulate=False)
    close(gx, x.grad, tol=2e-6, what='maxpool grad')
    ops.maxpool_bwd(g.to(dev), x.detach().to(dev), 3, stride, 'SAME', gx, accumulate=True)
    close(gx, 2 * x.grad, tol=2e-6, what='maxpool grad accumulate')


@pytest.mark.parametrize('C,r', [(7, 2), (64, 5), (3, 4)])
def test_lrn_fwd_bwd(dev, C, r):
    rng = np.random.RandomState(C)
    x = torch.tensor((rng.rand(2, 3, 4, C) * 60).astype(np.float32), requires_grad=True)
    bias, alpha, beta = 2.0, 1e-3, 0.75
    want = ref_lrn(x, r, bias, alpha, beta)
    got = ops.lrn_fwd(x.detach().to(dev), r, bias, alpha, beta)
    close(got, want, tol=3e-6, what='lrn')
    g = torch.tensor(rng.randn(*x.shape).astype(np.float32))
    (want * g).sum().backward()
    gx = torch.zeros(x.shape).to(dev)
    ops.lrn_bwd(g.to(dev), x.detach().to(dev), r, bias, alpha, beta, gx, accumulate=False)
    close(gx, x.grad, tol=1e-5, what='lrn grad')
    # same rule as torch's local_response_norm once its alpha/n convention is undone
    tl = F.local_response_norm(x.detach().permute(0, 3, 1, 2), 2 * r + 1, alpha * (2 * r + 1), beta, bias)
    close(got, tl.permute(0, 2, 3, 1), tol=3e-6, what='lrn vs torch')


def test_relu_and_copy_channels(dev):
    rng = np.random.RandomState(1)
    x = torch.tensor(rng.randn(3, 4, 5, 6).astype(np.float32))
    x[0, 0, 0, 0] = 0.0
    y = ops.relu_fwd(x.to(dev))
    np.testing.assert_array_equal(y.cpu().numpy(), np.maximum(x.numpy(), 0))
    g = torch.tensor(rng.randn(*x.shape).astype(np.float32))
    gx = torch.ones(x.shape).to(dev)
    ops.relu_bwd(g.to(dev), y, gx, accumulate=True)
    np.testing.assert_allclose(gx.cpu().numpy(), 1 + g.numpy() * (x.numpy() > 0), rtol=1e-6)
    cat = torch.zeros(3, 4, 5, 10).to(dev)
    ops.copy_channels(x.to(dev), 0, cat, 3, 6)
    np.testing.assert_array_equal(cat[..., 3:9].cpu().numpy(), x.numpy())
    back = torch.ones(3, 4, 5, 4).to(dev)
    ops.copy_channels(cat, 4, back, 0, 4, accumulate=True)       # slice gradient, accumulated
    np.testing.assert_array_equal(back.cpu().numpy(), 1 + x.numpy()[..., 1:5])


@pytest.mark.parametrize('H,W,k,stride,padding', [(7, 7, 7, 1, 'VALID'), (10, 15, 7, 1, 'VALID'), (6, 5, 3, 2, 'SAME'),
                                                   (5, 5, 2, 2, 'VALID'), (4, 7, 3, 1, 'SAME')])
def test_avgpool_fwd_bwd(dev, H, W, k, stride, padding):
    rng = np.random.RandomState(H * W)
    x = torch.tensor(rng.randn(2, H, W, 3).astype(np.float32), requires_grad=True)
    xp = x.permute(0, 3, 1, 2)
    if padding == 'SAME':                                        # TF counts only the cells inside the image
        ones = tf_pad(torch.ones(1, 1, H, W), k, stride)
        want = F.avg_pool2d(tf_pad(xp, k, stride), k, stride, divisor_override=1) / \
            F.avg_pool2d(ones, k, stride, divisor_override=1)
    else:
        want = F.avg_pool2d(xp, k, stride)
    want = want.permute(0, 2, 3, 1)
    got = ops.avgpool_fwd(x.detach().to(dev), k, stride, padding)
    close(got, want, tol=2e-6, what='avgpool')
    g = torch.tensor(rng.randn(*want.shape).astype(np.float32))
    (want * g).sum().backward()
    gx = torch.full(x.shape, 3.0).to(dev)
    ops.avgpool_bwd(g.to(dev), x.shape, k, stride, padding, gx, accumulate=False)
    close(gx, x.grad, tol=3e-6, what='avgpool grad')
    ops.avgpool_bwd(g.to(dev), x.shape, k, stride, padding, gx, accumulate=True)
    close(gx, 2 * x.grad, tol=3e-6, what='avgpool grad accumulate')
