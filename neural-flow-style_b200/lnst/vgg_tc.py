"""Tensor-core back end of the loss network (``csrc/conv_tc.cu``): bf16 NHWC activations, tcgen05
implicit-GEMM convolutions for every layer with >= 64 input and output channels, dedicated
CUDA-core kernels for conv1_1 (K = 27) and its data gradient, tcgen05 Gram matrices (F^T F,
MN-major operands, split-K) and Gram gradients (F x G, same kernel as the convolution with one
tap and a per-image B matrix).  Activations and gradients stay bf16 end to end; Gram matrices
and losses are fp32.

``split=True`` (``conv_math='bf16x3'``): the same kernels on [hi | lo] bf16 halves of every value -- three K passes per
convolution, fp32-tolerance results (csrc/conv_tc.cu ConvShape; include/lnst_b200.h "bf16x3").  Activation tensors
then have 2C physical channels."""
import weakref

import torch

from . import ops


def _pack(w):
    """HWIO fp32 [3,3,Cin,Cout] -> [9, Cout, Cin] bf16 (tap-major, K-major rows)."""
    return w.permute(0, 1, 3, 2).reshape(9, w.shape[3], w.shape[2]).to(torch.bfloat16).contiguous()


def _hilo(t):
    """fp32 [..., K] -> bf16 [..., 2K] = [hi | lo] along the last axis"""
    t = t.to(torch.float32)
    hi = t.to(torch.bfloat16)
    lo = (t - hi.to(torch.float32)).to(torch.bfloat16)
    return torch.cat([hi, lo], -1).contiguous()


def _pack2(w):
    """HWIO fp32 -> [9, Cout, 2*Cin] bf16 = [Whi | Wlo]"""
    return _hilo(w.permute(0, 1, 3, 2).reshape(9, w.shape[3], w.shape[2]))


class TensorCoreConvs:
    fuse_pool = True                                     # conv + 2x2 average pool from one epilogue (split mode)
    gray_dot = None                                      # (image, dots): set by the styler for ONE backward pass
    gray_dot_done = False
    # a pool's backward written by the data-gradient convolution above it: bit-identical, but the epilogue then writes four
    # masked pixels per accumulator row and outlasts the tile's MMAs (+4..17 us per step at C3): opt-in
    fuse_unpool = False
    fuse_gram = True                                     # Gram-loss gradient accumulated by the data-gradient convolution above it

    def __init__(self, net, split=False):
        # the network owns this object: a weak back-reference keeps the pair out of a reference cycle, so the
        # packed weights (~0.3 GB for VGG-19) are released with the network instead of waiting for the cycle collector
        self._net = weakref.ref(net)
        self.split = bool(split)
        pack = _pack2 if self.split else _pack
        self.wp, self.wdp = {}, {}
        self.first_bwd_tc = True
        self.gray_w = None
        for name, w in net.w.items():
            if w.shape[2] % 64 == 0 and w.shape[3] % 64 == 0:
                self.wp[name] = pack(w)
                self.wdp[name] = pack(net.wd[name])
            elif tuple(w.shape[2:]) == (3, 64):
                # conv1_1's data gradient (64 -> 3) as a 64 -> 16 tensor-core convolution: rows 3..15 zero
                wd16f = torch.zeros(9, 16, 64, dtype=torch.float32, device=w.device)
                wd16f[:, :3] = net.wd[name].to(torch.float32).permute(0, 1, 3, 2).reshape(9, 3, 64)
                self.wd16 = _hilo(wd16f) if self.split else wd16f.to(torch.bfloat16).contiguous()
                # gray render replicated to RGB (styler_base.py:41-43): x_c = 255*g - mean_c is folded into the weights
                from .vgg import _R_MEAN, _G_MEAN, _B_MEAN
                mean = torch.tensor([_R_MEAN, _G_MEAN, _B_MEAN], dtype=torch.float32, device=w.device)
                w32 = w.to(torch.float32)
                wm = (w32 * mean.view(1, 1, 3, 1)).sum(2).reshape(9, 64).contiguous()
                self.gray_w = ((255.0 * w32.sum(2)).reshape(9, 64).contiguous(), wm,
                               (net.b[name].to(torch.float32) - wm.sum(0)).contiguous())
                wdg = torch.zeros(9, 16, 64, dtype=torch.float32, device=w.device)
                wdg[:, 0] = 255.0 * net.wd[name].to(torch.float32).permute(0, 1, 3, 2).reshape(9, 3, 64).sum(1)
                self.wd16_gray = _hilo(wdg) if self.split else wdg.to(torch.bfloat16).contiguous()
                self.wg_gray = wdg[:, 0].contiguous()              # fp32 [9,64]: the CUDA-core kernel's weights

    @property
    def net(self):
        return self._net()

    # ---- network ----------------------------------------------------------------------------------
    def forward(self, x, layers, gray=None):
        """x fp32 [n,H,W,3], or ``gray`` fp32 [n,H,W] in 0..1 for a gray render (x is then not read).
        Returns the activation store (bf16 tensors, fp32 copies on demand)."""
        acts = {}
        cur = x
        sp = self.split
        self._layers = list(layers)                     # what gram() may fuse into (the layer above a style layer)
        self._gram_pending, self._gram_done = {}, set()
        pooled = None                                   # output of a pool layer already written by the convolution before it
        for i, name in enumerate(layers):
            if name.startswith('conv'):
                if name in self.wp and cur.dtype == torch.bfloat16:
                    nxt = layers[i + 1] if i + 1 < len(layers) else ''
                    if sp and self.fuse_pool and nxt.startswith('pool') and cur.shape[1] >= 2 and cur.shape[2] >= 2:
                        cur, pooled = ops.conv3x3_pool_bf16x3_tc(cur, self.wp[name], self.net.b[name], relu=True)
                    else:
                        cur = (ops.conv3x3_bf16x3_tc if sp else ops.conv3x3_bf16_tc)(cur, self.wp[name], self.net.b[name],
                                                                                   relu=True)
                elif gray is not None and name == layers[0] and tuple(self.net.w[name].shape[2:]) == (3, 64):
                    cur = (ops.conv_first_fwd_gray_x3 if sp else ops.conv_first_fwd_gray)(gray, *self.gray_w)
                elif cur.dtype == torch.float32 and tuple(self.net.w[name].shape[2:]) == (3, 64):
                    cur = (ops.conv_first_fwd_x3 if sp else ops.conv_first_fwd)(cur, self.net.w[name], self.net.b[name])
                elif sp:
                    raise NotImplementedError("conv_math='bf16x3' needs channel counts that are multiples of 64")
                else:
                    cur = ops.conv3x3_mixed(cur, self.net.w[name], self.net.b[name], relu=True, out_bf16=True)
            elif pooled is not None:
                cur, pooled = pooled, None
            else:
                cur = (ops.avgpool2_bf16x3_fwd if sp else ops.avgpool2_bf16_fwd)(cur)
            acts[name] = cur
        return _Acts(acts, split=sp)

    def backward(self, x, acts, layers, add_loss_grad, loss_layers, gray=False):
        g = None                                   # bf16 gradient of the current end point
        unpooled = False                           # the pool below was already handled by the convolution above it
        for i in range(len(layers) - 1, -1, -1):
            name = layers[i]
            if name in loss_layers:
                g = add_loss_grad(name, g)
            if g is None:
                continue
            if unpooled:                           # `name` is that pool: g is already the gradient of its input
                unpooled = False
                continue
            prev = layers[i - 1] if i > 0 else None
            prev_act = acts.raw[prev] if prev is not None else None
            mask = prev_act if (prev is not None and prev.startswith('conv')) else None
            sp = self.split
            if name.startswith('conv'):
                if name in self.wdp and prev is not None and sp and mask is not None and prev in self._gram_pending:
                    # the Gram-loss gradient of `prev` joins this data gradient inside the kernel (same ReLU mask)
                    g = ops.conv3x3_gram_bf16x3_tc(g, self.wdp[name], prev_act, self._gram_pending.pop(prev))
                    self._gram_done.add(prev)
                elif name in self.wdp and prev is not None and sp and self._can_unpool(layers, i, acts, g, loss_layers):
                    # data gradient + the average pool's backward + the ReLU mask below it from one epilogue
                    g = ops.conv3x3_unpool_bf16x3_tc(g, self.wdp[name], acts.raw[layers[i - 2]])
                    unpooled = True
                elif name in self.wdp and prev is not None:
                    g = (ops.conv3x3_bf16x3_tc if sp else ops.conv3x3_bf16_tc)(g, self.wdp[name], None, relu=False, mask=mask)
                elif prev is None and gray and tuple(self.net.w[name].shape[2:]) == (3, 64):
                    if getattr(self, 'first_bwd_direct', False) and ops._tma_ok():
                        # TMA-staged patch + CUDA cores with un-rounded fp32 weights: shared-memory-bandwidth bound
                        # (0.090 ms at C3 against 0.064 for the N = 16 MMA form), so not the default
                        g = ops.conv_first_bwd_gray_direct(g, sp, self.wg_gray)
                    elif self.gray_dot is not None:
                        # the render's normalisation needs sum(g_gray * image) next: reduced by the same kernel
                        img, dots = self.gray_dot
                        self.gray_dot, self.gray_dot_done = None, True
                        g = ops.conv_first_bwd_gray_dot_tc(g, self.wd16_gray, sp, img, dots)
                    else:
                        g = (ops.conv_first_bwd_gray_x3_tc if sp else ops.conv_first_bwd_gray_tc)(g, self.wd16_gray)  # d loss / d gray
                elif prev is None and tuple(self.net.w[name].shape[2:]) == (3, 64):
                    if sp:
                        g = ops.conv_first_bwd_x3_tc(g, self.wd16)
                    else:
                        g = ops.conv_first_bwd_tc(g, self.wd16) if self.first_bwd_tc else ops.conv_first_bwd(g, self.net.wd[name])
                elif sp:
                    raise NotImplementedError("conv_math='bf16x3' needs channel counts that are multiples of 64")
                else:
                    g = ops.conv3x3_mixed(g, self.net.wd[name], None, relu=False, out_bf16=(prev is not None), mask=mask)
            else:
                g = (ops.avgpool2_bf16x3_bwd if sp else ops.avgpool2_bf16_bwd)(g, mask, prev_act.shape)
        return g

    def _can_unpool(self, layers, i, acts, g, loss_layers):
        """conv layer i sits on a 2x2 average pool whose input is a conv layer's (ReLU) output with even extents, the pool's
        output carries no loss term of its own, and the layer's weights stream through shared memory (the kernel has
        no room for the staging tile next to resident weights)."""
        if not self.fuse_unpool or i < 2 or not layers[i - 1].startswith('pool') or not layers[i - 2].startswith('conv'):
            return False
        if layers[i - 1] in loss_layers:
            return False
        fine = acts.raw[layers[i - 2]]
        if fine.dtype != torch.bfloat16 or fine.shape[1] != 2 * g.shape[1] or fine.shape[2] != 2 * g.shape[2]:
            return False
        cin, cout = g.shape[3] // 2, fine.shape[3] // 2
        bn = 128 if cout % 128 == 0 else 64
        resident = cout == bn and 9 * (2 * cin // 64) * bn * 128 + 3 * 23552 <= 232448 - 2048 - 1024 - 512 - 16384 - 128
        return not resident

    # ---- losses -------------------------------------------------------------------------------------
    def gram(self, acts, name, Gs, weight, loss):
        F = acts.raw[name]
        P, ch = F.shape[1] * F.shape[2], F.shape[3] // (2 if self.split else 1)
        if ch % 64:
            raise NotImplementedError('tensor-core Gram needs a channel count that is a multiple of 64')
        if self.split:
            # a conv layer above `name` will compute the data gradient that lands here: it can carry F x Gd as well
            # (Gd pre-scaled by the coefficient styler_base.add_loss_grad applies: weight * 4 / (2 P C))
            layers = getattr(self, '_layers', [])
            i = layers.index(name) if name in layers else -1
            above = layers[i + 1] if 0 <= i < len(layers) - 1 else ''
            if self.fuse_gram and ch % 128 == 0 and above in self.wdp and name.startswith('conv') and Gs is not None:
                G, Gd2s = ops.gram_diff_bf16x3_tc(F, 2.0 * P * ch, Gs, weight, loss, gd_scale=weight * 4.0 / (2.0 * P * ch))
                self._gram_pending[name] = Gd2s
                return G, Gd2s
            return ops.gram_diff_bf16x3_tc(F, 2.0 * P * ch, Gs, weight, loss)
        return ops.gram_diff_bf16_tc(F, 2.0 * P * ch, Gs, weight, loss)

    def gram_grad(self, acts, name, handle, coef, g, relu_mask):
        if self.split:
            if name in self._gram_done:                           # already inside g (conv3x3_gram_bf16x3_tc)
                self._gram_done.discard(name)
                return g
            if name in self._gram_pending:                        # no gradient came from above: Gd2 carries the coefficient
                return ops.gram_bwd_bf16x3_tc(acts.raw[name], self._gram_pending.pop(name), 1.0, g, relu_mask, g)
            return ops.gram_bwd_bf16x3_tc(acts.raw[name], handle[1], coef, g, relu_mask, g)
        return ops.gram_bwd_bf16_tc(acts.raw[name], handle[1], coef, g, relu_mask, g)

    def content(self, acts, name, channel, weight, loss, g, relu_mask, target=None, amp=1.0):
        f = acts[name]                                          # fp32 copy
        n, P, ch = f.shape[0], f.shape[1] * f.shape[2], f.shape[3]
        to_f32, to_bf = (ops.from_split, ops.to_split) if self.split else (ops.to_f32, ops.to_bf16)
        g32 = to_f32(g) if g is not None else torch.empty_like(f)
        beta = 1.0 if g is not None else 0.0
        for v in range(n):
            if target is not None:
                ops.content_mse(f[v], target, amp, weight, loss[v:v + 1], g32[v], beta, relu_mask)
            else:
                ops.content_loss(f[v].reshape(P, ch), channel, weight, loss[v:v + 1], g32[v].reshape(P, ch), beta,
                                 relu_mask)
        return to_bf(g32)


class _Acts:
    """Activation store: ``raw`` holds the bf16 tensors; ``acts[name]`` hands out an fp32 copy
    (made on first use, cached)."""

    def __init__(self, raw, split=False):
        self.raw, self._f32, self.split = raw, {}, split

    def __getitem__(self, name):
        if name not in self._f32:
            self._f32[name] = (ops.from_split if self.split else ops.to_f32)(self.raw[name])
        return self._f32[name]
