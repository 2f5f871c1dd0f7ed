"""Tensor-core back end of the loss network (``csrc/conv_tc.cu``): bf16 NHWC activations, tcgen05
implicit-GEMM convolutions for every layer with >= 64 input and output channels, CUDA-core
mixed-precision kernels for conv1_1 and its data gradient.  Loss kernels work on fp32 copies of the
(few) style/content end points."""
import torch

from . import ops


def _pack(w):
    """HWIO fp32 [3,3,Cin,Cout] -> [9, Cout, Cin] bf16 (tap-major, K-major rows)."""
    return w.permute(0, 1, 3, 2).reshape(9, w.shape[3], w.shape[2]).to(torch.bfloat16).contiguous()


class TensorCoreConvs:
    def __init__(self, net):
        self.net = net
        self.wp, self.wdp = {}, {}
        for name, w in net.w.items():
            if w.shape[2] % 64 == 0 and w.shape[3] % 64 == 0:
                self.wp[name] = _pack(w)
                self.wdp[name] = _pack(net.wd[name])

    def forward(self, x, layers):
        """x fp32 [n,H,W,3].  Returns {name: bf16 activation}; fp32 views are made on demand."""
        acts = {}
        cur = x
        for name in layers:
            if name.startswith('conv'):
                if name in self.wp and cur.dtype == torch.bfloat16:
                    cur = ops.conv3x3_bf16_tc(cur, self.wp[name], self.net.b[name], relu=True)
                elif cur.dtype == torch.float32 and tuple(self.net.w[name].shape[2:]) == (3, 64):
                    cur = ops.conv_first_fwd(cur, self.net.w[name], self.net.b[name])
                else:
                    cur = ops.conv3x3_mixed(cur, self.net.w[name], self.net.b[name], relu=True, out_bf16=True)
            else:
                cur = ops.avgpool2_bf16_fwd(cur)
            acts[name] = cur
        return _Acts(acts)

    def backward(self, x, acts, layers, add_loss_grad, loss_layers):
        g = None                                   # bf16 gradient of the current end point
        for i in range(len(layers) - 1, -1, -1):
            name = layers[i]
            if name in loss_layers:
                # loss terms live in fp32: convert, accumulate, convert back (style layers only)
                g32 = add_loss_grad(name, acts[name], ops.to_f32(g) if g is not None else None)
                if g32 is not None:
                    g = ops.to_bf16(g32)
            if g is None:
                continue
            prev = layers[i - 1] if i > 0 else None
            prev_act = acts.raw[prev] if prev is not None else None
            mask = prev_act if (prev is not None and prev.startswith('conv')) else None
            if name.startswith('conv'):
                if name in self.wdp and prev is not None:
                    g = ops.conv3x3_bf16_tc(g, self.wdp[name], None, relu=False, mask=mask)
                elif prev is None and tuple(self.net.w[name].shape[2:]) == (3, 64):
                    g = ops.conv_first_bwd(g, self.net.wd[name])
                else:
                    g = ops.conv3x3_mixed(g, self.net.wd[name], None, relu=False, out_bf16=(prev is not None), mask=mask)
            else:
                g = ops.avgpool2_bf16_bwd(g, mask, prev_act.shape)
        return g


class _Acts:
    """Activation store: ``raw`` holds the bf16 tensors; ``acts[name]`` hands the loss kernels an
    fp32 copy (made on first use, cached)."""

    def __init__(self, raw):
        self.raw, self._f32 = raw, {}

    def __getitem__(self, name):
        if name not in self._f32:
            self._f32[name] = ops.to_f32(self.raw[name])
        return self._f32[name]
