"""Loss network: slim-style VGG-16/19 with **average** pooling and post-ReLU end points
(reference ``vgg.py:68-120``), forward + data-gradient only (the weights are frozen).

Weights are in slim layout ``{'conv1_1': (w[3,3,Cin,Cout], b[Cout]), ...}``.  ``load_weights``
reads them from ``<model_path>`` (an ``.npz`` with ``vgg_19/conv1/conv1_1/weights`` style keys, or
plain ``conv1_1/weights``); the TF checkpoint itself cannot be parsed offline.

Two arithmetic back ends share this class:
  * ``fp32``  -- CUDA-core implicit-GEMM kernels (``csrc/lossnet.cu``), exact fp32 FMA;
  * ``bf16``  -- tcgen05 tensor-core kernels (``csrc/conv_tc.cu``), bf16 operands / fp32 accumulate.
"""
import os

import numpy as np
import torch

from . import ops

# vgg.py:16-18
_R_MEAN, _G_MEAN, _B_MEAN = 0.485 * 255, 0.456 * 255, 0.406 * 255

_BLOCKS = {
    'vgg_19': [(2, 64), (2, 128), (4, 256), (4, 512), (4, 512)],
    'vgg_16': [(2, 64), (2, 128), (3, 256), (3, 512), (3, 512)],
}


def layer_order(model='vgg_19'):
    """End points in network order: conv1_1, conv1_2, pool1, conv2_1, ..."""
    out = []
    for b, (rep, _) in enumerate(_BLOCKS[model], start=1):
        out += ['conv%d_%d' % (b, i) for i in range(1, rep + 1)] + ['pool%d' % b]
    return out


def model_name(network):
    return 'vgg_16' if '16' in os.path.basename(str(network)) else 'vgg_19'


# torchvision's ImageNet normalisation (its VGG weights expect (x/255 - mean)/std; the reference feeds
# x - 255*mean with no std, vgg.py:18-20,50-53)
_TV_STD = (0.229, 0.224, 0.225)


def from_torchvision(state_dict, model='vgg_19'):
    """Slim-layout weights from a torchvision ``vgg16`` / ``vgg19`` state dict
    (``features.<i>.weight`` [Cout,Cin,3,3], ``features.<i>.bias``; batch-norm variants are not
    supported).  OIHW -> HWIO; the 1/(255*std_c) input scaling torchvision expects is folded into
    conv1_1 so that the engine's mean-only preprocessing (``vgg.py:50-53``) produces the same
    activations as torchvision's own pipeline (with the reference's average pooling)."""
    convs = sorted({int(k.split('.')[1]) for k in state_dict if k.startswith('features.') and k.endswith('.weight')})
    names = [n for n in layer_order(model) if n.startswith('conv')]
    if len(convs) != len(names):
        raise ValueError('state dict has %d conv layers, %s needs %d' % (len(convs), model, len(names)))
    out = {}
    for name, idx in zip(names, convs):
        w = torch.as_tensor(state_dict['features.%d.weight' % idx]).detach().to(torch.float32)
        b = torch.as_tensor(state_dict['features.%d.bias' % idx]).detach().to(torch.float32)
        if w.ndim != 4 or tuple(w.shape[2:]) != (3, 3):
            raise ValueError('%s: unexpected weight shape %s' % (name, tuple(w.shape)))
        w = w.permute(2, 3, 1, 0).contiguous()                # [3,3,Cin,Cout]
        if name == 'conv1_1':
            w = w / (255.0 * torch.tensor(_TV_STD).reshape(1, 1, 3, 1))
        out[name] = (w, b.contiguous())
    return out


def load_weights(path, model='vgg_19'):
    """Slim-layout weights for ``<model_path>`` (``config.network``, e.g. ``data/model/vgg_19.ckpt``), from
      * the TensorFlow V1 checkpoint itself (``lnst.tfckpt``: the slim model-zoo file the reference loads,
        ``vgg.py:115-120``), or -- same stem --
      * ``.npz``: an export of it (keys ``vgg_19/conv1/conv1_1/weights`` or ``conv1_1/weights``),
      * ``.pth`` / ``.pt``: a torchvision state dict (see ``from_torchvision``)."""
    stem = os.path.splitext(path)[0]
    found = None
    for cand in ([path] if os.path.isfile(path) else []) + [stem + e for e in ('.npz', '.pth', '.pt')]:
        if os.path.isfile(cand):
            found = cand
            break
    if found is None:
        raise FileNotFoundError(
            'loss-network weights not found: %s (TF V1 checkpoint) or %s(.npz|.pth|.pt).  Download the slim vgg_19 '
            'checkpoint, export it to .npz with keys "vgg_19/conv1/conv1_1/weights", save a torchvision state dict, '
            'or pass weights= to Styler.' % (path, stem))
    if found.endswith(('.pth', '.pt')):
        sd = torch.load(found, map_location='cpu', weights_only=True)
        return from_torchvision(sd.get('state_dict', sd) if isinstance(sd, dict) else sd, model)
    if found.endswith('.npz'):
        blob = np.load(found)
    else:
        from . import tfckpt
        blob = tfckpt.read(found, names=lambda n: '/conv' in n and n.endswith(('/weights', '/biases')))
    out = {}
    for name in layer_order(model):
        if not name.startswith('conv'):
            continue
        cands = ['%s/%s/%s' % (model, name.split('_')[0], name), name]
        for c in cands:
            if c + '/weights' in blob:
                out[name] = (torch.tensor(np.asarray(blob[c + '/weights']), dtype=torch.float32),
                             torch.tensor(np.asarray(blob[c + '/biases']), dtype=torch.float32))
                break
        else:
            raise KeyError('no weights for %s in %s' % (name, found))
    return out


class LossNet:
    def gray_path(self):
        """True when conv1_1 can run straight from a gray render (tensor-core back end, ``vgg_tc``)."""
        return getattr(self, 'tc', None) is not None and self.tc.gray_w is not None

    """Forward/backward through the prefix of the network that the requested end points need."""

    def __init__(self, weights, model, device, math='fp32'):
        self.model, self.device, self.math = model, device, math
        self.order = layer_order(model)
        self.w, self.b, self.wd = {}, {}, {}
        for name, (w, b) in weights.items():
            w = w.to(device=device, dtype=torch.float32).contiguous()
            self.w[name] = w
            self.b[name] = b.to(device=device, dtype=torch.float32).contiguous()
            # data-gradient weights: flip taps, swap in/out channels (HWIO with I=Cout, O=Cin)
            self.wd[name] = w.flip(0, 1).permute(0, 1, 3, 2).contiguous()
        if math in ('bf16', 'bf16x3'):
            from . import vgg_tc
            self.tc = vgg_tc.TensorCoreConvs(self, split=(math == 'bf16x3'))
        elif math != 'fp32':
            raise ValueError('conv_math must be fp32, bf16 or bf16x3')

    def prefix(self, wanted):
        wanted = [w for w in wanted if w != 'input']
        for w in wanted:
            if w not in self.order:
                raise KeyError('%s is not a %s end point' % (w, self.model))
        last = max([self.order.index(w) for w in wanted]) if wanted else -1
        return self.order[:last + 1]

    # ---- forward ---------------------------------------------------------------------------
    def forward(self, x, wanted, gray=None):
        """x [n,H,W,3] (mean-subtracted).  Returns {end point: activation [n,h,w,C] fp32}."""
        if self.math != 'fp32':
            return self.tc.forward(x, self.prefix(wanted), gray=gray)
        acts = {}
        cur = x
        for name in self.prefix(wanted):
            if name.startswith('conv'):
                cur = ops.conv3x3_f32(cur, self.w[name], self.b[name], relu=True)
            else:
                cur = ops.avgpool2_fwd(cur)
            acts[name] = cur
        return acts

    # ---- backward --------------------------------------------------------------------------
    def backward(self, x, acts, wanted, add_loss_grad, loss_layers, gray=False):
        """d loss / d x.  For every end point in ``loss_layers``, ``add_loss_grad(name, g)`` adds
        the loss terms that live there into ``g`` (None = nothing accumulated yet) through
        ``gram_grad`` / ``content`` below and returns the buffer.  Gradients held for conv end
        points are w.r.t. the PRE-activation (ReLU mask already applied).  ``g`` is in the back
        end's native type (fp32 or bf16)."""
        if self.math != 'fp32':
            return self.tc.backward(x, acts, self.prefix(wanted), add_loss_grad, loss_layers, gray=gray)
        layers = self.prefix(wanted)
        g = None
        for i in range(len(layers) - 1, -1, -1):
            name = layers[i]
            if name in loss_layers:
                g = add_loss_grad(name, g)
            if g is None:
                continue
            prev = layers[i - 1] if i > 0 else None
            prev_act = acts[prev] if prev is not None else None
            mask = prev_act if (prev is not None and prev.startswith('conv')) else None
            if name.startswith('conv'):
                g = ops.conv3x3_f32(g, self.wd[name], None, relu=False, mask=mask)
            else:
                g = ops.avgpool2_bwd(g, mask, prev_act.shape)
        return g

    # ---- losses on end points (styler_base.py:96-102,135-185) -------------------------------------
    def features_f32(self, acts, name):
        return acts[name]            # fp32 in both back ends (the tensor-core store converts on demand)

    def gram(self, acts, name, Gs, weight, loss, mask=None):
        """Per image v: G_v = F_v^T F_v/den_v - Gs; loss[v] += weight*sum(G_v^2), den_v = 2 h w C.  Gs None:
        no subtraction (style-target pass).  ``mask`` = (m [n,h,w] fp32, area [n] host floats): the masked
        variant of styler_base.py:165-169 -- F_v is the feature times m, den_v = 2 area_v C; m is a constant
        here (2-D colour mode: the density mask does not depend on the colours).  Returns a handle for
        ``gram_grad`` / ``gram_values``."""
        if self.math != 'fp32':
            if mask is not None:
                raise NotImplementedError("style_mask needs conv_math='fp32'")
            return self.tc.gram(acts, name, Gs, weight, loss)
        f = acts[name]
        n, P, ch = f.shape[0], f.shape[1] * f.shape[2], f.shape[3]
        out = {'G': [], 'den': [], 'Fm': [], 'mask': mask, 'weight': weight, 'Gs': Gs, 'tmp': {}}
        for v in range(n):
            G = torch.empty(ch, ch, dtype=torch.float32, device=self.device)
            Fv, den = f[v].reshape(P, ch), 2.0 * P * ch
            if mask is not None:
                Fv = ops.mul_bcast(f[v], mask[0][v]).reshape(P, ch)
                if torch.is_tensor(mask[1]):                   # area on the device (3-D style mask: it changes every step)
                    den = (mask[1][v:v + 1], 2.0 * ch)
                else:
                    den = 2.0 * float(mask[1][v]) * ch
            Gs_v = Gs[v] if isinstance(Gs, (list, tuple)) else Gs     # per-image target (style_mask_on_ref)
            ops.gram_diff(Fv, den, Gs_v, weight, G, loss[v:v + 1] if loss is not None else None)
            out['G'].append(G)
            out['den'].append(den)
            out['Fm'].append(Fv)
        return out

    def gram_values(self, handle):
        """fp32 [n,C,C] view of a ``gram`` handle."""
        return handle[0] if self.math != 'fp32' else torch.stack(handle['G'], 0)

    def gram_grad(self, acts, name, handle, coef, g, relu_mask):
        """g <- (g + coef_v * F G) [* (F > 0)] with coef_v = 4 weight / den_v (``coef`` is that value for the
        unmasked denominator; the fp32 handle carries its own); allocates g when None."""
        if self.math != 'fp32':
            return self.tc.gram_grad(acts, name, handle, coef, g, relu_mask)
        f = acts[name]
        n, P, ch = f.shape[0], f.shape[1] * f.shape[2], f.shape[3]
        beta = 1.0
        if g is None:
            g, beta = torch.empty_like(f), 0.0
        for v in range(n):
            den, Gv = handle['den'][v], handle['G'][v]
            if isinstance(den, tuple):                         # device denominator: the coefficient rides on G
                cv, Gv = 1.0, ops.scale_by_dev(Gv, 4.0 * handle['weight'], den)
            else:
                cv = 4.0 * handle['weight'] / den
            if handle['mask'] is None:
                ops.gram_bwd(f[v].reshape(P, ch), Gv, cv, beta, relu_mask, g[v].reshape(P, ch))
            else:
                tmp = torch.empty(P, ch, dtype=torch.float32, device=self.device)
                ops.gram_bwd(handle['Fm'][v], Gv, cv, 0.0, 0, tmp)
                handle['tmp'][v] = tmp                             # d loss / d (F m), reused by gram_mask_grad
                ops.masked_accumulate(tmp.reshape(f[v].shape), handle['mask'][0][v], f[v], relu_mask, g[v], beta)
        return g

    def gram_mask_grad(self, acts, name, handle, style_side=None):
        """d loss / d mask [n,h,w] of a masked ``gram`` handle, after ``gram_grad`` ran on it (the 3-D style mask is
        the render, styler_base.py:165-169): per pixel <d loss/d(F m), F>, plus the pixel-independent term through
        den = 2 area C: -4 w C sum(D o (D + Gs)) / den with D = G/den - Gs.

        ``style_side`` = (style feature [1,h,w,C], per-image handles of its masked Gram): ``style_mask_on_ref`` -- the
        target Gs = (Fs m)^T (Fs m)/den depends on the mask too: minus <4 w/den (Fs m) D, Fs> per pixel, and the two
        denominator terms combine to -4 w C sum(D o D) / den."""
        f = acts[name]
        n, P, ch = f.shape[0], f.shape[1] * f.shape[2], f.shape[3]
        dm = torch.empty(n, f.shape[1], f.shape[2], dtype=torch.float32, device=self.device)
        for v in range(n):
            D = handle['G'][v]
            Gs = handle['Gs'][v] if isinstance(handle['Gs'], (list, tuple)) else handle['Gs']
            if style_side is not None:
                s = (D * D).sum().reshape(1)
            else:
                s = (D * (D + Gs)).sum().reshape(1) if Gs is not None else (D * D).sum().reshape(1)
            den = handle['den'][v]
            if isinstance(den, tuple):
                ops.rowdot(handle['tmp'][v], f[v].reshape(P, ch), dm[v].reshape(P),
                           scalar=ops.scale_by_dev(s, -4.0 * handle['weight'] * ch, den), scale=1.0)
            else:
                ops.rowdot(handle['tmp'][v], f[v].reshape(P, ch), dm[v].reshape(P), scalar=s,
                           scale=-4.0 * handle['weight'] * ch / den)
            if style_side is not None:
                fs, hs = style_side
                tmp_s = torch.empty(P, ch, dtype=torch.float32, device=self.device)
                if isinstance(den, tuple):
                    ops.gram_bwd(hs[v]['Fm'][0], ops.scale_by_dev(D, -4.0 * handle['weight'], den), 1.0, 0.0, 0, tmp_s)
                else:
                    ops.gram_bwd(hs[v]['Fm'][0], D, -4.0 * handle['weight'] / den, 0.0, 0, tmp_s)
                ops.rowdot(tmp_s, fs[0].reshape(P, ch), dm[v].reshape(P), accumulate=True)
        return dm

    def content(self, acts, name, channel, weight, loss, g, relu_mask, target=None, amp=1.0):
        """Content loss on end point ``name``: channel activation (styler_base.py:143-148), or, with
        ``target`` (the content image's feature, fp32 [h,w,C]), mean((f - target*amp)^2) (:137-141)."""
        if self.math != 'fp32':
            return self.tc.content(acts, name, channel, weight, loss, g, relu_mask, target, amp)
        f = acts[name]
        n, P, ch = f.shape[0], f.shape[1] * f.shape[2], f.shape[3]
        beta = 1.0
        if g is None:
            g, beta = torch.empty_like(f), 0.0
        for v in range(n):
            if target is not None:
                ops.content_mse(f[v], target, amp, weight, loss[v:v + 1], g[v], beta, relu_mask)
            else:
                ops.content_loss(f[v].reshape(P, ch), channel, weight, loss[v:v + 1], g[v].reshape(P, ch), beta,
                                 relu_mask)
        return g
