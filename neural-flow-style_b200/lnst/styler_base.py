"""StylerBase: what the 2-D and 3-D stylers share -- configuration copy, loss network, style /
content targets and the feature-space losses (reference ``styler_base.py:11-345``).

The reference builds a TF graph and lets autodiff produce d(loss)/d(image); here the same
chain is an explicit sequence of kernel launches (``image_loss_and_grad``), all on one CUDA
stream with no host synchronisation.
"""
import os

import numpy as np
import torch
from PIL import Image

from . import _lib, ops
from .util import crop_ratio, resize
from .vgg import LossNet, load_weights, model_name

f32 = torch.float32


class StylerBase(object):
    def __init__(self, self_dict, weights=None, device=None, content_weights=None):
        # styler_base.py:14-15 -- every config attribute becomes an attribute of the styler
        for arg in vars(self_dict):
            setattr(self, arg, getattr(self_dict, arg))
        if not hasattr(self, 'view_mode'):
            self.view_mode = 'sequential'
        if not hasattr(self, 'conv_math'):
            self.conv_math = 'bf16x3'
        if getattr(self, 'w_hist', 0) > 0:
            raise NotImplementedError('histogram loss (w_hist > 0) is not built: DESIGN.md section 5 (out of scope, '
                                      'broken at the reference HEAD, styler_base.py:203-207)')
        # the masked Gram, v_batch groups and fed batches run on the exact fp32 loss-net path only: a reference
        # config that asks for them (test_dambreak2d.py:190 sets style_mask) gets that path instead of an error
        if self.conv_math != 'fp32' and 'vgg' in getattr(self, 'network', '') and (
                getattr(self, 'style_mask', False) or getattr(self, 'batch_size', 1) > 1 or getattr(self, 'v_batch', 1) > 1):
            import warnings
            warnings.warn("conv_math=%r -> 'fp32': style_mask / batch_size > 1 / v_batch > 1 run on the fp32 loss-network "
                          'path' % self.conv_math)
            self.conv_math = 'fp32'
        lib = _lib.get()                          # raises without the CUDA library / a GPU
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device()) if lib.kind == 'cuda' else torch.device('cpu')
        self.device = torch.device(device)
        self.model_path = os.path.join(self.data_dir, self.model_dir, self.network)   # :18
        if 'vgg' in self.model_path:
            if weights is None:
                weights = load_weights(self.model_path, model_name(self.network))
            self.net = LossNet(weights, model_name(self.network), self.device, math=self.conv_math)
        else:
            # inception5h: a frozen GraphDef read by tensor name (styler_base.py:17-31,53-57); ``weights`` may be
            # the parsed node list (lnst.graphdef.Node), else the .pb at model_path is read
            from . import graphdef
            from .graphnet import GraphNet
            nodes = weights if weights is not None else graphdef.load(self.model_path)
            self.net = GraphNet(nodes, self.device, pool1=bool(getattr(self, 'pool1', False)))
        # multi-net loss (engine extension, BASELINE.json configs[4]): the content loss lives on a second network
        self.net2 = None
        if getattr(self, 'content_network', ''):
            path2 = os.path.join(self.data_dir, self.model_dir, self.content_network)
            if 'vgg' in path2:
                w2 = content_weights if content_weights is not None else load_weights(path2, model_name(self.content_network))
                self.net2 = LossNet(w2, model_name(self.content_network), self.device, math='fp32')
            else:
                from . import graphdef
                from .graphnet import GraphNet
                nodes2 = content_weights if content_weights is not None else graphdef.load(path2)
                self.net2 = GraphNet(nodes2, self.device, pool1=bool(getattr(self, 'pool1', False)))
        self.content_img = None
        self.style_img = None
        self._content_feat = None              # set per octave by run() when a content target image is given
        self.rank, self.world = 0, 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.rank, self.world = torch.distributed.get_rank(), torch.distributed.get_world_size()

    # ---- targets (styler_base.py:311-345) ------------------------------------------------------
    def load_img(self, hw=None):
        self.content_img = None
        self.style_img = None
        if self.w_content > 0 and self.content_target:
            img = np.float32(Image.open(self.content_target))
            if img.shape[-1] == 4:
                img = img[..., :-1]
            if hw is not None:
                img = crop_ratio(img, hw[1] / hw[0])
            self.content_img = img
        if self.w_style > 0 and self.style_target:
            img = np.float32(Image.open(self.style_target))
            if self.style_tiling > 1:
                img = np.tile(img, (self.style_tiling, self.style_tiling, 1))
            if hw is not None:
                img = crop_ratio(img, hw[1] / hw[0])
            self.style_img = img

    # ---- semi-Lagrangian transport between frames (styler_base.py:59-74) ----------------------------------------
    def _transport(self, g, v, a, b, recursive=True):
        """Carry a per-cell field g [H,W,C] (or [D,H,W,C]) from frame a to frame b along the velocity fields
        v [N,H,W,dim] (normalised units per frame) by repeated order-1 advection -- the reference's temporal-coherence
        helper (its graph nodes ``self.adv / self.g / self.u`` are not built at HEAD, so nothing calls it there; kept
        with the reference's signature and semantics over ``lnst_advect``).  Returns a NumPy array like the reference."""
        dev = self.device
        gt = torch.as_tensor(np.ascontiguousarray(g, dtype=np.float32)).to(dev)
        vt = torch.as_tensor(np.ascontiguousarray(v, dtype=np.float32)).to(dev)
        if a < b:
            if recursive:
                for i in range(a, b):
                    gt = ops.advect(gt, vt[i].contiguous())
            else:                                                  # forward once
                gt = ops.advect(gt, (vt[a] * (b - a)).contiguous())
        elif a > b:
            if recursive:
                for i in reversed(range(b, a)):
                    gt = ops.advect(gt, (-vt[i]).contiguous())
            else:
                gt = ops.advect(gt, (-vt[a - 1] * (a - b)).contiguous())
        return gt.cpu().numpy()

    # ---- end points needed -----------------------------------------------------------------------
    def _wanted(self):
        w = []
        if self.w_style and self.style_img is not None:
            w += list(self.style_layer)
        if self.w_content and self.net2 is None:
            w.append(self.content_layer)
        return w

    def _net_hw(self, hw):
        """Size of the image fed to the net for a render of size hw (styler_base.py:35-37)."""
        if np.isclose(self.resize_scale, 1):
            return int(hw[0]), int(hw[1])
        s = np.float32(self.resize_scale)
        return int(np.float32(hw[0]) * s), int(np.float32(hw[1]) * s)

    def _target_tensor(self, img, hw):
        """Style / content target resampled to the net input size (styler_base.py:257-262)."""
        img = np.asarray(img, dtype=np.float32)
        if img.shape[-1] == 4:                                  # :252-255
            img = img[..., :-1] * (img[..., -1:] / 255)
        img = resize(img, self._net_hw(hw), order=3)
        x = torch.tensor(img, dtype=f32, device=self.device).reshape(1, img.shape[0], img.shape[1], 3)
        mean = torch.tensor([0.485 * 255, 0.456 * 255, 0.406 * 255], dtype=f32, device=self.device)
        return (x - mean).contiguous()

    def _style_feature(self, style_target, style_shp):
        """Gram matrices of the style target, one per style layer, already divided by
        2*h*w*C (styler_base.py:157-162,178-179, 249-278).  Returned as device tensors."""
        x = self._target_tensor(style_target, style_shp)
        acts = self.net.forward(x, list(self.style_layer))
        # kept for style_mask_on_ref (styler_base.py:171-173: the style feature itself is masked per frame)
        self._style_acts = {l: self.net.features_f32(acts, l) for l in self.style_layer} if getattr(
            self, 'style_mask_on_ref', False) and self.style_mask else None
        grams = []
        for l in self.style_layer:
            if 'input' in l:
                raise NotImplementedError("style layer 'input' is not built")
            handle = self.net.gram(acts, l, None, 0.0, None)
            grams.append(self.net.gram_values(handle)[0].contiguous())
        return grams

    @staticmethod
    def _feature_pixels(h, w, name):
        """h*w of end point ``name`` for an h x w input (2x2/2 VALID pooling between blocks)."""
        block = int(name[4]) if name.startswith('conv') else int(name[4]) + 1
        for _ in range(block - 1):
            h, w = h // 2, w // 2
        return h * w

    def masked_style_grams(self, masks, per_image=False):
        """Style Gram targets under ``style_mask_on_ref`` (styler_base.py:171-173): (Fs m)^T (Fs m) / (2 area C) with the
        render's own mask and area; ``masks`` as returned by ``style_masks_for``.  One image: a list of Gram matrices
        (per style layer).  ``per_image``: ({layer: [Gram per image]}, {layer: [handle per image]}) -- the 3-D styler,
        where every view has its own mask and the mask carries a gradient."""
        grams, by_layer, handles = [], {}, {}
        for l in self.style_layer:
            fs, (m, area) = self._style_acts[l], masks[l]
            if tuple(fs.shape[1:3]) != tuple(m.shape[1:3]):
                raise ValueError('style_mask_on_ref: style feature %s and render feature %s differ in size (the reference '
                                 'multiplies them elementwise)' % (tuple(fs.shape[1:3]), tuple(m.shape[1:3])))
            hs = [self.net.gram({l: fs}, l, None, 0.0, None, mask=(m[v:v + 1], area[v:v + 1]))
                  for v in range(m.shape[0] if per_image else 1)]
            by_layer[l] = [self.net.gram_values(h)[0].contiguous() for h in hs]
            handles[l] = hs
            grams.append(by_layer[l][0])
        return (by_layer, handles) if per_image else grams

    def _content_feature(self, content_target, content_shp):
        """Feature of the content target at ``content_layer`` (styler_base.py:233-247): fp32 [h,w,C] on the
        device; ``image_loss_and_grad`` compares every view's feature with it."""
        x = self._target_tensor(content_target, content_shp)
        net = self.net2 if self.net2 is not None else self.net
        acts = net.forward(x, [self.content_layer])
        feat = net.features_f32(acts, self.content_layer)[0].contiguous()
        if self.top_k > 0:                                         # :240-246: keep the k strongest logits of every row
            assert 'softmax2_pre_activation' in self.content_layer
            rows = feat.reshape(-1, feat.shape[-1])
            keep = torch.zeros_like(rows, dtype=torch.bool)
            keep.scatter_(1, torch.topk(rows.abs(), int(self.top_k), dim=1).indices, True)
            feat = torch.where(keep, rows, torch.zeros_like(rows)).reshape(feat.shape).contiguous()
        return feat

    # ---- feature-space losses + their gradient w.r.t. the net input ---------------------------------
    def style_masks_for(self, d_gray, net_hw, device_areas=False):
        """Per style layer: (m [n,h,w], area [n]) with m = d_gray resized to the layer's feature size by
        TF's legacy bicubic (styler_base.py:165-169).  The areas are read back once, here (2-D colour mode: d_gray does not
        depend on the optimised variable) -- or, ``device_areas``, stay a device tensor (3-D: the mask is the current
        render; the Gram kernels take their denominators from the device so that the step stays graph-capturable)."""
        out = {}
        for l in self.style_layer:
            h, w = net_hw
            block = int(l[4]) if l.startswith('conv') else int(l[4]) + 1
            for _ in range(block - 1):
                h, w = h // 2, w // 2
            m = ops.resize_bicubic_fwd(d_gray.reshape(d_gray.shape[0], d_gray.shape[1], d_gray.shape[2], 1).contiguous(),
                                       h, w)[..., 0].contiguous()
            area = m.sum(dim=(1, 2))
            out[l] = (m, area.contiguous() if device_areas else [float(a) for a in area.cpu().tolist()])
        return out

    def image_loss_and_grad(self, x, d_img, style_grams, loss, style_masks=None, gray=None, mask_grads=None, group=False,
                            share=None, style_side=None):
        """x [n,H,W,3] net input (one image per view), d_img the same before mean subtraction.
        Adds each image's total feature/TV loss into ``loss[v]`` and returns d loss_v / d x_v
        stacked [n,H,W,3] (styler_base.py:127-213).

        ``gray`` [n,H,W] (0..1): the render is one channel replicated to RGB and the loss net can start from it
        (``LossNet.gray_path``, no TV loss): x and d_img are not read (may be None) and the result is
        d loss_v / d gray_v [n,H,W].

        ``group``: the n images are one fed batch of the reference graph with batch_size = 1 (v_batch > 1): the Gram
        loss reads image 0 only (``_gram_matrix`` loops over range(batch_size), styler_base.py:98), the content and TV
        terms are means over the batch (:137-148, :212), i.e. weighted 1/n per image.

        ``share``: weight of the batch-mean terms (content, TV) when the images of one fed batch are evaluated in
        separate calls (batch_size > 1 in the 2-D styler: 1/batch_size); the Gram terms are sums over the batch.

        ``mask_grads`` (a dict, 3-D style mask): filled with {style layer: d loss / d mask [n,h,w]} -- there the mask
        is the render itself and carries a gradient (styler_base.py:165-169)."""
        n = x.shape[0] if gray is None else gray.shape[0]
        hw = (x.shape[1], x.shape[2]) if gray is None else (gray.shape[1], gray.shape[2])
        wanted = self._wanted()
        style_on = bool(self.w_style) and style_grams is not None
        acts = self.net.forward(x, wanted, gray=gray) if wanted else {}
        shapes = {}
        handles = {}
        if share is None:
            share = 1.0 / n if group else 1.0                      # per-image weight of the batch-mean terms
        acts_style = acts
        if group and style_on:                                     # image 0 only
            acts_style = {l: self.net.features_f32(acts, l)[:1] for l in self.style_layer}
        if style_on:
            for li, l in enumerate(self.style_layer):
                handles[l] = self.net.gram(acts_style, l, style_grams[li], self.w_style * self.w_style_layer[li], loss,
                                           mask=style_masks[l] if style_masks else None)

        def add_loss_grad(name, g):
            is_conv = 1 if (name.startswith('conv') and not hasattr(self.net, 'relu_masked')) else 0
            if style_on:
                for li, l in enumerate(self.style_layer):
                    if l != name:
                        continue
                    ch = (style_grams[li][0] if isinstance(style_grams[li], (list, tuple)) else style_grams[li]).shape[0]
                    P = self.net.feature_pixels(acts, name) if hasattr(self.net, 'feature_pixels') else \
                        self._feature_pixels(hw[0], hw[1], name)
                    coef = self.w_style * self.w_style_layer[li] * 4.0 / (2.0 * P * ch)
                    if group:                                      # cotangent on image 0, zeros on the others
                        if g is None:
                            g = torch.zeros_like(self.net.features_f32(acts, name))
                        self.net.gram_grad(acts_style, name, handles[l], coef, g[:1], is_conv)
                    else:
                        g = self.net.gram_grad(acts, name, handles[l], coef, g, is_conv)
            if self.w_content and self.net2 is None and self.content_layer == name:
                g = self.net.content(acts, name, self.content_channel, self.w_content * share, loss, g, is_conv,
                                     target=getattr(self, '_content_feat', None), amp=self.w_content_amp)
            return g

        g_x = self.net.backward(x, acts, wanted, add_loss_grad, set(wanted), gray=gray is not None) if wanted else None
        if g_x is None:
            g_x = torch.zeros_like(x if gray is None else gray)
        if mask_grads is not None and style_on and style_masks:
            for l in self.style_layer:                             # style_side: {layer: handles} under style_mask_on_ref
                side = (self._style_acts[l], style_side[l]) if style_side is not None else None
                mask_grads[l] = self.net.gram_mask_grad(acts, l, handles[l], style_side=side)
        if self.w_content and self.net2 is not None:               # multi-net loss: content term on the second network
            net2, cl = self.net2, self.content_layer
            relu2 = 1 if (cl.startswith('conv') and not hasattr(net2, 'relu_masked')) else 0
            acts2 = net2.forward(x, [cl])

            def add2(name, g):
                return net2.content(acts2, name, self.content_channel, self.w_content * share, loss, g, relu2,
                                    target=getattr(self, '_content_feat', None), amp=self.w_content_amp)

            g2 = net2.backward(x, acts2, [cl], add2, {cl})
            ops.axpy(g_x, g2, 1.0)
        if self.w_tv:
            g_tv = torch.empty_like(d_img[0])
            for v in range(n):
                ops.tv_loss(d_img[v], self.w_tv * share, loss[v:v + 1], g_tv)
                ops.axpy(g_x[v], g_tv, 1.0)
        return g_x
