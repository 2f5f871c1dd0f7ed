"""2-D particle colour styler -- drop-in for the reference's ``styler_2p.Styler``
(``styler_2p.py:14-314``): per-particle RGB colours are optimised so that the SPH colour splat of
a 2-D liquid matches a style image; positions and SPH densities are constants.

    c -> clip(c,0,1) -> SPH colour splat (/ particle density) -> clip(.,0,1) -> x255 -> -mean -> VGG
      -> Gram style loss [+ TV]                                      (``styler_2p.py:42-102``)
"""
import numpy as np
import torch

from . import _lib, ops
from .styler_base import StylerBase, f32
from .styler_3p import _Adam
from .util import octave_sizes
from .vgg import _R_MEAN, _G_MEAN, _B_MEAN


class Styler(StylerBase):
    def __init__(self, self_dict, weights=None, device=None, content_weights=None):
        StylerBase.__init__(self, self_dict, weights=weights, device=device, content_weights=content_weights)
        if self.style_mask and 'vgg' not in self.model_path:    # style_masks_for reads the VGG block number of a layer
            raise NotImplementedError('style_mask with a GraphDef loss network')

    # ---- graph pieces ----------------------------------------------------------------------------
    def _grid(self, res):
        return _lib.make_grid(2, res, self.domain, self.nsize, self.clip)

    def _scale(self):
        return 0.8 * (2 * self.radius) ** 2 * self.rest_density       # mass, transform.py:1349-1352

    def _gray(self, fr, res):
        """clip(p2g(p)/rho0, 0, 1): the density mask, constant per frame/octave (styler_2p.py:55-57,94)."""
        key = (fr['id'], tuple(res))
        if key not in self._cache:
            g = ops.splat_sph_fwd(fr['p'], None, self._grid(res), self.radius * self.support,
                                  self._scale() / self.rest_density)
            self._cache[key] = ops.clip_fwd(g, 0.0, 1.0)
        return self._cache[key]

    def _forward(self, fr, var, res):
        grid = self._grid(res)
        c_ = ops.clip_fwd(var, 0.0, 1.0)                              # :68
        d_raw = ops.splat_sph_fwd(fr['p'], None, grid, self.radius * self.support, self._scale(), pc=c_,
                                  pd=fr['r'], rest_density=self.rest_density)     # :74-75
        d = ops.clip_fwd(d_raw, 0.0, 1.0)                             # :88
        H, W = res
        gray3 = d.reshape(1, H, W, 3)
        nh, nw = self._net_hw((H, W))
        if (nh, nw) != (H, W):
            gray3 = ops.resize_bilinear_fwd(gray3, nh, nw)
        d_img = torch.empty(1, nh, nw, 3, dtype=f32, device=self.device)
        x = torch.empty_like(d_img)
        ops.to_net_input_fwd(gray3, 255.0, d_img, x)                  # styler_base.py:41-45
        return {'grid': grid, 'd_raw': d_raw, 'd': d, 'd_img': d_img, 'x': x, 'hw': (H, W)}

    def loss_and_grad(self, fr, var, res, style_grams, share=1.0):
        """``share``: 1/batch_size -- this frame's weight in the batch-mean terms of a joint loss (content, TV)"""
        st = self._forward(fr, var, res)
        loss = torch.zeros(1, dtype=f32, device=self.device)
        masks = None
        if self.style_mask and style_grams is not None:           # styler_base.py:165-169, test_dambreak2d.py:189
            key = (fr['id'], tuple(res), 'mask')
            if key not in self._cache:                             # d_gray is constant per (frame, octave)
                H_, W_ = res
                self._cache[key] = self.style_masks_for(self._gray(fr, res).reshape(1, H_, W_),
                                                        (st['x'].shape[1], st['x'].shape[2]))
            masks = self._cache[key]
            if self.style_mask_on_ref:                             # the mask is constant per (frame, octave): so is the target
                gkey = (fr['id'], tuple(res), 'grams_on_ref')
                if gkey not in self._cache:
                    self._cache[gkey] = self.masked_style_grams(masks)
                style_grams = self._cache[gkey]
        g_x = self.image_loss_and_grad(st['x'], st['d_img'], style_grams, loss, style_masks=masks, share=share)
        H, W = st['hw']
        g_d3 = ops.to_net_input_bwd(g_x, 3, 255.0, torch.empty(1, g_x.shape[1], g_x.shape[2], 3, dtype=f32,
                                                                device=self.device))
        if (g_x.shape[1], g_x.shape[2]) != (H, W):
            g_d3 = ops.resize_bilinear_bwd(g_d3, H, W)
        g_raw = ops.clip_bwd(g_d3.reshape(H, W, 3), st['d_raw'], 0.0, 1.0)
        g_c = ops.splat_sph_bwd_color(fr['p'], st['grid'], self.radius * self.support, self._scale(), fr['r'], 3,
                                      self.rest_density, g_raw)
        grad = ops.clip_bwd(g_c, var, 0.0, 1.0)
        return loss, grad

    def init_colors(self, n):
        """``styler_2p.py:189-192``: U(-5,5) + RGB mean, /255, a different draw per frame."""
        c = self.rng.uniform(-5, 5, [self.num_frames, n, 3]).astype(np.float32)
        c += np.array([_R_MEAN, _G_MEAN, _B_MEAN])
        c /= 255
        return c

    # ---- the optimisation loop (styler_2p.py:165-314) ------------------------------------------------
    def run(self, params, c_init=None):
        dev = self.device
        nf = self.num_frames
        if nf % self.batch_size:
            raise ValueError('num_frames must be a multiple of batch_size (the reference feeds p[t+i], styler_2p.py:237-240)')
        oct_size = octave_sizes(self.resolution, self.octave_n, self.octave_scale)
        frames = []
        for i in range(nf):
            frames.append({'id': i,
                           'p': torch.as_tensor(np.asarray(params['p'][i]), dtype=f32).to(dev).contiguous(),
                           'r': torch.as_tensor(np.asarray(params['r'][i]), dtype=f32).to(dev).contiguous()})
        if c_init is None:
            c_init = self.init_colors(frames[0]['p'].shape[0])
        g_opt = [torch.tensor(np.asarray(c_init[i]), dtype=f32).to(dev).contiguous() for i in range(nf)]   # copies

        loss_history, d_intm, opt_ = [], [], {}
        for octave in range(self.octave_n):
            res = oct_size[octave]
            self._cache = {}
            style_grams = None
            if self.w_style and self.style_img is not None:
                style_grams = self._style_feature(self.style_img, res)
            self._content_feat = None
            if self.w_content and self.content_img is not None:
                self._content_feat = self._content_feature(self.content_img, res)
            lr = self.lr[octave] if isinstance(self.lr, list) else self.lr
            loss_o, intm_o = [], []
            for step in range(self.iter):
                deltas = []
                # batch_size frames per sess.run (:236-262): ONE joint loss -- Gram terms summed over the batch, content
                # and TV means over it -- and one Adam op over the batch's variables: slot i of the group's optimizer is
                # shared by every frame fed into position i; no term couples the frames, so they run one after another
                B = self.batch_size
                for t in range(0, nf, B):
                    joint = None
                    for i in range(B):
                        fr = frames[t + i]
                        var = g_opt[t + i].clone()
                        adam = opt_.setdefault((t // self.frames_per_opt, i), _Adam())
                        l, grad = self.loss_and_grad(fr, var, res, style_grams, share=1.0 / B)
                        adam.step(var, grad, lr)
                        joint = l[0] if joint is None else joint + l[0]
                        deltas.append(ops.iterate_delta(var, 1.0, g_opt[t + i], None, 0, torch.empty_like(var)))   # :260-262
                        if step == self.iter - 1 and octave < self.octave_n - 1:
                            intm_o.append(self._out_image(fr, var, res))
                    loss_o.append(joint)
                if self.window_sigma > 0 and nf > 1:                                 # :276-277
                    sm = ops.temporal_gauss(torch.stack(deltas, 0), self.window_sigma)
                    deltas = [sm[j] for j in range(nf)]
                for t in range(nf):
                    ops.axpy(g_opt[t], deltas[t].contiguous(), 1.0)
            loss_history.append([float(v) for v in torch.stack(loss_o).cpu().tolist()] if loss_o else [])
            if octave < self.octave_n - 1:
                d_intm.append(np.stack(intm_o, 0))

        result = {'l': loss_history, 'd_intm': d_intm}
        res = oct_size[-1]
        self._cache = {}
        c_sty, d_sty = [], []
        for t in range(nf):
            fr = frames[t]
            dens = ops.clip_fwd((fr['r'] * (1.0 / self.rest_density)).contiguous(), 0.0, 1.0)     # :71
            c_sty.append(ops.mul_bcast(ops.clip_fwd(g_opt[t], 0.0, 1.0), dens.reshape(-1)).cpu().numpy())
            d_sty.append(self._out_image(fr, g_opt[t], res))
        result['c'] = c_sty
        result['d'] = np.array(d_sty)
        result['g_opt'] = [g.cpu().numpy() for g in g_opt]
        return result

    def _out_image(self, fr, var, res):
        """(d * d_gray * 255) as uint8 (styler_2p.py:100, 264-265, 306)."""
        st = self._forward(fr, var, res)
        out = ops.mul_bcast(st['d'], self._gray(fr, res).reshape(-1))
        return (out * 255).cpu().numpy().astype(np.uint8)
