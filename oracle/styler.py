"""Oracle restatement of the reference's Styler graphs and ``run()`` loops.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

* ``Oracle3P``  <- ``styler_3p.py:14-164`` (graph) and ``:229-438`` (loop)
* ``Oracle2P``  <- ``styler_2p.py:14-102`` (graph) and ``:165-314`` (loop)

``cfg`` is the reference's flat namespace (``config.py`` flags + driver-added attributes
``num_kernels``, ``kernel_scale``, ``rng``).  The pretrained network file is replaced by an
explicit ``weights`` dict (see ``oracle.vgg``); style/content targets are passed already
resized, one per octave (the reference resizes with skimage, which is unavailable here).

Two view modes for the multi-view 3-D loop:
  ``sequential`` -- the reference: one Adam step per view, iterates averaged
                    (``styler_3p.py:329-352``);
  ``allreduce``  -- mean gradient over views, ONE Adam step (the data-parallel variant
                    BASELINE.json's north_star asks for; a documented semantic change).

NaN handling.  ``p2g_wavg``'s ``where(wmap>eps, num/wmap, num)`` gives a NaN
gradient to every variable element that has an empty (wmap == 0) cell among its splat targets
(``transform.py:1703``).  The reference lets the TF variable (and Adam's m, v) go NaN and maps
NaN->0 only on the host copies (``styler_3p.py:337-340,360``), so the END-OF-ITERATION value of such
an element is always 0.  Inside a multi-view iteration the NaN variable is fed to the next view's
forward pass, where ``tf.clip_by_value`` = ``maximum(minimum(x, 1), -1)`` runs through the GPU
functors (``fminf/fmaxf``: the non-NaN operand wins) and reads it as +1 -- reproduced here (``_clip``)
and pinned by ``tests/golden/ref_density_sequential.npz``, the reference's own loop executed on the
TF stand-in.  (On TF's CPU kernels the NaN would propagate instead; the reference targets the GPU.)
"""
import numpy as np
import torch
from scipy.ndimage import gaussian_filter

from . import transform as T
from . import render as R
from . import vgg as V
from . import loss as L
from .adam import TFAdam


def octave_sizes(resolution, octave_n, octave_scale):
    """``styler_3p.py:241-246`` / ``styler_2p.py:173-178`` (np.int -> int)."""
    out, s = [], np.array(resolution)
    for _ in range(octave_n):
        out.append([int(v) for v in s])
        s = (s // octave_scale).astype(int)
    out.reverse()
    return out


def _nan_to_num(x):
    return torch.nan_to_num(x)  # NaN->0, +-inf->+-finfo.max, like np.nan_to_num


def _clip(x, lo, hi):
    """``tf.clip_by_value`` as TF-1.15 builds it: ``maximum(minimum(x, hi), lo)`` (``clip_ops.py``) with
    the GPU functors' NaN rule (Eigen ``numext::mini/maxi`` -> CUDA ``fminf/fmaxf``, IEEE minNum/maxNum:
    the non-NaN operand wins), so a NaN variable element reads as ``hi``.  Gradient: passes where
    ``lo <= x <= hi`` (``_MaximumMinimumGrad``), zero for NaN."""
    lo = torch.as_tensor(lo, dtype=x.dtype)
    hi = torch.as_tensor(hi, dtype=x.dtype)
    return torch.fmax(torch.fmin(x, hi), lo)


class _Base:
    def __init__(self, cfg, weights, dtype=torch.float32, content_weights=None):
        self.cfg = cfg
        # multi-net loss (engine extension, BASELINE.json configs[4]): content loss on a second GraphDef network
        self.nodes2 = list(content_weights) if content_weights is not None else None
        self.dtype = dtype
        self.nodes = None
        if isinstance(weights, (list, tuple)):         # a GraphDef loss network (inception5h): styler_base.py:17-31
            self.nodes = list(weights)
        else:
            self.weights = {k: (w.to(dtype), b.to(dtype)) for k, (w, b) in weights.items()}
        self.model = 'vgg_16' if '16' in cfg.network else 'vgg_19'

    # styler_base.py:91-94
    def _layer(self, ep, name):
        return ep['input'] if 'input' in name else ep[name]

    def _net(self, d_img):
        # only the prefix up to the deepest requested end point is evaluated (the reference
        # builds the whole net; the unused tail does not influence any fetched tensor)
        c = self.cfg
        want = set()
        if c.w_style:
            want |= set(c.style_layer)
        if c.w_content:
            want.add(c.content_layer)
        want.discard('input')
        if self.nodes2 is not None and c.w_content:
            from . import graphnet
            want.discard(c.content_layer)
            ep2 = graphnet.forward(d_img, self.nodes2, [c.content_layer], pool1=bool(getattr(c, 'pool1', False)))
            ep = dict(self._net_first(d_img, want)) if want else {'input': d_img}
            ep[c.content_layer] = ep2[c.content_layer]
            return ep
        return self._net_first(d_img, want)

    def _net_first(self, d_img, want):
        c = self.cfg
        if self.nodes is not None:
            from . import graphnet
            return graphnet.forward(d_img, self.nodes, sorted(want), pool1=bool(getattr(c, 'pool1', False)))
        order = [n for n, _, _ in V.layer_specs(self.model)]
        order_all = []
        for b in range(1, 6):
            order_all += [n for n in order if n.startswith('conv%d_' % b)] + ['pool%d' % b]
        last = None
        for n in order_all:
            if n in want:
                last = n
        return V.forward(d_img, self.weights, self.model, upto=last)

    # styler_base.py:249-278 (RGBA mask branch not restated: its result is discarded at HEAD,
    # ``style_feature *= mask`` multiplies the graph tensor, not the fetched array)
    def style_features(self, style_img):
        c = self.cfg
        img = torch.as_tensor(np.asarray(style_img)[..., :3], dtype=self.dtype)
        ep = self._net(torch.stack([img] * c.batch_size, 0))
        return [self._layer(ep, l).detach() for l in c.style_layer]

    def content_feature(self, content_img):
        c = self.cfg
        img = torch.as_tensor(np.asarray(content_img)[..., :3], dtype=self.dtype)
        ep = self._net(torch.stack([img] * c.batch_size, 0))
        feat = self._layer(ep, c.content_layer).detach()
        if getattr(c, 'top_k', 0) > 0:                         # styler_base.py:240-246
            assert 'softmax2_pre_activation' in c.content_layer
            feat = feat.clone()
            idx = torch.topk(feat.abs(), c.top_k, dim=1).indices
            keep = torch.zeros_like(feat, dtype=torch.bool).scatter_(1, idx, True)
            feat[~keep] = 0
        return feat

    # styler_base.py:127-231
    def total_loss(self, g, style_feats=None, content_feat=None):
        c = self.cfg
        total = 0
        ep = None
        if c.w_content or (c.w_style and style_feats is not None):
            ep = self._net(g['d_img'])
        if c.w_content:
            total = total + c.w_content * L.content_loss(
                self._layer(ep, c.content_layer), c.content_channel, content_feat, c.w_content_amp)
        if c.w_style and style_feats is not None:
            feats = [self._layer(ep, l) for l in c.style_layer]
            sl, _ = L.style_loss(feats, style_feats, c.w_style_layer, c.batch_size,
                                 d_gray=g['d_gray'], style_mask=c.style_mask,
                                 style_mask_on_ref=c.style_mask_on_ref)
            total = total + c.w_style * sl
        if c.w_tv:
            total = total + c.w_tv * L.tv_loss(g['d_img'])
        if getattr(c, 'w_density', 0) > 0 and g.get('d_vars') is not None:
            total = total + c.w_density * L.density_loss(g['d_vars'])
        if getattr(c, 'w_pressure', 0) > 0 and g.get('pressure') is not None:
            total = total + c.w_pressure * L.pressure_loss(g['pressure'])
        return total


class Oracle3P(_Base):
    def graph(self, p_list, r_list, var_list, res, rot_mats=None):
        """``styler_3p.py:42-164`` for one feed.  Lists hold ``batch_size`` frames."""
        c = self.cfg
        d, p_out, d_vars, pressure = [], [], [], []
        for i in range(len(p_list)):
            p_ = p_list[i].unsqueeze(0)
            if 'p' in c.target_field:
                p_ = p_ + var_list[i].unsqueeze(0)
            p_out.append(p_[0])
            if 'd' in c.target_field:
                r_opt = _clip(var_list[i], -1, 1)                # :74
                d_vars.append(r_opt)
                r_ = (r_list[i] + r_opt).unsqueeze(0)            # :76
                d_ = 0
                for k in range(c.num_kernels):                    # :79-87
                    support = c.support / (c.kernel_scale ** k)
                    d_ = d_ + T.p2g_wavg(p_, r_[..., k:k + 1], c.domain, res, c.radius, c.nsize,
                                         support=support, clip=c.clip, is_2d=False)
            else:
                d_ = T.p2g(p_, c.domain, res, c.radius, c.rest_density, c.nsize,
                           support=c.support, clip=c.clip, is_2d=False) / c.rest_density  # :90-91
            d.append(d_)
            if c.w_pressure > 0 and 'p' in c.target_field:
                pressure.append(d_)
        d = torch.cat(d, 0)
        d_out = R.field_post(d, c.k)                             # :112-128
        dr = d_out
        if c.rotate:
            dr = T.rotate(d_out, rot_mats)                       # :132-134
        d_gray = R.render(dr, c.transmit, c.render_liquid)       # :148-161
        d_img = R.to_loss_net_input(d_gray, c.resize_scale, c.target_field)
        return {'d_out': d_out, 'd_gray': d_gray, 'd_img': d_img, 'p_out': p_out,
                'd_vars': d_vars if d_vars else None,
                'pressure': torch.cat(pressure, 0) if pressure else None}

    def views(self):
        c = self.cfg
        return T.rot_mat(c.phi0, c.phi1, c.phi_unit, c.theta0, c.theta1, c.theta_unit,
                         sample_type=c.sample_type, rng=c.rng, nv=c.n_views)

    def loss_and_grad(self, p_list, r_list, var_list, res, rot_mats, style_feats, content_feat):
        vs = [v.detach().clone().requires_grad_(True) for v in var_list]
        g = self.graph(p_list, r_list, vs, res, rot_mats)
        loss = self.total_loss(g, style_feats, content_feat)
        grads = torch.autograd.grad(loss, vs)
        return float(loss.detach()), [gr.detach() for gr in grads]

    def run(self, params, style_targets=None, content_targets=None, view_mode='sequential'):
        """``styler_3p.py:229-438``."""
        c = self.cfg
        dt = self.dtype
        nf, B = c.num_frames, c.batch_size
        lr_list = None
        if abs(c.lr_scale - 1) > 1e-7:                                   # :237-238
            lr_list = [c.lr / c.lr_scale ** i for i in range(c.octave_n)]
        oct_size = octave_sizes(c.resolution, c.octave_n, c.octave_scale)
        p = [torch.as_tensor(x, dtype=dt) for x in params['p']]
        r = [torch.as_tensor(x, dtype=dt) for x in params['r']] if 'd' in c.target_field else None
        width = 3 if 'p' in c.target_field else c.num_kernels
        g_opt = [torch.zeros(p[i].shape[0], width, dtype=dt) for i in range(nf)]
        n_views = None
        rot_mats = None
        if c.rotate:
            rot_mats, views = self.views()                                # :137-145
            n_views = c.n_views if c.n_views is not None else len(views)
            assert n_views % c.v_batch == 0
        eye = [np.identity(3)] * B

        loss_history, d_intm, opt_ = [], [], {}
        for octave in range(c.octave_n):
            res = oct_size[octave]
            loss_o, intm_o = [], []
            content_feat = None
            if content_targets is not None:
                content_feat = self.content_feature(content_targets[octave])
            style_feats = None
            if style_targets is not None:
                style_feats = self.style_features(style_targets[octave])
            lr = lr_list[octave] if lr_list is not None else c.lr
            for step in range(c.iter):
                g_tmp = [None] * nf
                for t in range(0, nf, B * c.interp):
                    fr = [t + i * c.interp for i in range(B)]
                    pl = [p[f] for f in fr]
                    rl = [r[f] for f in fr] if r is not None else None
                    var = [g_opt[f].clone() for f in fr]                  # :312
                    opt_id = t // c.frames_per_opt
                    if opt_id not in opt_:
                        opt_[opt_id] = [TFAdam() for _ in range(B)]       # :315-323
                    adam = opt_[opt_id]
                    if c.rotate:
                        l_ = []
                        if view_mode == 'sequential':
                            acc = None
                            for i in range(0, n_views, c.v_batch):        # :329-340
                                lv, gr = self.loss_and_grad(pl, rl, var, res, rot_mats[i:i + c.v_batch],
                                                            style_feats, content_feat)
                                l_.append(lv)
                                var = [adam[j].step(var[j], gr[j], lr) for j in range(B)]
                                if acc is None:
                                    acc = [_nan_to_num(v) for v in var]
                                else:
                                    acc = [a + _nan_to_num(v) for a, v in zip(acc, var)]
                            g_new = [a / (n_views / c.v_batch) for a in acc]   # :351-352
                        else:
                            gsum = None
                            for i in range(0, n_views, c.v_batch):
                                lv, gr = self.loss_and_grad(pl, rl, var, res, rot_mats[i:i + c.v_batch],
                                                            style_feats, content_feat)
                                l_.append(lv)
                                gsum = gr if gsum is None else [a + b for a, b in zip(gsum, gr)]
                            nb = n_views // c.v_batch
                            var = [adam[j].step(var[j], gsum[j] / nb, lr) for j in range(B)]
                            g_new = var
                        loss_o.append(float(np.mean(l_)))                  # :342
                        if 'uniform' not in c.sample_type:                # :344-349
                            rot_mats, views = self.views()
                    else:
                        lv, gr = self.loss_and_grad(pl, rl, var, res, None, style_feats, content_feat)
                        loss_o.append(lv)                                  # :354-357
                        var = [adam[j].step(var[j], gr[j], lr) for j in range(B)]
                        g_new = var
                    for i, f in enumerate(fr):                            # :359-363
                        g_tmp[f] = _nan_to_num(g_new[i]) - g_opt[f]
                        if 'd' in c.target_field:
                            g_tmp[f] = g_tmp[f] * r[f][..., 0:1]
                    if step == c.iter - 1 and octave < c.octave_n - 1:    # :365-370
                        with torch.no_grad():
                            g = self.graph(pl, rl, var, res, eye if c.rotate else None)
                        intm_o.append(g['d_img'].numpy().astype(np.uint8))
                key = list(range(0, nf, c.interp))
                if c.window_sigma > 0 and nf > 1:                          # :382-383
                    stack = np.stack([g_tmp[f].numpy() for f in key], 0)
                    stack = gaussian_filter(stack, sigma=(c.window_sigma, 0, 0))
                    for j, f in enumerate(key):
                        g_tmp[f] = torch.as_tensor(stack[j])
                for f in key:                                              # :385-386
                    g_opt[f] = g_opt[f] + g_tmp[f]
            loss_history.append(loss_o)
            if octave < c.octave_n - 1:
                d_intm.append(np.concatenate(intm_o, 0))

        if c.interp > 1:                                                   # :392-397
            w = np.linspace(0, 1, c.interp + 1)
            for t in range(0, nf - 1, c.interp):
                for i in range(1, c.interp):
                    g_opt[t + i] = g_opt[t] * (1 - w[i]) + g_opt[t + c.interp] * w[i]

        result = {'l': loss_history, 'd_intm': d_intm, 'v': None, 'c': None, 'g_opt': g_opt}
        p_sty, v_sty, d_sty, r_sty = [None] * nf, [None] * nf, [None] * nf, [None] * nf
        res = oct_size[-1]
        for t in range(0, nf, B):                                          # :409-431
            fr = list(range(t, t + B))
            with torch.no_grad():
                g = self.graph([p[f] for f in fr], [r[f] for f in fr] if r is not None else None,
                               [g_opt[f] for f in fr], res, eye if c.rotate else None)
            for i, f in enumerate(fr):
                p_sty[f] = g['p_out'][i].numpy()
                if 'p' in c.target_field:
                    v_sty[f] = g_opt[f].numpy()
                d_sty[f] = g['d_out'][i].numpy()
                r_sty[f] = g['d_img'][i].numpy().astype(np.uint8)
        result['p'] = p_sty
        if 'p' in c.target_field:
            result['v'] = v_sty
        result['d'] = np.array(d_sty)
        result['r'] = np.array(r_sty)
        return result


class Oracle2P(_Base):
    def graph(self, p_list, r_list, var_list, res):
        """``styler_2p.py:42-102``."""
        c = self.cfg
        d, d_gray, col = [], [], []
        for i in range(len(p_list)):
            p_ = p_list[i].unsqueeze(0)
            r_ = r_list[i].unsqueeze(0)
            dg = T.p2g(p_, c.domain, res, c.radius, c.rest_density, c.nsize, support=c.support,
                       clip=c.clip) / c.rest_density                       # :55-57
            d_gray.append(dg)
            c_ = _clip(var_list[i].unsqueeze(0), 0, 1)                     # :68
            col.append(c_[0] * torch.clamp(r_[0] / c.rest_density, 0, 1))  # :71
            d.append(T.p2g(p_, c.domain, res, c.radius, c.rest_density, c.nsize, support=c.support,
                           clip=c.clip, pc=c_, pd=r_))                     # :74-75
        d = torch.clamp(torch.cat(d, 0), 0, 1)                             # :88
        d_gray = torch.clamp(torch.cat(d_gray, 0), 0, 1)                   # :94
        d_img = R.to_loss_net_input(d, c.resize_scale, c.target_field)
        return {'d_out': d * d_gray, 'd_gray': d_gray, 'd_img': d_img, 'c': col,
                'd_vars': None, 'pressure': None}

    def init_colors(self, n):
        """``styler_2p.py:189-192``."""
        c = self.cfg
        c_opt = c.rng.uniform(-5, 5, [c.num_frames, n, 3]).astype(np.float32)
        c_opt += np.array(V.MEAN_RGB)
        c_opt /= 255
        return c_opt

    def run(self, params, style_targets=None, content_targets=None, c_init=None):
        """``styler_2p.py:165-314``."""
        c = self.cfg
        dt = self.dtype
        nf, B = c.num_frames, c.batch_size
        oct_size = octave_sizes(c.resolution, c.octave_n, c.octave_scale)
        p = [torch.as_tensor(x, dtype=dt) for x in params['p']]
        r = [torch.as_tensor(x, dtype=dt) for x in params['r']]
        if c_init is None:
            c_init = self.init_colors(p[0].shape[0])
        g_opt = [torch.as_tensor(c_init[i], dtype=dt) for i in range(nf)]
        loss_history, d_intm, opt_ = [], [], {}
        for octave in range(c.octave_n):
            res = oct_size[octave]
            loss_o, intm_o = [], []
            content_feat = None
            if content_targets is not None:
                content_feat = self.content_feature(content_targets[octave])
            style_feats = None
            if style_targets is not None:
                style_feats = self.style_features(style_targets[octave])
            lr = c.lr[octave] if isinstance(c.lr, list) else c.lr
            for step in range(c.iter):
                g_tmp = [None] * nf
                for t in range(0, nf, B):
                    fr = list(range(t, t + B))
                    pl, rl = [p[f] for f in fr], [r[f] for f in fr]
                    var = [g_opt[f].clone() for f in fr]
                    opt_id = t // c.frames_per_opt
                    if opt_id not in opt_:
                        opt_[opt_id] = [TFAdam() for _ in range(B)]
                    adam = opt_[opt_id]
                    vs = [v.detach().clone().requires_grad_(True) for v in var]
                    g = self.graph(pl, rl, vs, res)
                    loss = self.total_loss(g, style_feats, content_feat)
                    grads = torch.autograd.grad(loss, vs)
                    loss_o.append(float(loss.detach()))
                    var = [adam[j].step(var[j], grads[j].detach(), lr) for j in range(B)]
                    for i, f in enumerate(fr):
                        g_tmp[f] = _nan_to_num(var[i]) - g_opt[f]          # :260-262
                    if step == c.iter - 1 and octave < c.octave_n - 1:
                        with torch.no_grad():
                            g = self.graph(pl, rl, var, res)
                        intm_o.append((g['d_out'].numpy() * 255).astype(np.uint8))
                if c.window_sigma > 0 and nf > 1:                          # :276-277
                    stack = gaussian_filter(np.stack([x.numpy() for x in g_tmp], 0),
                                            sigma=(c.window_sigma, 0, 0))
                    g_tmp = [torch.as_tensor(stack[j]) for j in range(nf)]
                for f in range(nf):
                    g_opt[f] = g_opt[f] + g_tmp[f]
            loss_history.append(loss_o)
            if octave < c.octave_n - 1:
                d_intm.append(np.concatenate(intm_o, 0))
        result = {'l': loss_history, 'd_intm': d_intm, 'g_opt': g_opt}
        c_sty, d_sty = [None] * nf, [None] * nf
        res = oct_size[-1]
        for t in range(0, nf, B):
            fr = list(range(t, t + B))
            with torch.no_grad():
                g = self.graph([p[f] for f in fr], [r[f] for f in fr], [g_opt[f] for f in fr], res)
            for i, f in enumerate(fr):
                c_sty[f] = g['c'][i].numpy()
                d_sty[f] = (g['d_out'][i].numpy() * 255).astype(np.uint8)
        result['c'] = c_sty
        result['d'] = np.array(d_sty)
        return result
