"""Tensor-level wrappers over the C-ABI (``_lib``): allocate outputs with torch, pass raw
pointers + the current CUDA stream.  No arithmetic happens here."""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import ptr

f32 = torch.float32


def _s(t):
    return _lib.stream_ptr(t.device)


def _u8(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _harr(vals):
    return (C.c_float * len(vals))(*[float(v) for v in vals])


def cells(grid):
    return (grid.res[0] if grid.dim == 3 else 1) * grid.res[1] * grid.res[2]


# ---- splats --------------------------------------------------------------------------------
def splat_sph_fwd(p, disp, grid, h, scale, out=None, pc=None, pd=None, rest_density=1000.0):
    n = p.shape[0]
    ch = 1 if pc is None else pc.shape[-1]
    shape = ([grid.res[0]] if grid.dim == 3 else []) + [grid.res[1], grid.res[2]] + ([ch] if pc is not None else [])
    if out is None:
        out = torch.zeros(shape, dtype=f32, device=p.device)
    else:
        out.zero_()
    _lib.get().call('lnst_splat_sph_fwd', ptr(p), ptr(disp), n, C.byref(grid), h, scale, ptr(pc), ptr(pd), ch,
                    rest_density, ptr(out), _s(p))
    return out


def splat_sph_bwd_pos(p, disp, grid, h, scale, g_out):
    g_p = torch.empty_like(p)
    _lib.get().call('lnst_splat_sph_bwd_pos', ptr(p), ptr(disp), p.shape[0], C.byref(grid), h, scale, ptr(g_out),
                    ptr(g_p), _s(p))
    return g_p


def splat_sph_bwd_color(p, grid, h, scale, pd, ch, rest_density, g_out):
    g_pc = torch.empty(p.shape[0], ch, dtype=f32, device=p.device)
    _lib.get().call('lnst_splat_sph_bwd_color', ptr(p), p.shape[0], C.byref(grid), h, scale, ptr(pd), ch,
                    rest_density, ptr(g_out), ptr(g_pc), _s(p))
    return g_pc


def splat_wavg_wmap(p, grid, hs):
    wmap = torch.empty(len(hs), cells(grid), dtype=f32, device=p.device)
    _lib.get().call('lnst_splat_wavg_wmap', ptr(p), p.shape[0], C.byref(grid), _harr(hs), len(hs), ptr(wmap), _s(p))
    return wmap


def _b(box):
    return C.byref(box) if box is not None else None


def splat_wavg_fwd(p, r, var, grid, hs, wmap, num, out, box=None):
    """With ``box``: only its cells are combined; ``num`` must be zero on entry and is zero again on exit."""
    _lib.get().call('lnst_splat_wavg_fwd_box', ptr(p), ptr(r), ptr(var), p.shape[0], C.byref(grid), _harr(hs),
                    len(hs), ptr(wmap), ptr(num), ptr(out), _b(box), _s(p))
    return out


def splat_cells(p, grid):
    """(cell int32 [N] -- linear index over the unflipped [D,H,W] grid, -1 outside / padding, -2 rounded onto the far face;
    rel fp32 [N,3] -- offset from the cell centre): what the per-cell particle lists of the gather splat are built from"""
    n = p.shape[0]
    cell = torch.empty(n, dtype=torch.int32, device=p.device)
    rel = torch.empty(n, 3, dtype=f32, device=p.device)
    _lib.get().call('lnst_splat_cells', ptr(p), n, C.byref(grid), _u8(cell), ptr(rel), _s(p))
    return cell, rel


def cell_lists(p, grid):
    """(cstart int32 [V+1], order int32 [Nv], rel fp32 [Nv,3]) or None when a particle's cell index rounds out of the grid"""
    cell, rel = splat_cells(p, grid)
    if bool((cell == -2).any()):
        return None
    V = cells(grid)
    valid = torch.nonzero(cell >= 0).flatten()
    cv = cell[valid].to(torch.int64)
    srt = torch.argsort(cv, stable=True)
    order = valid[srt]
    counts = torch.bincount(cv, minlength=V)
    cstart = torch.zeros(V + 1, dtype=torch.int32, device=p.device)
    cstart[1:] = torch.cumsum(counts, 0).to(torch.int32)
    return cstart, order.to(torch.int32).contiguous(), rel[order].contiguous()


def splat_wavg_fwd_gather(lists, r, var, grid, hs, out, box=None):
    """gather form of ``splat_wavg_fwd`` (csrc/tiles_tma.cu): no atomics, no num / wmap volumes, TMA store of the tiles"""
    cstart, order, rel = lists
    _lib.get().call('lnst_splat_wavg_fwd_gather', _u8(cstart), _u8(order), ptr(rel), ptr(r), ptr(var), C.byref(grid),
                    _harr(hs), len(hs), ptr(out), _b(box), _s(r))
    return out


def splat_wavg_bwd(p, var, grid, hs, wmap, g_out, g_var):
    _lib.get().call('lnst_splat_wavg_bwd', ptr(p), ptr(var), p.shape[0], C.byref(grid), _harr(hs), len(hs),
                    ptr(wmap), ptr(g_out), ptr(g_var), _s(p))
    return g_var


def splat_wavg_coef(wmap):
    coef = torch.empty_like(wmap)
    _lib.get().call('lnst_splat_wavg_coef', ptr(wmap), wmap.shape[0], wmap.shape[1], ptr(coef), _s(wmap))
    return coef


def splat_wavg_bwd_coef(p, var, grid, hs, coef, g_out, g_var):
    _lib.get().call('lnst_splat_wavg_bwd_coef', ptr(p), ptr(var), p.shape[0], C.byref(grid), _harr(hs), len(hs),
                    ptr(coef), ptr(g_out), ptr(g_var), _s(p))
    return g_var


# ---- field ---------------------------------------------------------------------------------
# TMA-tiled volume kernels (csrc/tiles_tma.cu) replace the SIMT ones on the GPU whenever TMA can address the volume:
# rows of a multiple of 4 floats and a 16-byte aligned base.  LNST_TMA=0 (or ``ops.USE_TMA = False``) keeps the SIMT
# kernels -- the A/B switch of the parity tests and microbenchmarks.
USE_TMA = os.environ.get('LNST_TMA', '1') not in ('0', '')
# The TMA-slab ray-march BACKWARD is correct but slower than the gather kernel on the B200 (C3, exact intervals: 170 us
# against 151 us -- its per-slab CTA barriers and 25 % occupancy cost more than the merged atomics save; DESIGN.md
# section 3), so it is opt-in: LNST_TMA_BWD=1 or ``ops.USE_TMA_BWD = True``.
USE_TMA_BWD = os.environ.get('LNST_TMA_BWD', '0') not in ('0', '')
ADVECT_REACH = 2


def _tma_ok(*vols):
    lib = _lib.get()
    if not (USE_TMA and lib.kind == 'cuda' and lib.has_tma):
        return False
    return all(v is None or (v.shape[-1] % 4 == 0 and v.data_ptr() % 16 == 0) for v in vols)


def smooth3_relu_fwd(d, out, k, box=None):
    D, H, W = d.shape
    if int(k) > 0 and _tma_ok(d):
        _lib.get().call('lnst_smooth3_relu_fwd_tma', ptr(d), ptr(out), D, H, W, int(k), _b(box), _s(d))
        return out
    _lib.get().call('lnst_smooth3_relu_fwd_box', ptr(d), ptr(out), D, H, W, int(k), _b(box), _s(d))
    return out


def smooth3_relu_bwd(g_out, out, g_in, k, box=None):
    D, H, W = out.shape
    if int(k) > 0 and _tma_ok(g_out, out):
        _lib.get().call('lnst_smooth3_relu_bwd_tma', ptr(g_out), ptr(out), ptr(g_in), D, H, W, int(k), _b(box), _s(out))
        return g_in
    _lib.get().call('lnst_smooth3_relu_bwd_box', ptr(g_out), ptr(out), ptr(g_in), D, H, W, int(k), _b(box), _s(out))
    return g_in


def fill_box(vol, box, value=0.0):
    D, H, W = vol.shape
    _lib.get().call('lnst_fill_box', ptr(vol), D, H, W, _b(box), float(value), _s(vol))
    return vol


# ---- render --------------------------------------------------------------------------------
def rotate_fwd(vol, rot):
    D, H, W = vol.shape
    nv = rot.shape[0]
    out = torch.empty(nv, D, H, W, dtype=f32, device=vol.device)
    _lib.get().call('lnst_rotate_fwd', ptr(vol), ptr(rot), nv, D, H, W, ptr(out), _s(vol))
    return out


def rotate_bwd(g_out, rot, g_vol=None):
    """d loss / d vol of ``rotate_fwd``: 8-corner scatter of g_out [nv,D,H,W], summed over the views"""
    nv, D, H, W = g_out.shape
    if g_vol is None:
        g_vol = torch.zeros(D, H, W, dtype=f32, device=g_out.device)
    _lib.get().call('lnst_rotate_bwd', ptr(g_out), ptr(rot), nv, D, H, W, ptr(g_vol), _s(g_out))
    return g_vol


def ray_intervals(rot, shape, box, bricks, out=None):
    """int32 [n_views,H,W,2]: inclusive depth range of every ray that can touch the box / occupied bricks."""
    D, H, W = shape
    nv = rot.shape[0]
    if out is None:
        out = torch.empty(nv, H, W, 2, dtype=torch.int32, device=rot.device)
    _lib.get().call('lnst_ray_intervals', ptr(rot), nv, D, H, W, _b(box), _u8(bricks), _u8(out), _s(rot))
    return out


def ray_intervals_exact(rot, shape, box, touch, out=None):
    """int32 [n_views,H,W,2] like ``ray_intervals``, from the per-voxel footprint mask ``touch`` (uint8 [D,H,W])"""
    D, H, W = shape
    nv = rot.shape[0]
    if out is None:
        out = torch.empty(nv, H, W, 2, dtype=torch.int32, device=rot.device)
    _lib.get().call('lnst_ray_intervals_exact', ptr(rot), nv, D, H, W, _b(box), _u8(touch), _u8(out), _s(rot))
    return out


def raymarch_fwd(vol, rot, tau, liquid, img, stot, box=None, intervals=None, stats=None):
    """``stats`` [2 nv] (zero on entry): the kernel also reduces stats[2v] = max of view v (rotated march only)."""
    D, H, W = vol.shape
    nv = 1 if rot is None else rot.shape[0]
    if rot is not None and min(D, H, W) >= 2 and D * H * W < 2 ** 31 - 1 and _tma_ok(vol):
        _lib.get().call('lnst_raymarch_fwd_max_tma', ptr(vol), ptr(rot), nv, D, H, W, float(tau), int(bool(liquid)), _b(box),
                        _u8(intervals), ptr(img), ptr(stot), ptr(stats), _s(vol))
        return img, stot
    _lib.get().call('lnst_raymarch_fwd_max_box', ptr(vol), ptr(rot), nv, D, H, W, float(tau), int(bool(liquid)), _b(box),
                    _u8(intervals), ptr(img), ptr(stot), ptr(stats), _s(vol))
    return img, stot


def raymarch_bwd(vol, rot, tau, liquid, stot, g_img, g_vol, box=None, intervals=None, norm=None):
    """``norm`` = (img, stats, dots): g_img is the cotangent of img / max(img) and the kernel applies the normalisation's
    gradient while loading it (rotated march only)."""
    D, H, W = vol.shape
    nv = 1 if rot is None else rot.shape[0]
    if norm is None and USE_TMA_BWD and rot is not None and not liquid and min(D, H, W) >= 2 and D * H * W < 2 ** 30 - 1 \
            and _tma_ok(vol):
        _lib.get().call('lnst_raymarch_bwd_tma', ptr(vol), ptr(rot), nv, D, H, W, float(tau), _b(box), _u8(intervals),
                        ptr(stot), ptr(g_img), ptr(g_vol), _s(vol))
        return g_vol
    n_img, n_stats, n_dots = norm if norm is not None else (None, None, None)
    _lib.get().call('lnst_raymarch_bwd_norm_box', ptr(vol), ptr(rot), nv, D, H, W, float(tau), int(bool(liquid)), _b(box),
                    _u8(intervals), ptr(stot), ptr(g_img), ptr(n_img), ptr(n_stats), ptr(n_dots), ptr(g_vol), _s(vol))
    return g_vol


def normalize_ties_fwd(img, stats, gray):
    """gray = img / max and stats[2v+1] = ties in one pass (stats[2v] final, stats[2v+1] zero on entry)"""
    _lib.get().call('lnst_normalize_ties_fwd', ptr(img), ptr(stats), img.shape[0], img[0].numel(), ptr(gray), _s(img))
    return gray


def image_max(img, stats):
    nv = img.shape[0]
    _lib.get().call('lnst_image_max', ptr(img), nv, img[0].numel(), ptr(stats), _s(img))
    return stats


def normalize_fwd(img, stats, gray):
    _lib.get().call('lnst_normalize_fwd', ptr(img), ptr(stats), img.shape[0], img[0].numel(), ptr(gray), _s(img))
    return gray


def normalize_bwd(img, stats, g_gray, dots, g_img):
    _lib.get().call('lnst_normalize_bwd', ptr(img), ptr(stats), ptr(g_gray), img.shape[0], img[0].numel(),
                    ptr(dots), ptr(g_img), _s(img))
    return g_img


def resize_bilinear_fwd(x, oh, ow):
    n, H, W, ch = x.shape
    out = torch.empty(n, oh, ow, ch, dtype=f32, device=x.device)
    _lib.get().call('lnst_resize_bilinear_fwd', ptr(x), n, H, W, ch, oh, ow, ptr(out), _s(x))
    return out


def resize_bilinear_bwd(g_out, H, W):
    n, oh, ow, ch = g_out.shape
    g_in = torch.empty(n, H, W, ch, dtype=f32, device=g_out.device)
    _lib.get().call('lnst_resize_bilinear_bwd', ptr(g_out), n, H, W, ch, oh, ow, ptr(g_in), _s(g_out))
    return g_in


def resize_bicubic_fwd(x, oh, ow):
    n, H, W, ch = x.shape
    out = torch.empty(n, oh, ow, ch, dtype=f32, device=x.device)
    _lib.get().call('lnst_resize_bicubic_fwd', ptr(x), n, H, W, ch, oh, ow, ptr(out), _s(x))
    return out


def resize_bicubic_bwd(g_out, H, W):
    """gradient of resize_bicubic_fwd: g_out [n,OH,OW,C] -> [n,H,W,C]"""
    n, OH, OW, ch = g_out.shape
    g_in = torch.empty(n, H, W, ch, dtype=f32, device=g_out.device)
    _lib.get().call('lnst_resize_bicubic_bwd', ptr(g_out), n, H, W, ch, OH, OW, ptr(g_in), _s(g_out))
    return g_in


def rowdot(a, b, out, scalar=None, scale=0.0, accumulate=False):
    """out[p] (+)= sum_c a[p,c] b[p,c] + scale*scalar[0]"""
    P, ch = a.shape
    _lib.get().call('lnst_rowdot', ptr(a), ptr(b), ch, P, ptr(scalar), float(scale), int(bool(accumulate)), ptr(out),
                    _s(a))
    return out


def masked_accumulate(t, m, f, relu, g, beta):
    """g = beta*g + t * m[..., None] [* (f > 0)]"""
    _lib.get().call('lnst_masked_accumulate', ptr(t), ptr(m), ptr(f), int(relu), t.shape[-1], float(beta), ptr(g),
                    t.numel(), _s(t))
    return g


def to_net_input_fwd(gray, s, d_img, x):
    """gray [n,H,W,Cg] -> d_img, x [n,H,W,3]"""
    n = gray.shape[0]
    _lib.get().call('lnst_to_net_input_fwd', ptr(gray), n, gray[0].numel() // gray.shape[-1], gray.shape[-1],
                    float(s), ptr(d_img), ptr(x), _s(gray))
    return d_img, x


def to_net_input_bwd(g_x, cg, s, g_gray):
    n = g_x.shape[0]
    _lib.get().call('lnst_to_net_input_bwd', ptr(g_x), n, g_x[0].numel() // 3, cg, float(s), ptr(g_gray), _s(g_x))
    return g_gray


# ---- loss net (fp32) -------------------------------------------------------------------------
def conv3x3_f32(x, w, b, relu, mask=None, y=None):
    n, H, W, cin = x.shape
    cout = w.shape[-1]
    if y is None:
        y = torch.empty(n, H, W, cout, dtype=f32, device=x.device)
    _lib.get().call('lnst_conv3x3_f32', ptr(x), ptr(w), ptr(b), ptr(mask), ptr(y), n, H, W, cin, cout, int(relu),
                    _s(x))
    return y


def avgpool2_fwd(x, y=None):
    n, H, W, ch = x.shape
    if y is None:
        y = torch.empty(n, H // 2, W // 2, ch, dtype=f32, device=x.device)
    _lib.get().call('lnst_avgpool2_fwd', ptr(x), ptr(y), n, H, W, ch, _s(x))
    return y


def avgpool2_bwd(g_y, mask, shape, g_x=None):
    n, H, W, ch = shape
    if g_x is None:
        g_x = torch.empty(n, H, W, ch, dtype=f32, device=g_y.device)
    _lib.get().call('lnst_avgpool2_bwd', ptr(g_y), ptr(mask), ptr(g_x), n, H, W, ch, _s(g_y))
    return g_x


# ---- losses --------------------------------------------------------------------------------
def gram_diff(F, denom, Gs, weight, G, loss):
    """F [P,C] (one image).  G = F^T F/denom - Gs; loss += weight*sum(G^2).  ``denom``: a host number, or
    (den_dev [1] device tensor, den_scale): denom = den_scale * den_dev[0], never read back."""
    P, ch = F.shape
    if isinstance(denom, tuple):
        _lib.get().call('lnst_gram_diff_dev', ptr(F), P, ch, 0.0, ptr(denom[0]), float(denom[1]), ptr(Gs), float(weight),
                        ptr(G), ptr(loss), _s(F))
    else:
        _lib.get().call('lnst_gram_diff', ptr(F), P, ch, float(denom), ptr(Gs), float(weight), ptr(G), ptr(loss), _s(F))
    return G


def scale_by_dev(x, num, den):
    """x * num / (den_scale * den_dev[0]) with den = (den_dev [1] device tensor, den_scale) -> new tensor like x"""
    out = torch.empty_like(x)
    _lib.get().call('lnst_scale_by_dev', ptr(x), x.numel(), float(num), ptr(den[0]), float(den[1]), ptr(out), _s(x))
    return out


def gram_bwd(F, G, coef, beta, relu_mask, g_F):
    P, ch = F.shape
    _lib.get().call('lnst_gram_bwd', ptr(F), ptr(G), P, ch, float(coef), float(beta), int(relu_mask), ptr(g_F), _s(F))
    return g_F


def content_loss(F, channel, weight, loss, g_F, beta, relu_mask):
    P, ch = F.shape
    _lib.get().call('lnst_content_loss', ptr(F), P, ch, int(channel), float(weight), ptr(loss), ptr(g_F), float(beta),
                    int(relu_mask), _s(F))


def content_mse(F, target, amp, weight, loss, g_F, beta, relu_mask):
    """loss += weight*mean((F - target*amp)^2); g_F = beta*g_F + weight*grad [* (F>0)]."""
    _lib.get().call('lnst_content_mse', ptr(F), ptr(target), F.numel(), float(amp), float(weight), ptr(loss), ptr(g_F),
                    float(beta), int(relu_mask), _s(F))


def tv_loss(d_img, weight, loss, g_img):
    H, W, ch = d_img.shape
    _lib.get().call('lnst_tv_loss', ptr(d_img), H, W, ch, float(weight), ptr(loss), ptr(g_img), _s(d_img))


# ---- optimiser / glue ------------------------------------------------------------------------
def adam_step_dev(var, grad, m, v, state, lr, gscale=1.0, beta1=0.9, beta2=0.999, eps=1e-8):
    """TF ApplyAdam with the beta powers / lr_t in the device tensor ``state`` [3] (graph-replayable)."""
    _lib.get().call('lnst_adam_step_dev', ptr(var), ptr(grad), ptr(m), ptr(v), var.numel(), ptr(state), float(lr),
                    beta1, beta2, eps, float(gscale), _s(var))
    return var


def adam_iterate_dev(g_opt, grad, m, v, state, lr, gscale, mask, mask_stride, var_out, delta, apply,
                     beta1=0.9, beta2=0.999, eps=1e-8):
    """var = g_opt; Adam step; delta = (nan_to_num(var) - g_opt) [* mask]; optionally g_opt += delta -- one kernel."""
    width = g_opt.shape[-1]
    _lib.get().call('lnst_adam_iterate_dev', ptr(g_opt), ptr(grad), ptr(m), ptr(v), g_opt.numel(), ptr(state),
                    float(lr), beta1, beta2, eps, float(gscale), ptr(mask), width, int(mask_stride), ptr(var_out),
                    ptr(delta), int(bool(apply)), _s(g_opt))
    return var_out, delta


def adam_step(var, grad, m, v, lr_t, gscale=1.0, beta1=0.9, beta2=0.999, eps=1e-8):
    _lib.get().call('lnst_adam_step', ptr(var), ptr(grad), ptr(m), ptr(v), var.numel(), float(lr_t), beta1, beta2,
                    eps, float(gscale), _s(var))


def sum_scale(x, scale, out=None):
    """out[0] = scale * sum(x) for a few loss slots -- no framework reduction kernel in the replayed step"""
    if out is None:
        out = torch.empty(1, dtype=f32, device=x.device)
    _lib.get().call('lnst_sum_scale', ptr(x), x.numel(), float(scale), ptr(out), _s(x))
    return out


def zeros(n, device):
    """fp32 [n] zeroed on the current stream by the library (a memset node in the step's graph)"""
    x = torch.empty(n, dtype=f32, device=device)
    _lib.get().call('lnst_zero', ptr(x), n, _s(x))
    return x


def iterate_accumulate(acc, var, first):
    _lib.get().call('lnst_iterate_accumulate', ptr(acc), ptr(var), var.numel(), int(first), _s(var))


def iterate_delta(g_new, scale, g_opt, mask, mask_stride, delta):
    width = g_new.shape[-1]
    _lib.get().call('lnst_iterate_delta', ptr(g_new), float(scale), ptr(g_opt), ptr(mask), width, int(mask_stride),
                    g_new.numel(), ptr(delta), _s(g_new))
    return delta


def temporal_gauss(x, sigma):
    """x [T, ...] -> gaussian_filter along axis 0"""
    y = torch.empty_like(x)
    _lib.get().call('lnst_temporal_gauss', ptr(x), ptr(y), x.shape[0], x[0].numel(), float(sigma), _s(x))
    return y


def axpy(y, x, a):
    _lib.get().call('lnst_axpy', ptr(y), ptr(x), float(a), y.numel(), _s(y))
    return y


def clip_fwd(x, lo, hi):
    y = torch.empty_like(x)
    _lib.get().call('lnst_clip_fwd', ptr(x), float(lo), float(hi), ptr(y), x.numel(), _s(x))
    return y


def clip_bwd(g, x, lo, hi, scale=1.0):
    gx = torch.empty_like(x)
    _lib.get().call('lnst_clip_bwd', ptr(g), ptr(x), float(lo), float(hi), float(scale), ptr(gx), x.numel(), _s(x))
    return gx


def mul_bcast(a, b):
    """a [..., C] * b [...] (b broadcast over the last axis of a)"""
    out = torch.empty_like(a)
    _lib.get().call('lnst_mul_bcast', ptr(a), ptr(b), a.shape[-1], ptr(out), a.numel(), _s(a))
    return out


def advect(d, vel):
    """d [X,Y,(Z),C], vel [X,Y,(Z),dim] (normalised units)"""
    dim = vel.shape[-1]
    dims = (C.c_int32 * 3)(*([int(s) for s in d.shape[:dim]] + [1] * (3 - dim)))
    out = torch.empty_like(d)
    if dim == 3 and d.shape[-1] == 1 and d.numel() < 2 ** 31 - 1 and d.shape[2] % 4 == 0 and _tma_ok(d.reshape(d.shape[:3])):
        # scalar field: the source box of every output tile is staged in shared memory by TMA (csrc/tiles_tma.cu);
        # ADVECT_REACH = the back-trace length in cells the box is sized for (longer ones gather from global memory)
        _lib.get().call('lnst_advect3_tma', ptr(d), ptr(vel), d.shape[0], d.shape[1], d.shape[2], int(ADVECT_REACH),
                        ptr(out), _s(d))
        return out
    _lib.get().call('lnst_advect', ptr(d), ptr(vel), dim, dims, d.shape[-1], ptr(out), _s(d))
    return out


# ---- grid -> particle, resimulation step (csrc/gather.cu) ---------------------------------------
def _dims3(shape, dim):
    return (C.c_int32 * 3)(*([int(s) for s in shape[:dim]] + [1] * (3 - dim)))


def g2p(g, p, disp=None, linear=False):
    """g [n0,n1,(n2),C] sampled at p (+disp) [N,dim] (normalised, grid axis order) -> [N,C]
    (transform.py:771-1231)"""
    dim = p.shape[-1]
    out = torch.empty(p.shape[0], g.shape[-1], dtype=f32, device=g.device)
    _lib.get().call('lnst_g2p', ptr(g), dim, _dims3(g.shape, dim), g.shape[-1], ptr(p), ptr(disp), p.shape[0],
                    int(bool(linear)), ptr(out), _s(g))
    return out


def rk4_advect(u, x, time_step, linear=False, want_v=False):
    """x_adv = x + time_step * RK4-blended velocity sampled from u [n0,n1,(n2),dim] (test_smokegun_resim.py:36-55)"""
    dim = x.shape[-1]
    assert u.shape[-1] == dim
    x_adv = torch.empty_like(x)
    v = torch.empty_like(x) if want_v else None
    _lib.get().call('lnst_rk4_advect', ptr(u), dim, _dims3(u.shape, dim), ptr(x), x.shape[0], float(time_step),
                    int(bool(linear)), ptr(x_adv), ptr(v), _s(u))
    return (x_adv, v) if want_v else x_adv


def pressure_loss(d_rec, rest_density, weight, loss, g_d=None):
    """loss += weight*mean(where(d>0, d-rho0, 0)^2); g_d (optional, like d_rec) <- its gradient"""
    _lib.get().call('lnst_pressure_loss', ptr(d_rec), d_rec.numel(), float(rest_density), float(weight), ptr(loss),
                    ptr(g_d), _s(d_rec))


def pressure_reg(d, rest_density, w_mean, g_scale, loss, n_loss, g_d):
    """loss[:n_loss] += w_mean * mean(pr^2); g_d += g_scale * pr, pr = where(d > 0, d - rho0, 0) (styler_3p.py:96-98)"""
    _lib.get().call('lnst_pressure_reg', ptr(d), d.numel(), float(rest_density), float(w_mean), float(g_scale), ptr(loss),
                    int(n_loss), ptr(g_d), _s(d))


def density_reg(var, weight, g_weight, loss, n_loss, grad):
    """styler_base.py:217-223 on clip(var, -1, 1): loss[:n_loss] += weight * (...); grad += g_weight * d(...)/d var"""
    sums = torch.empty(2, dtype=f32, device=var.device)
    _lib.get().call('lnst_density_reg', ptr(var), var.numel(), float(weight), float(g_weight), ptr(sums), ptr(loss),
                    int(n_loss), ptr(grad), _s(var))


def sub_fliph(a, b, out=None):
    """a - flip_H(b) for [D,H,W] volumes"""
    D, H, W = a.shape
    if out is None:
        out = torch.empty_like(a)
    _lib.get().call('lnst_sub_fliph', ptr(a), ptr(b), ptr(out), D, H, W, _s(a))
    return out


# ---- GraphDef loss networks (csrc/graphnet.cu), fp32 NHWC -----------------------------------------------
def same_pad(size, k, stride):
    """TF 'SAME': (output size, pad before) -- the odd cell goes after."""
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return out, total // 2


def conv_out(x_shape, w_shape, stride, padding):
    n, H, W, _ = x_shape
    kh, kw = int(w_shape[0]), int(w_shape[1])
    if padding == 'SAME':
        (OH, pt), (OW, pl) = same_pad(H, kh, stride), same_pad(W, kw, stride)
    else:
        OH, OW, pt, pl = (H - kh) // stride + 1, (W - kw) // stride + 1, 0, 0
    return OH, OW, pt, pl


def conv2d_f32(x, w, bias, stride=1, padding='SAME', relu=False, out=None, ch_off=0):
    """x [n,H,W,Cin], w [kh,kw,Cin,Cout] -> [n,OH,OW,Cout]; with ``out`` [n,OH,OW,Ctot] the result is written into
    channels [ch_off, ch_off+Cout)."""
    n, H, W, cin = x.shape
    kh, kw, _, cout = w.shape
    OH, OW, pt, pl = conv_out(x.shape, w.shape, stride, padding)
    if out is None:
        out = torch.empty(n, OH, OW, cout, dtype=f32, device=x.device)
    ld = out.shape[-1]
    _lib.get().call('lnst_conv2d_f32', ptr(x), ptr(w), ptr(bias), C.c_void_p(out.data_ptr() + 4 * ch_off), n, H, W, cin,
                    cout, kh, kw, stride, pt, pl, OH, OW, ld, int(bool(relu)), _s(x))
    return out


def conv2d_bwd_data_f32(g_y, w, x_shape, stride, padding, g_x, accumulate, ch_off=0, relu_y=None):
    """``relu_y``: post-ReLU output of the convolution, same layout (and channel offset) as g_y -> g_y * (relu_y > 0)"""
    n, H, W, cin = x_shape
    kh, kw, _, cout = w.shape
    OH, OW, pt, pl = conv_out(x_shape, w.shape, stride, padding)
    mask = None if relu_y is None else C.c_void_p(relu_y.data_ptr() + 4 * ch_off)
    assert relu_y is None or relu_y.shape == g_y.shape
    _lib.get().call('lnst_conv2d_bwd_data_f32', C.c_void_p(g_y.data_ptr() + 4 * ch_off), mask, g_y.shape[-1], ptr(w), ptr(g_x),
                    n, H, W, cin, cout, kh, kw, stride, pt, pl, OH, OW, int(bool(accumulate)), _s(g_y))
    return g_x


def relu_fwd(x):
    y = torch.empty_like(x)
    _lib.get().call('lnst_relu_fwd', ptr(x), ptr(y), x.numel(), _s(x))
    return y


def relu_bwd(g_y, y, g_x, accumulate):
    _lib.get().call('lnst_relu_bwd', ptr(g_y), ptr(y), ptr(g_x), y.numel(), int(bool(accumulate)), _s(y))
    return g_x


def maxpool_fwd(x, k, stride, padding='SAME'):
    n, H, W, ch = x.shape
    OH, OW, pt, pl = conv_out(x.shape, (k, k), stride, padding)
    y = torch.empty(n, OH, OW, ch, dtype=f32, device=x.device)
    _lib.get().call('lnst_maxpool_fwd', ptr(x), ptr(y), n, H, W, ch, k, stride, pt, pl, OH, OW, _s(x))
    return y


def maxpool_bwd(g_y, x, k, stride, padding, g_x, accumulate):
    n, H, W, ch = x.shape
    OH, OW, pt, pl = conv_out(x.shape, (k, k), stride, padding)
    _lib.get().call('lnst_maxpool_bwd', ptr(g_y), ptr(x), ptr(g_x), n, H, W, ch, k, stride, pt, pl, OH, OW,
                    int(bool(accumulate)), _s(x))
    return g_x


def avgpool_fwd(x, k, stride, padding='VALID'):
    n, H, W, ch = x.shape
    OH, OW, pt, pl = conv_out(x.shape, (k, k), stride, padding)
    y = torch.empty(n, OH, OW, ch, dtype=f32, device=x.device)
    _lib.get().call('lnst_avgpool_fwd', ptr(x), ptr(y), n, H, W, ch, k, stride, pt, pl, OH, OW, _s(x))
    return y


def avgpool_bwd(g_y, x_shape, k, stride, padding, g_x, accumulate):
    n, H, W, ch = x_shape
    OH, OW, pt, pl = conv_out(x_shape, (k, k), stride, padding)
    _lib.get().call('lnst_avgpool_bwd', ptr(g_y), ptr(g_x), n, H, W, ch, k, stride, pt, pl, OH, OW,
                    int(bool(accumulate)), _s(g_y))
    return g_x


def lrn_fwd(x, depth_radius, bias, alpha, beta):
    y = torch.empty_like(x)
    _lib.get().call('lnst_lrn_fwd', ptr(x), ptr(y), x.numel() // x.shape[-1], x.shape[-1], int(depth_radius), float(bias),
                    float(alpha), float(beta), _s(x))
    return y


def lrn_bwd(g_y, x, depth_radius, bias, alpha, beta, g_x, accumulate):
    _lib.get().call('lnst_lrn_bwd', ptr(g_y), ptr(x), ptr(g_x), x.numel() // x.shape[-1], x.shape[-1], int(depth_radius),
                    float(bias), float(alpha), float(beta), int(bool(accumulate)), _s(x))
    return g_x


def copy_channels(src, src_off, dst, dst_off, ch, accumulate=False):
    """dst[..., dst_off:dst_off+ch] (+)= src[..., src_off:src_off+ch] for NHWC tensors with the same pixel count"""
    pixels = src.numel() // src.shape[-1]
    _lib.get().call('lnst_copy_channels', C.c_void_p(src.data_ptr() + 4 * src_off), src.shape[-1],
                    C.c_void_p(dst.data_ptr() + 4 * dst_off), dst.shape[-1], ch, pixels, int(bool(accumulate)), _s(src))
    return dst


# ---- loss net (tensor-core path, bf16 NHWC) ---------------------------------------------------
bf16 = torch.bfloat16


def conv3x3_bf16_tc(x, w_packed, bias, relu, mask=None, y=None):
    n, H, W, cin = x.shape
    cout = w_packed.shape[1]
    if y is None:
        y = torch.empty(n, H, W, cout, dtype=bf16, device=x.device)
    _lib.get().call('lnst_conv3x3_bf16_tc', ptr(x), ptr(w_packed), ptr(bias), ptr(mask), ptr(y), n, H, W, cin, cout,
                    int(relu), _s(x))
    return y


def conv3x3_mixed(x, w, b, relu, out_bf16, mask=None):
    n, H, W, cin = x.shape
    cout = w.shape[-1]
    y = torch.empty(n, H, W, cout, dtype=bf16 if out_bf16 else f32, device=x.device)
    _lib.get().call('lnst_conv3x3_mixed', ptr(x), int(x.dtype == bf16), ptr(w), ptr(b), ptr(mask), ptr(y),
                    int(out_bf16), n, H, W, cin, cout, int(relu), _s(x))
    return y


def gram_diff_bf16_tc(F, denom, Gs, weight, loss, want_bf16=True):
    """F bf16 [n,h,w,C] -> (G fp32 [n,C,C] = F^T F/denom - Gs, Gd bf16 copy); loss[n] += weight*sum(G^2)."""
    n, h, w, ch = F.shape
    G = torch.empty(n, ch, ch, dtype=f32, device=F.device)
    Gd = torch.empty(n, ch, ch, dtype=bf16, device=F.device) if want_bf16 else None
    _lib.get().call('lnst_gram_diff_bf16_tc', ptr(F), n, h * w, ch, float(denom), ptr(Gs), float(weight), ptr(G),
                    ptr(Gd), ptr(loss), _s(F))
    return G, Gd


def gram_bwd_bf16_tc(F, Gd, coef, addend, relu_mask, g=None):
    n, h, w, ch = F.shape
    if g is None:
        g = torch.empty_like(F)
    _lib.get().call('lnst_gram_bwd_bf16_tc', ptr(F), ptr(Gd), float(coef), ptr(addend), int(relu_mask), ptr(g), n, h,
                    w, ch, _s(F))
    return g


def conv_first_fwd(x, w, b):
    """VGG conv1_1: x fp32 [n,H,W,3] -> bf16 [n,H,W,64] (bias + ReLU)."""
    n, H, W, _ = x.shape
    y = torch.empty(n, H, W, 64, dtype=bf16, device=x.device)
    _lib.get().call('lnst_conv_first_fwd', ptr(x), ptr(w), ptr(b), ptr(y), n, H, W, _s(x))
    return y


def conv_first_bwd(g, wd):
    """data gradient of conv1_1: g bf16 [n,H,W,64] -> fp32 [n,H,W,3]."""
    n, H, W, _ = g.shape
    gx = torch.empty(n, H, W, 3, dtype=f32, device=g.device)
    _lib.get().call('lnst_conv_first_bwd', ptr(g), ptr(wd), ptr(gx), n, H, W, _s(g))
    return gx


def conv_first_bwd_tc(g, wd16):
    """data gradient of conv1_1 on tensor cores: g bf16 [n,H,W,64], wd16 bf16 [9,16,64] -> fp32 [n,H,W,3]."""
    n, H, W, _ = g.shape
    gx = torch.empty(n, H, W, 3, dtype=f32, device=g.device)
    _lib.get().call('lnst_conv_first_bwd_tc', ptr(g), ptr(wd16), ptr(gx), n, H, W, _s(g))
    return gx


def conv_first_fwd_gray(gray, ws, wm, bsum):
    """VGG conv1_1 on a gray render: gray fp32 [n,H,W] -> bf16 [n,H,W,64] (see lnst_conv_first_fwd_gray)."""
    n, H, W = gray.shape
    y = torch.empty(n, H, W, 64, dtype=bf16, device=gray.device)
    _lib.get().call('lnst_conv_first_fwd_gray', ptr(gray), ptr(ws), ptr(wm), ptr(bsum), ptr(y), n, H, W, _s(gray))
    return y


def conv_first_bwd_gray_tc(g, wd16_gray):
    """data gradient of conv1_1 w.r.t. the gray render: g bf16 [n,H,W,64] -> fp32 [n,H,W]."""
    n, H, W, _ = g.shape
    gg = torch.empty(n, H, W, dtype=f32, device=g.device)
    _lib.get().call('lnst_conv_first_bwd_gray_tc', ptr(g), ptr(wd16_gray), ptr(gg), n, H, W, _s(g))
    return gg


def avgpool2_bf16_fwd(x):
    n, H, W, ch = x.shape
    y = torch.empty(n, H // 2, W // 2, ch, dtype=bf16, device=x.device)
    _lib.get().call('lnst_avgpool2_bf16_fwd', ptr(x), ptr(y), n, H, W, ch, _s(x))
    return y


def avgpool2_bf16_bwd(g_y, mask, shape):
    n, H, W, ch = shape
    g_x = torch.empty(n, H, W, ch, dtype=bf16, device=g_y.device)
    _lib.get().call('lnst_avgpool2_bf16_bwd', ptr(g_y), ptr(mask), ptr(g_x), n, H, W, ch, _s(g_y))
    return g_x


def to_bf16(x):
    y = torch.empty(x.shape, dtype=bf16, device=x.device)
    _lib.get().call('lnst_f32_to_bf16', ptr(x), ptr(y), x.numel(), _s(x))
    return y


def to_f32(x):
    y = torch.empty(x.shape, dtype=f32, device=x.device)
    _lib.get().call('lnst_bf16_to_f32', ptr(x), ptr(y), x.numel(), _s(x))
    return y


# ---- bf16x3 ("split") tensor-core path: rows carry [hi | lo] bf16 halves of every fp32 value (csrc/conv_tc.cu ConvShape)
def conv3x3_bf16x3_tc(x, w_packed2, bias, relu, mask=None):
    """x bf16 [n,H,W,2*Cin] split, w_packed2 bf16 [9,Cout,2*Cin] = [Whi | Wlo] -> y bf16 [n,H,W,2*Cout] split"""
    n, H, W, c2 = x.shape
    cout = w_packed2.shape[1]
    y = torch.empty(n, H, W, 2 * cout, dtype=bf16, device=x.device)
    _lib.get().call('lnst_conv3x3_bf16x3_tc', ptr(x), ptr(w_packed2), ptr(bias), ptr(mask), ptr(y), n, H, W, c2 // 2,
                    cout, int(relu), _s(x))
    return y


def conv3x3_pool_bf16x3_tc(x, w_packed2, bias, relu, mask=None):
    """conv3x3_bf16x3_tc plus the 2x2 average pool of its output from the same epilogue -> (y, y_pool [n,H/2,W/2,2*Cout])"""
    n, H, W, c2 = x.shape
    cout = w_packed2.shape[1]
    y = torch.empty(n, H, W, 2 * cout, dtype=bf16, device=x.device)
    yp = torch.empty(n, H // 2, W // 2, 2 * cout, dtype=bf16, device=x.device)
    _lib.get().call('lnst_conv3x3_pool_bf16x3_tc', ptr(x), ptr(w_packed2), ptr(bias), ptr(mask), ptr(y), ptr(yp), n, H, W,
                    c2 // 2, cout, int(relu), _s(x))
    return y, yp


def gram_diff_bf16x3_tc(F, denom, Gs, weight, loss, gd_scale=1.0):
    """F split bf16 [n,h,w,2C] -> (G fp32 [n,C,C] = F^T F/denom - Gs, Gd2 bf16 [n,C,2C] = split(gd_scale * G));
    loss[n] += weight*sum(G^2)"""
    n, h, w, c2 = F.shape
    ch = c2 // 2
    G2 = torch.empty(n, c2, c2, dtype=f32, device=F.device)
    G = torch.empty(n, ch, ch, dtype=f32, device=F.device)
    Gd2 = torch.empty(n, ch, c2, dtype=bf16, device=F.device)
    _lib.get().call('lnst_gram_diff_scaled_bf16x3_tc', ptr(F), n, h * w, ch, float(denom), ptr(Gs), float(weight),
                    float(gd_scale), ptr(G2), ptr(G), ptr(Gd2), ptr(loss), _s(F))
    return G, Gd2


def conv3x3_unpool_bf16x3_tc(x, w_packed2, fine_act):
    """data gradient through a conv whose input is a 2x2 average pool's output, written at the pool's input level under the
    ReLU mask ``fine_act`` [n,2H,2W,2*Cout] -> g_fine (same shape)"""
    n, H, W, c2 = x.shape
    cout = w_packed2.shape[1]
    g_fine = torch.empty_like(fine_act)
    _lib.get().call('lnst_conv3x3_unpool_bf16x3_tc', ptr(x), ptr(w_packed2), ptr(fine_act), ptr(g_fine), n, H, W, c2 // 2, cout,
                    _s(x))
    return g_fine


def conv3x3_gram_bf16x3_tc(x, w_packed2, F, Gd2s):
    """relu_mask(F) * (x (*) w + F x Gd2s): a data gradient and the Gram-loss gradient of the layer it lands on -> split y"""
    n, H, W, c2 = x.shape
    cout = w_packed2.shape[1]
    y = torch.empty(n, H, W, 2 * cout, dtype=bf16, device=x.device)
    _lib.get().call('lnst_conv3x3_gram_bf16x3_tc', ptr(x), ptr(w_packed2), ptr(F), ptr(Gd2s), ptr(y), n, H, W, c2 // 2, cout,
                    _s(x))
    return y


def gram_bwd_bf16x3_tc(F, Gd2, coef, addend, relu_mask, g=None):
    n, h, w, c2 = F.shape
    if g is None:
        g = torch.empty_like(F)
    _lib.get().call('lnst_gram_bwd_bf16x3_tc', ptr(F), ptr(Gd2), float(coef), ptr(addend), int(relu_mask), ptr(g), n, h,
                    w, c2 // 2, _s(F))
    return g


def conv_first_fwd_x3(x, w, b):
    n, H, W, _ = x.shape
    y = torch.empty(n, H, W, 128, dtype=bf16, device=x.device)
    _lib.get().call('lnst_conv_first_fwd_x3', ptr(x), ptr(w), ptr(b), ptr(y), n, H, W, _s(x))
    return y


def conv_first_fwd_gray_x3(gray, ws, wm, bsum):
    n, H, W = gray.shape
    y = torch.empty(n, H, W, 128, dtype=bf16, device=gray.device)
    _lib.get().call('lnst_conv_first_fwd_gray_x3', ptr(gray), ptr(ws), ptr(wm), ptr(bsum), ptr(y), n, H, W, _s(gray))
    return y


def conv_first_bwd_x3_tc(g, wd16_2):
    n, H, W, _ = g.shape
    gx = torch.empty(n, H, W, 3, dtype=f32, device=g.device)
    _lib.get().call('lnst_conv_first_bwd_x3_tc', ptr(g), ptr(wd16_2), ptr(gx), n, H, W, _s(g))
    return gx


def conv_first_bwd_gray_x3_tc(g, wd16_gray2):
    n, H, W, _ = g.shape
    gg = torch.empty(n, H, W, dtype=f32, device=g.device)
    _lib.get().call('lnst_conv_first_bwd_gray_x3_tc', ptr(g), ptr(wd16_gray2), ptr(gg), n, H, W, _s(g))
    return gg


def conv_first_bwd_gray_dot_tc(g, wd16_gray, split, img, dots):
    """conv1_1's gray data gradient plus dots[v] += sum g_gray[v] * img[v] from the same kernel -> g_gray fp32 [n,H,W]"""
    n, H, W, _ = g.shape
    gg = torch.empty(n, H, W, dtype=f32, device=g.device)
    _lib.get().call('lnst_conv_first_bwd_gray_dot_tc', ptr(g), ptr(wd16_gray), ptr(gg), ptr(img), ptr(dots), int(bool(split)),
                    n, H, W, _s(g))
    return gg


def avgpool2_bf16x3_fwd(x):
    n, H, W, c2 = x.shape
    y = torch.empty(n, H // 2, W // 2, c2, dtype=bf16, device=x.device)
    _lib.get().call('lnst_avgpool2_bf16x3_fwd', ptr(x), ptr(y), n, H, W, c2 // 2, _s(x))
    return y


def avgpool2_bf16x3_bwd(g_y, mask, shape):
    n, H, W, c2 = shape
    g_x = torch.empty(n, H, W, c2, dtype=bf16, device=g_y.device)
    _lib.get().call('lnst_avgpool2_bf16x3_bwd', ptr(g_y), ptr(mask), ptr(g_x), n, H, W, c2 // 2, _s(g_y))
    return g_x


def to_split(x):
    """fp32 [..., C] -> bf16 [..., 2C] = [hi | lo]"""
    C_ = x.shape[-1]
    y = torch.empty(tuple(x.shape[:-1]) + (2 * C_,), dtype=bf16, device=x.device)
    _lib.get().call('lnst_f32_to_bf16x3', ptr(x), ptr(y), x.numel() // C_, C_, _s(x))
    return y


def from_split(x):
    """bf16 [..., 2C] split -> fp32 [..., C] = hi + lo"""
    C_ = x.shape[-1] // 2
    y = torch.empty(tuple(x.shape[:-1]) + (C_,), dtype=f32, device=x.device)
    _lib.get().call('lnst_bf16x3_to_f32', ptr(x), ptr(y), x.numel() // (2 * C_), C_, _s(x))
    return y


def conv_first_bwd_gray_direct(g, split, wg):
    """d loss / d gray render from the conv1_1 gradient: g bf16 [n,H,W,64] (or [n,H,W,128] split), wg fp32 [9,64] -> [n,H,W]"""
    n, H, W, _ = g.shape
    gg = torch.empty(n, H, W, dtype=f32, device=g.device)
    _lib.get().call('lnst_conv_first_bwd_gray_direct', ptr(g), int(bool(split)), ptr(wg), ptr(gg), n, H, W, _s(g))
    return gg
