#!/usr/bin/env python
"""Per-kernel micro-benchmark at a bench workload's shapes (developer tool; GPU only).

    python tools/kbench.py [--workload C3] [--reps 10] [--only raymarch]

Builds the scene exactly like bench.py, runs one loss+gradient evaluation to obtain realistic
intermediate tensors, then times each hot entry point alone with CUDA events (mean of --reps
back-to-back launches after 2 warm-ups) and prints algorithmic GB/s next
to the measured HBM peak.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'neural-flow-style_b200'), os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

import bench  # noqa: E402


def timeit(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='C3')
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--only', default='')
    ap.add_argument('--rm-slab', dest='rm_slab', type=int, default=0)
    args = ap.parse_args()
    from lnst import _lib, ops, synth
    from lnst.styler_3p import Styler
    dev = torch.device('cuda:0')
    lib = _lib.get()
    if args.rm_slab:
        lib.call('lnst_set_raymarch_slab', args.rm_slab)
    wl = bench.WORKLOADS[args.workload]
    cfg = bench.make_cfg(args.workload, 'allreduce', 'bf16')
    p, r, sty = bench.make_scene(args.workload)
    st = Styler(cfg, weights=synth.vgg_weights(), device=dev)
    st.style_img = sty
    res = [wl['res']] * 3
    grams = st._style_feature(sty, res[1:])
    st.num_frames = 1
    frames, _ = st.upload({'p': p, 'r': r})
    ws = st._workspace(res, frames)
    box = ws['box']
    Vb = ws.get('box_cells', res[0] ** 3)
    fr = frames[0]
    var = torch.zeros(fr['p'].shape[0], 2, device=dev)
    rot = st._rot_all if wl['rotate'] else None
    loss, grad = st.loss_and_grad(fr, var, ws, rot, grams)
    torch.cuda.synchronize()
    D, H, W = res
    V, P, N = D * H * W, H * W, fr['p'].shape[0]
    nv = 1 if rot is None else rot.shape[0]
    ds = ws['ds']
    img = torch.empty(nv, H, W, device=dev)
    stot = torch.empty(nv, H, W, device=dev)
    ops.raymarch_fwd(ds, rot, st.transmit, False, img, stot)
    g_img = torch.randn(nv, H, W, device=dev)
    g_ds = torch.zeros(D, H, W, device=dev)
    key = (fr['id'], tuple(res))
    wmap = st._wmap(fr, res, ws['grid'])
    hs = st._supports()
    g_d = torch.randn(D, H, W, device=dev)
    gvar = torch.empty_like(var)
    out = {}

    def add(name, fn, nbytes):
        if args.only and args.only not in name:
            return
        ms = timeit(fn, args.reps)
        out[name] = {'ms': round(ms, 4), 'alg_GBps': round(nbytes / ms / 1e6, 1), 'alg_MB': round(nbytes / 1e6, 1)}
        print('%-28s %8.4f ms  %8.1f GB/s algorithmic (%.1f MB)' % (name, ms, nbytes / ms / 1e6, nbytes / 1e6), flush=True)

    print('active box cells %d of %d (%.3f)' % (Vb, V, Vb / V))
    add('raymarch_fwd(full)', lambda: ops.raymarch_fwd(ds, rot, st.transmit, False, img, stot), nv * (4 * V + 8 * P))
    add('raymarch_bwd(full)', lambda: ops.raymarch_bwd(ds, rot, st.transmit, False, stot, g_img, g_ds), nv * (8 * V + 8 * P))
    add('raymarch_fwd(box only)', lambda: ops.raymarch_fwd(ds, rot, st.transmit, False, img, stot, box), nv * (4 * Vb + 8 * P))
    bricks = ops.ray_intervals(rot, ds.shape, box, ws['bricks']) if rot is not None else None
    add('ray_intervals', lambda: ops.ray_intervals(rot, ds.shape, box, ws['bricks']), nv * 8 * P)
    add('raymarch_fwd', lambda: ops.raymarch_fwd(ds, rot, st.transmit, False, img, stot, box, bricks), nv * (4 * Vb + 8 * P))
    exact = ops.ray_intervals_exact(rot, ds.shape, box, ws['touch']) if rot is not None and ws.get('touch') is not None else None
    if exact is not None:
        live = lambda t: int((t[..., 1] - t[..., 0] + 1).clamp(min=0).sum())
        print('live samples: bricks %d, exact %d' % (live(bricks), live(exact)))
        add('ray_intervals_exact', lambda: ops.ray_intervals_exact(rot, ds.shape, box, ws['touch']), nv * 8 * P)
        add('raymarch_fwd(exact iv)', lambda: ops.raymarch_fwd(ds, rot, st.transmit, False, img, stot, box, exact), nv * (4 * Vb + 8 * P))
        add('raymarch_bwd(exact iv)', lambda: ops.raymarch_bwd(ds, rot, st.transmit, False, stot, g_img, g_ds, box, exact), nv * (8 * Vb + 8 * P))
    if rot is not None:
        lib.call('lnst_set_raymarch_merge', 0)
        add('raymarch_bwd(merge=0)', lambda: ops.raymarch_bwd(ds, rot, st.transmit, False, stot, g_img, g_ds, box),
            nv * (8 * Vb + 8 * P))
        lib.call('lnst_set_raymarch_merge', 2)
    add('raymarch_bwd(box only)', lambda: ops.raymarch_bwd(ds, rot, st.transmit, False, stot, g_img, g_ds, box), nv * (8 * Vb + 8 * P))
    add('raymarch_bwd', lambda: ops.raymarch_bwd(ds, rot, st.transmit, False, stot, g_img, g_ds, box, bricks), nv * (8 * Vb + 8 * P))
    add('splat_wavg_fwd', lambda: ops.splat_wavg_fwd(fr['p'], fr['r'], var, ws['grid'], hs, wmap, ws['num'], ws['d'], box),
        N * (12 + 16) + 4 * Vb)
    lists = ops.cell_lists(fr['p'], ws['grid'])
    if lists is not None and lib.has_tma:
        dg = torch.zeros_like(ws['d'])
        add('splat_wavg_fwd(gather+TMA store)', lambda: ops.splat_wavg_fwd_gather(lists, fr['r'], var, ws['grid'], hs, dg, box),
            N * (12 + 16) + 4 * Vb)
        dref = ops.splat_wavg_fwd(fr['p'], fr['r'], var, ws['grid'], hs, wmap, ws['num'], torch.zeros_like(ws['d']), box)
        print('gather vs scatter max rel diff %.2e' % float((dg - dref).abs().max() / dref.abs().max()))
    add('splat_wavg_bwd(generic)', lambda: ops.splat_wavg_bwd(fr['p'], var, ws['grid'], hs, wmap, g_d, gvar), N * (12 + 16) + 4 * Vb)
    coef = ops.splat_wavg_coef(wmap)
    add('splat_wavg_bwd', lambda: ops.splat_wavg_bwd_coef(fr['p'], var, ws['grid'], hs, coef, g_d, gvar), N * (12 + 16) + 4 * Vb)
    add('smooth3_relu_fwd', lambda: ops.smooth3_relu_fwd(ws['d'], ws['ds'], st.k, box), 8 * Vb)
    add('smooth3_relu_bwd', lambda: ops.smooth3_relu_bwd(g_ds, ds, ws['g_d'], st.k, box), 12 * Vb)
    # semi-Lagrangian advect microbenchmark: 128^3 scalar density, 3-channel velocity U(+-2 cells)
    ad = torch.rand(128, 128, 128, 1, device=dev)
    av = ((torch.rand(128, 128, 128, 3, device=dev) * 4 - 2) * (2.0 / 127)).contiguous()
    add('advect 128^3 (TMA-staged tiles)', lambda: ops.advect(ad, av), 4 * 128 ** 3 * 5)
    ops.USE_TMA = False
    add('advect 128^3 (gather kernel)', lambda: ops.advect(ad, av), 4 * 128 ** 3 * 5)
    ops.USE_TMA = True
    x = torch.randn(nv, H, W, 3, device=dev)
    d_img = torch.empty_like(x)

    def lossnet():
        l = torch.zeros(nv, device=dev)
        return st.image_loss_and_grad(x, d_img, grams, l)
    lib.call('lnst_set_conv_halo', 0)
    st.net.tc.first_bwd_tc = False
    add('lossnet fwd+bwd (per-tap convs)', lossnet, 0)
    lib.call('lnst_set_conv_halo', 2)
    st.net.tc.first_bwd_tc = True
    add('lossnet fwd+bwd (all views)', lossnet, 0)
    add('zero g_ds', lambda: ops.fill_box(g_ds, box, 0.0), 4 * Vb)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
