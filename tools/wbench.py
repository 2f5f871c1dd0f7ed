#!/usr/bin/env python
"""Micro-benchmark of the widening-row components on a B200 (developer tool; GPU only):

    python tools/wbench.py [--reps 10] [--only g2p,rk4,resim,inception,c5]

  g2p / rk4   lnst_g2p (cubic, 3 channels) and lnst_rk4_advect, N = 2^20 particles in a 200x300x200 velocity grid
  resim       SimG2P.optimize, one frame of the resimulation driver's configuration (test_smokegun_resim.py:336-372)
  inception   GraphNet forward + data-gradient, full-width synthetic inception5h graph, 300x300 input (the smokegun
              driver's net input), content layer mixed4d_3x3_bottleneck_pre_relu
  c5          one allreduce-mode iteration of BASELINE configs[4] scaled to one GPU: 256^3, 9 views, inception
              semantic (mixed4d_3x3_bottleneck_pre_relu ch 139) + VGG-19 style, N = 2^21

One JSON line per item: mean ms over --reps launches after 2 warm-ups (CUDA events), algorithmic GB/s or TFLOP/s.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'neural-flow-style_b200'), os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


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


def conv_flops(nodes, acts):
    """2*pixels*k*k*Cin*Cout over the Conv2D nodes that ran"""
    by = {n.name: n for n in nodes}
    fl = 0
    for name, a in acts.items():
        n = by.get(name)
        if n is None or n.op != 'BiasAdd':
            continue
        conv = by.get(n.inputs[0])
        if conv is None or conv.op != 'Conv2D':
            continue
        w = by[conv.inputs[1]].attr['value']
        fl += 2 * a.shape[0] * a.shape[1] * a.shape[2] * int(np.prod(w.shape))
    return fl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--only', default='g2p,rk4,resim,inception,c5')
    args = ap.parse_args()
    only = set(args.only.split(','))
    from lnst import _lib, ops, synth
    dev = torch.device('cuda:0')
    _lib.get()
    rng = np.random.RandomState(0)

    if only & {'g2p', 'rk4'}:
        D, H, W, N = 200, 300, 200, 1 << 20
        u = torch.tensor(rng.uniform(-0.01, 0.01, (D, H, W, 3)).astype(np.float32)).to(dev)
        cells = np.sort(rng.randint(0, D * H * W, N))                                  # cell-sorted like the drivers' sets
        z, y, x = np.unravel_index(cells, (D, H, W))
        p = torch.tensor(((np.stack([z, y, x], -1) + rng.rand(N, 3)) / [D, H, W]).astype(np.float32)).to(dev)
        if 'g2p' in only:
            ms = timeit(lambda: ops.g2p(u, p), args.reps)
            print(json.dumps({'item': 'lnst_g2p cubic C=3', 'ms': ms, 'GB/s_algorithmic': (N * 24 + u.numel() * 4) / ms / 1e6}))
            ms = timeit(lambda: ops.g2p(u, p, linear=True), args.reps)
            print(json.dumps({'item': 'lnst_g2p linear C=3', 'ms': ms, 'GB/s_algorithmic': (N * 24 + u.numel() * 4) / ms / 1e6}))
        if 'rk4' in only:
            ms = timeit(lambda: ops.rk4_advect(u, p, 0.5), args.reps)
            print(json.dumps({'item': 'lnst_rk4_advect', 'ms': ms, 'GB/s_algorithmic': (N * 24 + u.numel() * 4) / ms / 1e6}))

    if 'resim' in only:
        import argparse as _a
        from lnst.resim import SimG2P
        c = _a.Namespace(resolution=[200, 300, 200], domain=[200, 300, 200], disc=1, radius=0.5, nsize=1, support=4,
                         rest_density=1000, threshold=0.01, lr=0.0005, iter=20, octave_n=2, octave_scale=2)
        zz, yy, xx = np.meshgrid(*[(np.arange(n) + .5) / n for n in c.resolution], indexing='ij')
        d = np.exp(-(((zz - .5) / .15) ** 2 + ((yy - .75) / .12) ** 2 + ((xx - .2) / .12) ** 2)).astype(np.float32)
        d *= d > 0.05
        u = np.zeros(d.shape + (3,), np.float32)
        u[..., 1] = -0.01
        sim = SimG2P(c, device=dev)
        p0, pid = sim.sample(d, disc=1, threshold=0)
        import time
        sim.optimize(p0, pid, d, u)
        torch.cuda.synchronize()
        t0 = time.time()
        res = sim.optimize(p0, pid, d, u)
        torch.cuda.synchronize()
        print(json.dumps({'item': 'SimG2P.optimize (1 frame, 20 Adam iterations, 2 octaves)', 'particles': int(p0.shape[0]),
                          'wall_ms': (time.time() - t0) * 1e3, 'loss': [res['l'][0], res['l'][-1]]}))

    if 'inception' in only:
        from lnst.graphnet import GraphNet
        nodes = synth.inception5h_nodes(upto='mixed4d')
        net = GraphNet(nodes, dev)
        x = torch.tensor(rng.uniform(-120, 130, (1, 300, 300, 3)).astype(np.float32)).to(dev)
        layer = 'mixed4d_3x3_bottleneck_pre_relu'
        acts = net.forward(x, [layer])
        fl = conv_flops(nodes, acts)
        ms_f = timeit(lambda: net.forward(x, [layer]), args.reps)
        g = torch.ones_like(acts[layer])
        ms_b = timeit(lambda: net.backward(x, acts, [layer], lambda n, gg: g, {layer}), args.reps)
        print(json.dumps({'item': 'GraphNet inception5h -> mixed4d_3x3_bottleneck_pre_relu @300x300', 'fwd_ms': ms_f,
                          'bwd_ms': ms_b, 'fwd_GFLOP': fl / 1e9, 'fwd_TFLOP/s': fl / ms_f / 1e9, 'bwd_TFLOP/s': fl / ms_b / 1e9}))

    if 'c5' in only:
        from helpers import smoke_cfg
        from lnst.styler_3p import Styler
        res, N = 256, 1 << 21
        cfg = smoke_cfg(res=res, iter=1, rotate=True, n_views=9, view_mode='allreduce', conv_math='bf16', transmit=0.01,
                        style_layer=['conv2_1', 'conv3_1'], w_style_layer=[0.5, 0.5],
                        content_network='tensorflow_inception_graph.pb', w_content=1.0,
                        content_layer='mixed4d_3x3_bottleneck_pre_relu', content_channel=139)
        p, r = synth.smoke_particles(N, 2, seed=123)
        st = Styler(cfg, weights=synth.vgg_weights(), content_weights=synth.inception5h_nodes(upto='mixed4d'), device=dev)
        st.style_img = synth.style_image(res, res)
        st.cuda_graphs = True
        cfg.iter = 3
        st.iter = 3
        st.run({'p': p, 'r': r})                                   # eager + capture + one replay
        torch.cuda.synchronize()
        import time
        st.iter = 13
        t0 = time.time()
        st.run({'p': p, 'r': r})
        torch.cuda.synchronize()
        print(json.dumps({'item': 'C5 (256^3, 9 views, inception semantic + VGG-19 style), whole run() of 13 iterations '
                                  'incl. setup', 'wall_ms_per_iteration_upper_bound': (time.time() - t0) * 1e3 / 13}))


if __name__ == '__main__':
    main()
