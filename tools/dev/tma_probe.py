"""developer probe: run one TMA-tiled kernel in isolation (under compute-sanitizer) -- python tools/tma_probe.py smooth|rm"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'neural-flow-style_b200'))
import torch
from lnst import ops, _lib
dev = torch.device('cuda:0')
what = sys.argv[1]
D = H = W = int(sys.argv[2]) if len(sys.argv) > 2 else 24
d = torch.rand(D, H, W, device=dev)
if what == 'smooth':
    o = ops.smooth3_relu_fwd(d, torch.zeros_like(d), 3, None)
    torch.cuda.synchronize(); print('fwd ok', float(o.sum()))
    ops.USE_TMA = False
    o2 = ops.smooth3_relu_fwd(d, torch.zeros_like(d), 3, None)
    print('max diff', float((o - o2).abs().max()))
    ops.USE_TMA = True
    g = ops.smooth3_relu_bwd(d, o, torch.zeros_like(d), 3, None)
    torch.cuda.synchronize(); print('bwd ok', float(g.sum()))
else:
    rot = torch.eye(3).reshape(1, 9).to(dev)
    img = torch.empty(1, H, W, device=dev); st = torch.empty(1, H, W, device=dev)
    ops.raymarch_fwd(d, rot, 0.01, False, img, st)
    torch.cuda.synchronize(); print('rm ok', float(img.sum()))
