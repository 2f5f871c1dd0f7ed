"""developer probe: lane utilisation of the ray-march kernels for different pixel -> warp mappings (C3 exact intervals)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'neural-flow-style_b200'), os.path.join(ROOT, 'tests')]
import torch
import bench
from lnst import synth, ops
from lnst.styler_3p import Styler
wl = 'C3'
dev = torch.device('cuda:0')
cfg = bench.make_cfg(wl, 'allreduce', 'bf16')
p, r, sty = bench.make_scene(wl)
st = Styler(cfg, weights=synth.vgg_weights(), device=dev)
st.num_frames = 1
frames, _ = st.upload({'p': p, 'r': r})
res = [200] * 3
ws = st._workspace(res, frames)
rot = st._rot_all
for name, iv in (('bricks', ops.ray_intervals(rot, res, ws['box'], ws['bricks'])), ('exact', ops.ray_intervals_exact(rot, res, ws['box'], ws['touch']))):
    lo, hi = iv[..., 0].long(), iv[..., 1].long()
    ln = (hi - lo + 1).clamp(min=0)
    live = int(ln.sum())
    nv, H, W = lo.shape
    big = 10 ** 6
    def util(th, tw):
        # warps = th x tw pixel patches (th*tw = 32)
        Hp, Wp = (H + th - 1) // th * th, (W + tw - 1) // tw * tw
        L = torch.full((nv, Hp, Wp), big, device=dev); Hh = torch.full((nv, Hp, Wp), -big, device=dev)
        e = ln > 0
        L[:, :H, :W] = torch.where(e, lo, torch.full_like(lo, big)); Hh[:, :H, :W] = torch.where(e, hi, torch.full_like(hi, -big))
        L = L.reshape(nv, Hp // th, th, Wp // tw, tw).amin(dim=(2, 4)); Hh = Hh.reshape(nv, Hp // th, th, Wp // tw, tw).amax(dim=(2, 4))
        un = (Hh - L + 1).clamp(min=0)
        un = torch.where(un > big, torch.zeros_like(un), un)
        return int(un.sum()) * 32
    print(name, 'live samples', live)
    for th, tw in ((1, 32), (2, 16), (4, 8), (8, 4)):
        slots = util(th, tw)
        print('   warp %dx%-2d lane-slots %d utilisation %.3f' % (th, tw, slots, live / slots))
