"""developer probe: per-shape time of the GraphDef convolutions inside one C5 step (python tools/dev/c5_probe.py)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'neural-flow-style_b200'), os.path.join(ROOT, 'tests')]
import torch
import bench
from lnst import synth
from lnst.styler_3p import Styler, _Adam

wl = sys.argv[1] if len(sys.argv) > 1 else 'C5'
dev = torch.device('cuda:0')
cfg = bench.make_cfg(wl, 'allreduce', 'bf16x3')
p, r, sty = bench.make_scene(wl)
st = Styler(cfg, weights=synth.vgg_weights(), device=dev, content_weights=bench.content_nodes(wl))
st.style_img = sty
res = [bench.WORKLOADS[wl]['res']] * 3
grams = st._style_feature(sty, res[1:])
st.num_frames = 1
frames, _ = st.upload({'p': p, 'r': r})
ws = st._workspace(res, frames)
fr = frames[0]
g_opt = torch.zeros(fr['p'].shape[0], 2, device=dev)
adam = _Adam()
st.fuse_apply = True
from lnst import _lib
lib = _lib.get()
for _ in range(2):
    st.frame_step(fr, g_opt, adam, ws, grams, cfg.lr)
torch.cuda.synchronize()
orig = lib.call
stats = {}

def prof(name, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(name, *a); e1.record()
    v = lambda x: x.value if hasattr(x, 'value') else x
    key = name
    if name in ('lnst_conv2d_bwd_data_f32', 'lnst_conv2d_f32'):
        o = 4 if name == 'lnst_conv2d_f32' else 5
        key = name + str(tuple(v(x) for x in a[o:o + 8]))
    stats.setdefault(key, []).append((e0, e1))
lib.call = prof
st.frame_step(fr, g_opt, adam, ws, grams, cfg.lr)
torch.cuda.synchronize()
lib.call = orig
rows = sorted(((sum(a.elapsed_time(b) for a, b in v), len(v), k) for k, v in stats.items()), reverse=True)
tot = sum(x[0] for x in rows)
print('eager step %.2f ms' % tot)
for ms, n, k in rows[:14]:
    print('%8.3f ms  x%-3d %s' % (ms, n, k))
