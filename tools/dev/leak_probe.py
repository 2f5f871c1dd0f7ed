#!/usr/bin/env python
"""Which device tensors outlive a Styler.run (developer tool; GPU only)."""
import gc
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in (ROOT, os.path.join(ROOT, 'neural-flow-style_b200'), os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import torch  # noqa: E402

import bench  # noqa: E402
from lnst import synth  # noqa: E402
from lnst.styler_3p import Styler  # noqa: E402

dev = torch.device('cuda:0')
p, r, sty = bench.make_scene('C3')


def live():
    out = {}
    for o in gc.get_objects():
        try:
            if isinstance(o, torch.Tensor) and o.is_cuda and o.numel() * o.element_size() >= (1 << 20):
                out[id(o)] = o
        except Exception:
            pass
    return out


def stats(tag):
    s = torch.cuda.memory_stats()
    print(tag, {k: s.get(k, 0) >> 20 for k in ('allocated_bytes.all.current', 'reserved_bytes.all.current')}, flush=True)


for call in range(3):
    cfg = bench.make_cfg('C3', 'allreduce', 'bf16x3')
    cfg.iter = 20
    st = Styler(cfg, weights=synth.vgg_weights(), device=dev)
    st.style_img = sty
    out = st.run({'p': p, 'r': r})
    torch.cuda.synchronize()
    stats('call %d before del' % call)
    del st, out
    stats('call %d after del ' % call)
    n = gc.collect()
    stats('call %d after gc(%d)' % (call, n))
    lv = live()
    print('  live >=1MiB python-visible tensors:', sorted((tuple(t.shape), str(t.dtype)) for t in lv.values()), flush=True)
    for t in list(lv.values())[:6]:
        refs = [type(x).__name__ + (':' + ','.join(list(x.keys())[:6]) if isinstance(x, dict) else '') for x in gc.get_referrers(t)]
        print('   ', tuple(t.shape), refs[:5], flush=True)
print(torch.cuda.memory_summary(abbreviated=True)[:1500])
