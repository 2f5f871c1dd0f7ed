#!/usr/bin/env python
"""A/B of the graph-replayed C3 step with one switch flipped (developer tool; GPU only).

    python tools/dev/ab_step.py fuse_pool | gram_split3 | first_col [--steps 40]
"""
import argparse
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in (ROOT, os.path.join(ROOT, 'neural-flow-style_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, _p)
import torch  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('switch')
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--rounds', type=int, default=3)
    a = ap.parse_args()
    from lnst import _lib, vgg_tc
    ctx = bench.Ctx()
    lib = _lib.get()

    def set_switch(on):
        if a.switch == 'fuse_pool':
            vgg_tc.TensorCoreConvs.fuse_pool = bool(on)
        elif a.switch == 'gram_split3':
            lib.call('lnst_set_gram_split3', int(on))
        elif a.switch == 'first_col':
            lib.call('lnst_set_conv_first_col', int(on))
        elif a.switch == 'merge':
            lib.call('lnst_set_raymarch_merge', 2 if on else 1)
        else:
            raise SystemExit('unknown switch')
    for rnd in range(a.rounds):
        for on in (1, 0):
            set_switch(on)
            m = bench.measure_step(ctx, 'C3', 'allreduce', 'bf16x3', a.steps, 5, full=False)
            print('round %d  %s=%d  %.4f ms/step  (%.1f it/s)' % (rnd, a.switch, on, m['ms_step'], m['value']), flush=True)
            torch.cuda.empty_cache()
    set_switch(1)


if __name__ == '__main__':
    main()
