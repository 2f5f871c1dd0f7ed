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
        elif a.switch == 'first_mma':
            lib.call('lnst_set_conv_first_mma', int(not on))      # the MMA form is the non-default arm
        elif a.switch == 'first_col':
            lib.call('lnst_set_conv_first_col', int(on))
        elif a.switch.startswith('lib:'):                   # default library (1) against a variant build (0)
            _lib.set_for_testing(None if on else _lib.Lib(os.path.join(ROOT, a.switch[4:]), 'cuda'))
            ctx.lib = _lib.get()
        elif a.switch == 'fuse_unpool':
            vgg_tc.TensorCoreConvs.fuse_unpool = bool(on)
        elif a.switch.startswith('slab'):                   # slab8 / slab16 against the default 12 planes
            lib.call('lnst_set_raymarch_slab', 12 if on else int(a.switch[4:]))
        elif a.switch == 'fuse_gram':
            vgg_tc.TensorCoreConvs.fuse_gram = bool(on)
        elif a.switch == 'fuse_glue':
            from lnst.styler_3p import Styler
            Styler.fuse_glue = bool(on)
        elif a.switch == 'merge':
            lib.call('lnst_set_raymarch_merge', 2 if on else 1)
        else:
            raise SystemExit('unknown switch')
    import threading
    import time
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    for rnd in range(a.rounds):
        for on in (1, 0):
            set_switch(on)
            samples, stop = [], threading.Event()

            def sample():
                while not stop.is_set():
                    samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
                    time.sleep(0.004)
            th = threading.Thread(target=sample, daemon=True)
            th.start()
            m = bench.measure_step(ctx, 'C3', 'allreduce', 'bf16x3', a.steps, 5, full=False)
            stop.set()
            th.join()
            tail = samples[-max(1, int(a.steps * m['ms_step'] / 4.5)):]          # the samples of the timed region
            clk = sorted(c for c, _ in tail)[len(tail) // 2]
            pw = max(p for _, p in tail)
            print('round %d  %s=%d  %.4f ms/step  (%.1f it/s)  sm %d MHz  power max %.0f W' % (rnd, a.switch, on, m['ms_step'], m['value'], clk, pw), flush=True)
            torch.cuda.empty_cache()
    set_switch(1)


if __name__ == '__main__':
    main()
