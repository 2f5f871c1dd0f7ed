#!/usr/bin/env python
"""Where the wall clock of ``Styler(config).run(params)`` goes (developer tool; GPU only).

    python tools/dev/run_phases.py [--workload C3] [--iters 20] [--calls 3]

Wraps the phases of ``run`` (upload, workspace, style target, every step-runner call, inference) with
host timers bracketed by device synchronisation and prints one JSON object per call.  The synchronisations
serialise host and device, so the sum is an upper bound of the un-instrumented wall clock printed next to it.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in (ROOT, os.path.join(ROOT, 'neural-flow-style_b200'), os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='C3')
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--calls', type=int, default=3)
    args = ap.parse_args()
    from lnst import styler_3p, synth
    from lnst.styler_3p import Styler, StepRunner
    dev = torch.device('cuda:0')
    p, r, sty = bench.make_scene(args.workload)
    acc = {}

    def wrap(cls, name, label=None):
        fn = getattr(cls, name)
        lab = label or name

        def timed(*a, **k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn(*a, **k)
            torch.cuda.synchronize()
            acc.setdefault(lab, []).append(time.perf_counter() - t0)
            return out
        setattr(cls, name, timed)
        return fn

    originals = [(Styler, n, wrap(Styler, n)) for n in ('upload', '_workspace', '_style_feature', 'infer')]
    originals.append((StepRunner, '__call__', wrap(StepRunner, '__call__', 'step')))

    def wrap_plain(cls, name, lab):                           # no synchronisation: legal inside a stream capture
        fn = getattr(cls, name)

        def timed(*a, **k):
            n0 = torch.cuda.memory_stats().get('num_device_alloc', 0)
            f0 = torch.cuda.memory_stats().get('num_device_free', 0)
            t0 = time.perf_counter()
            out = fn(*a, **k)
            acc.setdefault(lab, []).append(time.perf_counter() - t0)
            acc.setdefault(lab + '_cudaMalloc/Free', []).append(
                (torch.cuda.memory_stats().get('num_device_alloc', 0) - n0, torch.cuda.memory_stats().get('num_device_free', 0) - f0))
            return out
        setattr(cls, name, timed)
        originals.append((cls, name, fn))
    wrap_plain(torch.cuda.CUDAGraph, 'capture_begin', 'capture_begin')
    wrap_plain(torch.cuda.CUDAGraph, 'capture_end', 'capture_end')
    wrap_plain(Styler, 'frame_step', 'frame_step')

    def mk():
        cfg = bench.make_cfg(args.workload, 'allreduce', 'bf16x3')
        cfg.iter = args.iters
        st = Styler(cfg, weights=synth.vgg_weights(), device=dev)
        st.style_img = sty
        return st
    rows = []
    for call in range(args.calls):
        acc.clear()
        t0 = time.perf_counter()
        st = mk()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        out = st.run({'p': p, 'r': r})
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        steps = acc.pop('step', [])
        row = {'call': call, 'construct_ms': 1e3 * (t1 - t0), 'run_ms(instrumented)': 1e3 * (t2 - t1)}
        for k, v in acc.items():
            if k.endswith('Free'):
                row[k] = v[:3]
            else:
                row[k + '_ms'] = [round(1e3 * x, 2) for x in v][:4]
        row['mem'] = {k: torch.cuda.memory_stats().get(k, 0) for k in ('num_device_alloc', 'num_device_free', 'reserved_bytes.all.current', 'allocated_bytes.all.current', 'active_bytes.all.current')}
        row['step_ms'] = {'first(eager)': round(1e3 * steps[0], 2), 'second(capture)': round(1e3 * steps[1], 2) if len(steps) > 1 else None,
                          'rest_sum': round(1e3 * sum(steps[2:]), 2), 'n': len(steps)}
        known = sum(sum(v) for k, v in acc.items() if k in ('upload', '_workspace', '_style_feature', 'infer')) + sum(steps)
        row['other_ms (D2H, result assembly, glue)'] = round(1e3 * (t2 - t1 - known), 2)
        rows.append(row)
        print(json.dumps(row), flush=True)
        del st, out
    for cls, n, fn in originals:
        setattr(cls, n, fn)
    walls = []
    for call in range(args.calls):
        st = mk()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        out = st.run({'p': p, 'r': r})
        torch.cuda.synchronize()
        walls.append(round(1e3 * (time.perf_counter() - t1), 2))
        del st, out
    print(json.dumps({'uninstrumented_run_ms': walls}), flush=True)


if __name__ == '__main__':
    main()
