#!/usr/bin/env python
"""Style-optimisation iterations/sec on BASELINE.json's headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the reference's ``for step in range(iter)`` body (``styler_3p.py:301``)
for one frame: 9 rotated views of a 200^3 smoke volume (N = 2^20 particles, 2 kernels,
VGG-19 style loss on conv2_1 + conv3_1), mean view gradient, one Adam update
(``view_mode='allreduce'``; views are sharded over the ranks and the particle gradient is
all-reduced over NCCL).  Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement".

``value`` is measured on the fp32-tolerance tensor-core path (``conv_math='bf16x3'``); the same line carries
``value_bf16`` (single-pass bf16 operands, outside the fp32 tolerance), ``value_sequential`` (the reference-exact view
mode: one Adam step per view), ``e2e_run`` (wall clock of ``Styler(config).run(params)`` from host arrays to the result
dict for the reference's 20-iteration budget) and, at N = 1, ``configs``: one entry per other BASELINE.json
configuration (C1 2-D colour, C2 128^3 single view, C4 60-frame sequence, C5 256^3 multi-net).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'neural-flow-style_b200'), os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (res, n_particles, rotate, n_views)
    'C3': dict(res=200, n=1 << 20, rotate=True, n_views=9,
               desc='smokegun-like 200^3, N=2^20 particles x 2 kernels, 9 rotated views, VGG-19 conv2_1+conv3_1'),
    'C2': dict(res=128, n=1 << 18, rotate=False, n_views=1,
               desc='smokegun-like 128^3, N=2^18 particles x 2 kernels, 1 view, VGG-19 conv2_1+conv3_1'),
    'tiny': dict(res=32, n=1 << 13, rotate=True, n_views=9, desc='debug 32^3'),
    # BASELINE.json configs[4] on one GPU's share: multi-net loss -- semantic term on an inception5h GraphDef (seeded
    # synthetic weights, full channel widths), style term on VGG-19.  Not a default bench line (a parity-test config);
    # `--workload C5` measures it, tools/wbench.py times its pieces.
    'C5': dict(res=256, n=1 << 21, rotate=True, n_views=9,
               content=dict(layer='mixed4d_3x3_bottleneck_pre_relu', channel=139, upto='mixed4d', width_div=1),
               desc='smokegun-like 256^3, N=2^21 particles x 2 kernels, 9 rotated views, inception5h '
                    'mixed4d_3x3_bottleneck_pre_relu ch 139 semantic + VGG-19 conv2_1+conv3_1 style (multi-net loss)'),
    'C5tiny': dict(res=24, n=1 << 12, rotate=True, n_views=3,
                   content=dict(layer='mixed3b_3x3_bottleneck_pre_relu', channel=5, upto='mixed3b', width_div=8),
                   desc='debug 24^3 multi-net'),
}
ACTIVE_CELLS = 0   # cells of the workload's active box (set by run_engine)
KERNELS_PER_CALL = {'lnst_splat_wavg_fwd_box': 2, 'lnst_adam_step_dev': 2, 'lnst_image_max': 2, 'lnst_normalize_bwd': 2, 'lnst_gram_diff': 2,
                    'lnst_gram_diff_bf16_tc': 2, 'lnst_gram_diff_bf16x3_tc': 2, 'lnst_density_reg': 2}
TENSOR_BOUND = ('lnst_conv3x3_f32', 'lnst_conv2d_f32', 'lnst_conv2d_bwd_data_f32', 'lnst_conv3x3_bf16_tc', 'lnst_gram_diff', 'lnst_gram_bwd', 'lnst_gram_diff_bf16_tc',
                'lnst_gram_bwd_bf16_tc', 'lnst_conv3x3_bf16x3_tc', 'lnst_gram_diff_bf16x3_tc', 'lnst_gram_bwd_bf16x3_tc')
# bf16x3 entry points execute three bf16 MMA passes per algorithmic (fp32) multiply-add
MMA_PASSES = {'lnst_conv3x3_bf16x3_tc': 3, 'lnst_gram_bwd_bf16x3_tc': 3,
              'lnst_gram_diff_bf16x3_tc': 2.6}   # hi^T hi + hi^T lo + lo^T hi on the tiles with row block <= column block: 3 (C = 128) .. 2.25 (C = 256)
# entry points that are the same kernel with one more output: timed and counted under the base name
ALIASES = {'lnst_conv3x3_pool_bf16x3_tc': 'lnst_conv3x3_bf16x3_tc', 'lnst_raymarch_fwd_max_tma': 'lnst_raymarch_fwd_tma',
           'lnst_raymarch_fwd_max_box': 'lnst_raymarch_fwd_box', 'lnst_raymarch_bwd_norm_box': 'lnst_raymarch_bwd_box',
           'lnst_conv_first_bwd_gray_dot_tc': 'lnst_conv_first_bwd_gray_x3_tc',
           'lnst_conv3x3_gram_bf16x3_tc': 'lnst_conv3x3_bf16x3_tc', 'lnst_gram_diff_scaled_bf16x3_tc': 'lnst_gram_diff_bf16x3_tc'}


def make_cfg(wl, view_mode, conv_math):
    from helpers import smoke_cfg
    w = WORKLOADS[wl]
    extra = {}
    if w.get('content'):
        extra = dict(content_network='tensorflow_inception_graph.pb', w_content=1.0, content_layer=w['content']['layer'],
                     content_channel=w['content']['channel'])
    return smoke_cfg(res=w['res'], iter=1, rotate=w['rotate'], n_views=w['n_views'], view_mode=view_mode,
                     conv_math=conv_math, transmit=0.01, style_layer=['conv2_1', 'conv3_1'], w_style_layer=[0.5, 0.5],
                     **extra)


def content_nodes(wl):
    """The second (GraphDef) loss network of a multi-net workload, else None."""
    from lnst import synth
    c = WORKLOADS[wl].get('content')
    return synth.inception5h_nodes(width_div=c['width_div'], upto=c['upto']) if c else None


def make_scene(wl):
    from lnst import synth
    w = WORKLOADS[wl]
    p, r = synth.smoke_particles(w['n'], 2, seed=123)
    sty = synth.style_image(w['res'], w['res'])
    return p, r, sty


def algorithmic_units(name, a, nk=2):
    """(bytes, flops) one launch of entry point ``name`` must move/compute (fp32; compulsory inputs
    read once, outputs written once -- SURVEY.md section 8d; DESIGN.md 'Roofline model')."""
    def v(x):
        return x.value if hasattr(x, 'value') else x
    def box_cells(b, V):
        """cells of an LnstBox argument (None = whole volume): the compulsory, non-zero part of a volume"""
        if b is None:
            return V
        b = b._obj
        return (b.hi[0] - b.lo[0] + 1) * (b.hi[1] - b.lo[1] + 1) * (b.hi[2] - b.lo[2] + 1)
    if name == 'lnst_conv3x3_pool_bf16x3_tc':            # + the pooled copy of the output
        n, H, W, ci, co = [v(x) for x in a[6:11]]
        return (4 * n * H * W * (ci + co) + 4 * n * (H // 2) * (W // 2) * co + 4 * 9 * ci * co, 2 * n * H * W * 9 * ci * co)
    if name == 'lnst_conv3x3_gram_bf16x3_tc':            # data gradient + F x Gd of the layer it lands on
        n, H, W, ci, co = [v(x) for x in a[5:10]]
        return (4 * n * H * W * (ci + 2 * co) + 4 * 9 * ci * co + 4 * n * co * co,
                2 * n * H * W * 9 * ci * co + 2 * n * H * W * co * co)
    if name == 'lnst_conv_first_bwd_gray_dot_tc':
        n, H, W = [v(x) for x in a[6:9]]
        return (n * H * W * (8 + 256), 2 * n * H * W * 9 * 64)
    if name == 'lnst_normalize_ties_fwd':
        return (8 * v(a[2]) * v(a[3]), 0)
    name = ALIASES.get(name, name)
    if name in ('lnst_raymarch_fwd_box', 'lnst_raymarch_bwd_box', 'lnst_raymarch_fwd_tma', 'lnst_raymarch_bwd_tma'):
        nv, D, H, W = [v(x) for x in a[2:6]]
        V, P = box_cells(a[7] if name == 'lnst_raymarch_bwd_tma' else a[8], D * H * W), H * W
        return (nv * (4 * V + 8 * P), 0) if 'fwd' in name else (nv * (8 * V + 8 * P), 0)   # box cells, not bricks
    if name in ('lnst_smooth3_relu_fwd_box', 'lnst_smooth3_relu_bwd_box', 'lnst_smooth3_relu_fwd_tma', 'lnst_smooth3_relu_bwd_tma'):
        fwd = 'fwd' in name
        D, H, W = [v(x) for x in (a[2:5] if fwd else a[3:6])]
        return ((8 if fwd else 12) * box_cells(a[6] if fwd else a[7], D * H * W), 0)
    if name in ('lnst_splat_wavg_fwd_box', 'lnst_splat_wavg_bwd', 'lnst_splat_wavg_bwd_coef'):
        fwd = 'fwd' in name
        n = v(a[3] if fwd else a[2])
        g = (a[4] if fwd else a[3])._obj
        V = g.res[0] * g.res[1] * g.res[2]
        if fwd:
            V = box_cells(a[10], V)
        elif ACTIVE_CELLS:
            V = ACTIVE_CELLS
        return (n * (12 + 8 * nk) + 4 * V, 0)
    if name == 'lnst_fill_box':
        return (4 * box_cells(a[4], v(a[1]) * v(a[2]) * v(a[3])), 0)
    if name == 'lnst_adam_step':
        return (28 * v(a[4]), 0)
    if name == 'lnst_adam_iterate_dev':                  # g_opt, grad, m, v in; m, v, var, delta, g_opt out (+ mask/width)
        return (40 * v(a[4]), 0)
    if name == 'lnst_conv_first_bwd_gray_direct':
        n, H, W = [v(x) for x in a[4:7]]
        return (n * H * W * (4 + (256 if v(a[1]) else 128)), 2 * n * H * W * 9 * 64)
    if name in ('lnst_conv_first_fwd_gray', 'lnst_conv_first_bwd_gray_tc', 'lnst_conv_first_fwd_gray_x3', 'lnst_conv_first_bwd_gray_x3_tc'):
        n, H, W = [v(x) for x in (a[5:8] if 'fwd' in name else a[3:6])]
        return (n * H * W * (4 + 128), 2 * n * H * W * 9 * 64)
    if name in ('lnst_avgpool2_bf16_fwd', 'lnst_avgpool2_bf16_bwd'):
        return (0, 0)
    if name in ('lnst_conv3x3_f32', 'lnst_conv3x3_bf16_tc', 'lnst_conv3x3_bf16x3_tc'):
        n, H, W, ci, co = [v(x) for x in a[5:10]]
        eb = 2 if name.endswith('bf16_tc') else 4            # bf16x3 rows carry hi + lo = 4 bytes per value
        return (eb * n * H * W * (ci + co) + eb * 9 * ci * co, 2 * n * H * W * 9 * ci * co)
    if name in ('lnst_gram_diff_bf16_tc', 'lnst_gram_diff_bf16x3_tc'):
        n, P, ch = v(a[1]), v(a[2]), v(a[3])
        return (n * (2 * P * ch + 6 * ch * ch), 2 * n * P * ch * ch)
    if name in ('lnst_gram_bwd_bf16_tc', 'lnst_gram_bwd_bf16x3_tc'):
        n, H, W, ch = [v(x) for x in a[6:10]]
        return (n * (6 * H * W * ch + 2 * ch * ch), 2 * n * H * W * ch * ch)
    if name in ('lnst_conv2d_f32', 'lnst_conv2d_bwd_data_f32'):      # GraphDef network (multi-net workloads)
        o = 4 if name == 'lnst_conv2d_f32' else 5
        n, H, W, ci, co, kh, kw = [v(x) for x in a[o:o + 7]]
        OH, OW = v(a[o + 10]), v(a[o + 11])
        return (4 * (n * H * W * ci + n * OH * OW * co + kh * kw * ci * co), 2 * n * OH * OW * kh * kw * ci * co)
    if name == 'lnst_gram_diff':
        P, ch = v(a[1]), v(a[2])
        return (4 * P * ch + 4 * ch * ch, 2 * P * ch * ch)
    if name == 'lnst_gram_bwd':
        P, ch = v(a[2]), v(a[3])
        return (8 * P * ch + 4 * ch * ch, 2 * P * ch * ch)
    return (0, 0)


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
# (profiles/), keyed by (workload, entry point); None where no capture exists.
TRAFFIC = {   # bytes per launch, profiles/r1_ncu_full_final2.csv (mean over the entry point's launches of one C3 step)
    ('C3', 'lnst_conv3x3_bf16_tc'): 34.6e6,        # 3 x halo<64> (58.4 MB) + halo<128> resident (11.7) + 4 x halo<128> streamed (22.5)
    ('C3', 'lnst_raymarch_bwd_box'): 25.4e6,
    ('C3', 'lnst_raymarch_fwd_box'): 13.1e6,
    ('C3', 'lnst_splat_wavg_fwd_box'): 86.7e6,     # num kernel 40.4 + combine 46.3
    ('C3', 'lnst_splat_wavg_bwd_coef'): 40.0e6,
    ('C3', 'lnst_smooth3_relu_bwd_box'): 21.8e6,
    ('C3', 'lnst_smooth3_relu_fwd_box'): 11.8e6,
    ('C3', 'lnst_adam_iterate_dev'): 42.1e6,
}


def _load_traffic():
    """round-2 captures: profiles/r2_traffic.json = {workload: {entry point: DRAM bytes per launch}} (ncu --set full)"""
    path = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
    if os.path.exists(path):
        try:
            for wl_, d in json.load(open(path)).items():
                for k, val in d.items():
                    if not k.startswith('_'):
                        TRAFFIC[(wl_, k)] = val
        except Exception:
            pass


_load_traffic()


def roofline_all(table, hbm_peak, tf_peak):
    """Per entry point: achieved algorithmic GB/s or TFLOP/s and its fraction of the measured peak."""
    out = {}
    for name, d in table.items():
        ms = d['ms'] / d['calls']
        if not ms:
            continue
        if name in TENSOR_BOUND and d['flops']:
            a = d['flops'] / (ms * 1e-3) / 1e12
            out[name] = {'bound': 'tensor', 'achieved_tflops': round(a, 2), 'frac': round(a / tf_peak, 4)}
            if name in MMA_PASSES:                          # what the tensor pipe executes for it
                out[name]['mma_tflops'] = round(a * MMA_PASSES[name], 2)
                out[name]['mma_frac'] = round(a * MMA_PASSES[name] / tf_peak, 4)
        elif d['bytes']:
            a = d['bytes'] / (ms * 1e-3) / 1e9
            out[name] = {'bound': 'hbm', 'achieved_gbs': round(a, 1), 'frac': round(a / hbm_peak, 4)}
    return out


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md's clocks line).  NVML is polled from a
    thread every 2 ms -- `nvidia-smi -lms 100` saw nothing of a 20 ms timed region -- with nvidia-smi as the fallback."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')
    BITS = {0x4: 'sw_power_cap', 0x8: 'hw_slowdown', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
            0x80: 'hw_power_brake_slowdown'}

    def __init__(self, index):
        self.index, self.proc, self.lines, self.nvml, self.samples = index, None, [], None, []
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:                                            # CUDA_VISIBLE_DEVICES may renumber: find the device by UUID
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                uuid = uuid if uuid.startswith('GPU-') else 'GPU-' + uuid
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, 'encode') else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml, self.h = pynvml, h
        except Exception:
            self.nvml = None

    def start(self):
        if self.nvml is not None:
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop.is_set():
            try:
                self.samples.append((n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM),
                                     n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h), 0.0))
            except Exception:
                pass
            time.sleep(0.001)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.th.join()
            n = self.nvml
            sm = [c for c, _, _ in self.samples]
            reasons = set()
            for _, bits, _ in self.samples:
                for b, name in self.BITS.items():
                    if bits & b:
                        reasons.add(name)
            try:
                mx = float(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM))
            except Exception:
                mx = None
            return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                    'samples': len(sm), 'source': 'NVML polled during the timed region'}
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        # per-kernel numbers here are event-timed one launch at a time (eager re-issue): the BURST bf16 figure applies
        return p['hbm_gbs'], p['bf16_tflops'], 'measured (MEASURED_PEAKS.json: hbm_gbs, bf16_tflops burst)'
    return 6650.0, 1590.0, 'fallback (B200_PROFILING.md: 6.65 TB/s, 1.59 PFLOP/s burst)'


def _finish(world, dist):
    """End of a rank's work.  With several ranks: meet at a barrier, then leave without tearing the NCCL
    communicator down (ncclCommDestroy with captured collectives alive has been seen to block at exit)."""
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


# ---------------------------------------------------------------------------------------------
class Ctx:
    """process-wide state of the engine arm: ranks, device, the loaded library"""

    def __init__(self):
        import torch.distributed as dist
        from lnst import _lib
        self.dist = dist
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(self.local)
        self.real_stdout = None
        if self.world > 1:
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            # stdout carries exactly ONE JSON line: NCCL prints its version banner with a plain printf when the
            # communicator comes up (NCCL_DEBUG=VERSION/INFO on the box), so fd 1 points at stderr until the line is due
            sys.stdout.flush()
            self.real_stdout = os.dup(1)
            os.dup2(2, 1)
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local))
        self.dev = torch.device('cuda', self.local)
        self.lib = _lib.get()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = torch.tensor([float(x)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def measure_step(ctx, wl, view_mode, conv_math, steps, warmup, full=True):
    """One workload of the 3-D density styler, timed as K CUDA-graph replays of the step (device-resident inputs);
    ``full`` adds the end-to-end loop (host buffers in / out every step), the per-entry-point timing and the clocks."""
    from lnst import ops, synth
    from lnst.styler_3p import Styler, _Adam
    global ACTIVE_CELLS
    dev, lib, world, rank = ctx.dev, ctx.lib, ctx.world, ctx.rank
    cfg = make_cfg(wl, view_mode, conv_math)
    p, r, sty = make_scene(wl)
    styler = Styler(cfg, weights=synth.vgg_weights(), device=dev, content_weights=content_nodes(wl))
    styler.style_img = sty
    res = [WORKLOADS[wl]['res']] * 3
    grams = styler._style_feature(sty, res[1:])
    styler.num_frames = 1
    frames, _ = styler.upload({'p': p, 'r': r})          # device-resident, cell-sorted particles
    ws = styler._workspace(res, frames)                  # + the active box of the particle cloud
    fr = frames[0]
    ACTIVE_CELLS = ws.get('box_cells', 0)
    g_opt = torch.zeros(fr['p'].shape[0], 2, device=dev)
    adam = _Adam()
    lr = cfg.lr
    view_sequential = cfg.rotate and cfg.view_mode == 'sequential'
    runner = styler.step_runner(fr, g_opt, adam, ws, grams, lr)   # eager once, then one CUDA graph per step
    styler.fuse_apply = True                             # g_opt += delta inside the fused Adam/iterate kernel

    def step():
        var, loss, delta = runner()
        if view_sequential:
            ops.axpy(g_opt, delta, 1.0)
        return loss

    for _ in range(max(warmup, 3)):
        step()
    ctx.barrier()
    # ---- timed region: exactly K steps -------------------------------------------------------------
    clocks = ClockSampler(ctx.local)
    ctx.barrier()
    if rank == 0 and full:
        clocks.start()
    launches0 = lib.launches
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        loss = step()
    t1.record()
    ctx.barrier()
    clk = clocks.stop() if (rank == 0 and full) else None
    ms_step = ctx.max_over_ranks(t0.elapsed_time(t1)) / steps
    out = {'ms_step': ms_step, 'value': 1000.0 / ms_step, 'calls_per_step': (lib.launches - launches0) / steps,
           'loss': float(loss), 'clocks': clk, 'cuda_graph': bool(runner.graph is not None),
           'active_cells': ACTIVE_CELLS, 'res': res[0]}
    if not full:
        return out

    # ---- end-to-end: host buffers in, host result out, every step (the reference's sess.run boundary:
    # p, r fed and the variable initialised from the host every step, the variable and the loss read
    # back, styler_3p.py:312,331,334).  p and r do not depend on the previous step, so their upload for
    # step i+1 runs on a copy stream under step i (into staging buffers, then one device copy into the
    # buffers the step's graph reads); the variable's round trip is sequential by construction.
    # With several ranks every rank feeds only ITS 1/W slice of p, r and the variable over PCIe and the slices are
    # all-gathered over NVLink (one collective per tensor), and reads back its slice of the variable (+ the loss).
    n_rows = fr['p'].shape[0]
    chunk = (n_rows + world - 1) // world
    lo_r, hi_r = min(rank * chunk, n_rows), min((rank + 1) * chunk, n_rows)
    hp = fr['p'][lo_r:hi_r].cpu().pin_memory()           # host copies in the engine's (cell-sorted) order
    hr = fr['r'][lo_r:hi_r].cpu().pin_memory()
    hg = g_opt[lo_r:hi_r].cpu().pin_memory()
    hl = torch.zeros(1).pin_memory()
    sp = torch.zeros(world * chunk, fr['p'].shape[1], device=dev)      # staging, padded to W equal chunks
    sr = torch.zeros(world * chunk, fr['r'].shape[1], device=dev)
    sg = torch.zeros(world * chunk, g_opt.shape[1], device=dev)
    cs = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    e2e_steps, e2e_warm = max(3, steps), 2               # K timed steps after two untimed ones (pipeline in steady state)

    def gather(full):
        if world > 1:
            ctx.dist.all_gather_into_tensor(full, full[rank * chunk:(rank + 1) * chunk])

    ctx.barrier()
    uploaded, consumed = torch.cuda.Event(), torch.cuda.Event()
    with torch.cuda.stream(cs):
        sp[lo_r:hi_r].copy_(hp, non_blocking=True)
        sr[lo_r:hi_r].copy_(hr, non_blocking=True)
        uploaded.record(cs)
    w0 = None
    for i in range(e2e_warm + e2e_steps):
        if i == e2e_warm:                                # the timed region: exactly e2e_steps steps, each with its own
            torch.cuda.synchronize()                     # H2D of p, r and the variable and its D2H of variable + loss
            ctx.barrier()                                # (p, r of step i travel under step i - 1: one upload per timed step)
            w0 = time.perf_counter()
        main.wait_event(uploaded)
        gather(sp)
        gather(sr)
        fr['p'].copy_(sp[:n_rows], non_blocking=True)
        fr['r'].copy_(sr[:n_rows], non_blocking=True)
        consumed.record(main)
        with torch.cuda.stream(cs):                      # next step's particle upload, under this step
            cs.wait_event(consumed)
            sp[lo_r:hi_r].copy_(hp, non_blocking=True)
            sr[lo_r:hi_r].copy_(hr, non_blocking=True)
            uploaded.record(cs)
        sg[lo_r:hi_r].copy_(hg, non_blocking=True)
        gather(sg)
        g_opt.copy_(sg[:n_rows], non_blocking=True)
        l = step()
        hg.copy_(g_opt[lo_r:hi_r], non_blocking=True)
        hl.copy_(l.reshape(1), non_blocking=True)
        main.synchronize()
    torch.cuda.synchronize()
    ctx.barrier()
    e2e_s = ctx.max_over_ranks(time.perf_counter() - w0)
    out['e2e'] = {'value': e2e_steps / e2e_s, 'unit': 'iters/s',
                  'h2d_bytes_per_step': world * (hp.numel() * 4 + hr.numel() * 4 + hg.numel() * 4),
                  'd2h_bytes_per_step': world * (hg.numel() * 4 + 4), 'steps': e2e_steps, 'warmup': e2e_warm,
                  'per_rank': 'each rank moves 1/%d of the rows over PCIe; slices all-gathered over NVLink' % world if world > 1 else 'single rank',
                  'boundary': "the reference's sess.run boundary per step (styler_3p.py:312,331,334): p, r and the variable "
                              'uploaded from pinned host memory, variable + loss read back'}

    # ---- per-kernel timing: the same steps issued eagerly (a graph replay cannot carry events), every
    # entry point bracketed by CUDA events on the launching stream -----------------------------------
    prof = {}
    orig_call = lib.call

    def profiling_call(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_call(name, *a)
        e1.record()
        prof.setdefault(ALIASES.get(name, name), []).append((e0, e1, algorithmic_units(name, a)))

    prof_steps = max(3, min(steps, 10))

    def eager_step():
        var, loss_e, delta = styler.frame_step(fr, g_opt, adam, ws, grams, lr)
        if view_sequential:                               # otherwise applied inside lnst_adam_iterate_dev
            ops.axpy(g_opt, delta, 1.0)

    for _ in range(2):                                    # the eager path's own allocator warm-up (untimed)
        eager_step()
    ctx.barrier()
    lib.call = profiling_call
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(prof_steps):
        eager_step()
    p1.record()
    ctx.barrier()
    lib.call = orig_call
    out['eager_ms_step'] = p0.elapsed_time(p1) / prof_steps
    out['prof_steps'] = prof_steps
    table = {}
    for name, evs in prof.items():
        ms = [a.elapsed_time(b) for a, b, _ in evs]
        table[name] = {'calls': len(ms) / prof_steps, 'ms': float(sum(ms)) / prof_steps,
                       'bytes': sum(u[0] for _, _, u in evs) / len(evs), 'flops': sum(u[1] for _, _, u in evs) / len(evs)}
    out['table'] = table
    return out


def roofline_entry(wl, table, ms_step, hbm_peak, tf_peak, src, extra):
    dominant = max(table, key=lambda k: table[k]['ms'])
    d = table[dominant]
    avg_ms = d['ms'] / d['calls']
    if dominant in TENSOR_BOUND:
        achieved = d['flops'] / (avg_ms * 1e-3) / 1e12
        roof = {'kernel': dominant, 'bound': 'tensor', 'achieved': achieved, 'peak': tf_peak, 'unit': 'TFLOP/s',
                'frac': achieved / tf_peak}
        if dominant in MMA_PASSES:
            roof['note'] = ('achieved = algorithmic (fp32) flops; the bf16x3 kernel executes %d bf16 MMA passes per '
                            'algorithmic multiply-add: tensor pipe at %.1f TFLOP/s = %.3f of peak'
                            % (MMA_PASSES[dominant], achieved * MMA_PASSES[dominant], achieved * MMA_PASSES[dominant] / tf_peak))
    else:
        achieved = d['bytes'] / (avg_ms * 1e-3) / 1e9
        roof = {'kernel': dominant, 'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s',
                'frac': achieved / hbm_peak}
    roof.update({'traffic': TRAFFIC.get((wl, dominant)), 'peak_source': src, 'avg_launch_ms': avg_ms,
                 'share_of_step': d['ms'] / ms_step})
    roof.update(extra)
    return roof


def time_run(make_styler, params, iters_a, iters_b, **run_kw):
    """steady-state iterations/s of a drop-in ``Styler.run``.  3-D styler: CUDA events the run loop records at the start of
    every iteration (``Styler.iter_events``; iterations 0 and 1 are the eager pass and the graph captures).  Stylers without
    that hook (2-D): two runs with different iteration budgets, the difference of their wall clocks over the difference in
    iterations (set-up and the final inference cancel; minimum over repeats)."""
    st = make_styler(iters_b)
    hook = hasattr(st, 'frame_step')
    if hook:
        st.iter_events = []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = st.run(params, **run_kw)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ev = st.iter_events
        per_iter = ev[2].elapsed_time(ev[-1]) / (len(ev) - 3) * 1e-3
        del st
        return {'value': 1.0 / per_iter, 'unit': 'iters/s', 'ms_per_iter': 1e3 * per_iter, 'run_wall_s': wall,
                'run_iters': iters_b, 'final_loss': float(np.asarray(out['l'][-1]).reshape(-1)[-1]),
                'timed': 'CUDA events around iterations 2..%d of one Styler.run' % (iters_b - 1)}
    del st
    walls = {iters_a: [], iters_b: []}
    for k, it in enumerate((iters_a, iters_b, iters_a, iters_b, iters_a)):   # the first run is an untimed warm-up
        st = make_styler(it)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = st.run(params, **run_kw)
        torch.cuda.synchronize()
        if k:
            walls[it].append(time.perf_counter() - t0)
        del st
    walls = [min(walls[iters_a]), min(walls[iters_b])]
    per_iter = max((walls[1] - walls[0]) / (iters_b - iters_a), 1e-9)
    return {'value': 1.0 / per_iter, 'unit': 'iters/s', 'ms_per_iter': 1e3 * per_iter, 'run_wall_s': walls[1],
            'run_iters': iters_b, 'final_loss': float(np.asarray(out['l'][-1]).reshape(-1)[-1]),
            'timed': 'wall-clock difference of runs with %d and %d iterations' % (iters_a, iters_b)}


def other_configs(ctx, conv_math, hbm_peak, tf_peak, src):
    """BASELINE.json configs[0], [1], [3], [4] on this GPU; the headline workload is configs[2]."""
    from helpers import dam_cfg, liquid_cfg
    from lnst import synth
    res = {}
    # C2: 128^3, one view -------------------------------------------------------------------------------------------
    try:
        m = measure_step(ctx, 'C2', 'allreduce', conv_math, steps=20, warmup=3, full=True)
        res['C2'] = {'workload': WORKLOADS['C2']['desc'], 'value': m['value'], 'unit': 'iters/s', 'ms_per_step': m['ms_step'],
                     'e2e': m['e2e']['value'], 'conv_math': conv_math,
                     'roofline': roofline_entry('C2', m['table'], m['ms_step'], hbm_peak, tf_peak, src, {})}
    except Exception as e:                                 # pragma: no cover
        res['C2'] = {'error': repr(e)}
    torch.cuda.empty_cache()
    # C1: 2-D colour, 256^2, conv1_1, 50 iterations (the reference's own CPU-runnable case) ------------------------------
    try:
        from lnst.styler_2p import Styler as Styler2
        cfg_of = lambda it: dam_cfg(resolution=[256, 256], domain=[6.4, 6.4], radius=0.025, nsize=4, support=4, iter=it,
                                    lr=0.01, octave_n=1, style_layer=['conv1_1'], w_style_layer=[1.0], w_style=1, w_tv=0,
                                    style_mask=False, conv_math='fp32')
        p2, r2 = synth.dam_particles_2d(cfg_of(1).domain)
        sty2 = synth.style_image(256, 256)
        c0 = np.random.RandomState(5).uniform(0.2, 0.8, (1, p2[0].shape[0], 3)).astype(np.float32)

        def mk2(it):
            st = Styler2(cfg_of(it), weights=synth.vgg_weights(), device=ctx.dev)
            st.style_img = sty2
            return st
        r_ = time_run(mk2, {'p': p2, 'r': r2}, 50, 150, c_init=c0)
        r_.update({'workload': 'C1: dambreak2d-like single frame, 2-D 256x256 colour field (81 splat taps), VGG-19 conv1_1, '
                               '50 Adam iterations; N = %d particles' % p2[0].shape[0], 'conv_math': 'fp32 (Cin = 3: CUDA cores)',
                   })
        res['C1'] = r_
    except Exception as e:                                 # pragma: no cover
        res['C1'] = {'error': repr(e)}
    torch.cuda.empty_cache()
    # C4: 60-frame liquid sequence, 128^3, position mode, temporal filter ---------------------------------------------------
    try:
        from lnst.styler_3p import Styler as Styler3
        nf, n4 = 60, 200000
        cfg4 = lambda it: liquid_cfg(res=128, iter=it, num_frames=nf, window_sigma=9, frames_per_opt=1, lr=0.002,
                                     conv_math=conv_math, style_layer=['conv2_1', 'conv3_1'], w_style_layer=[0.5, 0.5])
        p4 = synth.liquid_particles(n4, num_frames=nf)
        sty4 = synth.style_image(128, 128)

        def mk4(it):
            st = Styler3(cfg4(it), weights=synth.vgg_weights(), device=ctx.dev)
            st.style_img = sty4
            return st
        r_ = time_run(mk4, {'p': p4}, 4, 24)
        r_.update({'workload': 'C4: chocolate-like sequence, %d frames, 128^3, position mode (SPH splat, liquid render), '
                               'N = %d particles per frame, temporal Gaussian sigma 9; one iteration = all %d frames'
                               % (nf, n4, nf), 'conv_math': conv_math, 'frame_steps_per_s': nf / (r_['ms_per_iter'] * 1e-3)})
        res['C4'] = r_
    except Exception as e:                                 # pragma: no cover
        res['C4'] = {'error': repr(e)}
    torch.cuda.empty_cache()
    # C5: 256^3, 9 views, multi-net (inception5h semantic + VGG-19 style) ---------------------------------------------------
    try:
        m = measure_step(ctx, 'C5', 'allreduce', conv_math, steps=3, warmup=3, full=True)
        res['C5'] = {'workload': WORKLOADS['C5']['desc'], 'value': m['value'], 'unit': 'iters/s', 'ms_per_step': m['ms_step'],
                     'e2e': m['e2e']['value'], 'conv_math': conv_math + ' (VGG), fp32 CUDA cores (GraphDef network)',
                     'kernel_table_ms_per_step': {k: round(v['ms'], 3) for k, v in sorted(m['table'].items(), key=lambda kv: -kv[1]['ms'])[:8]},
                     'roofline': roofline_entry('C5', m['table'], m['ms_step'], hbm_peak, tf_peak, src, {})}
    except Exception as e:                                 # pragma: no cover
        res['C5'] = {'error': repr(e)}
    torch.cuda.empty_cache()
    return res


def run_engine(args):
    from lnst import synth
    ctx = Ctx()
    world, rank, dist, lib = ctx.world, ctx.rank, ctx.dist, ctx.lib
    conv_math = args.conv_math or ('bf16x3' if lib.has_tc else 'fp32')
    wl = args.workload
    m = measure_step(ctx, wl, args.view_mode, conv_math, args.steps, args.warmup, full=True)
    ms_step, table, res = m['ms_step'], m['table'], [m['res']] * 3
    extras = {}
    if conv_math == 'bf16x3':                                # the single-pass bf16 number beside it (outside the fp32 tolerance)
        torch.cuda.empty_cache()
        b = measure_step(ctx, wl, args.view_mode, 'bf16', min(args.steps, 10), 3, full=False)
        extras['value_bf16'] = b['value']
        extras['ms_per_step_bf16'] = b['ms_step']
    if world == 1 and args.view_mode == 'allreduce' and WORKLOADS[wl]['rotate'] and not args.quick:
        torch.cuda.empty_cache()
        q = measure_step(ctx, wl, 'sequential', conv_math, min(args.steps, 10), 3, full=False)
        extras['value_sequential'] = q['value']                # the reference-exact view mode (styler_3p.py:329-352)
        extras['ms_per_step_sequential'] = q['ms_step']
    hbm_peak, tf_peak, src = peaks()
    if world == 1 and not args.quick:
        torch.cuda.empty_cache()
        # wall clock of the drop-in call itself: host NumPy lists -> result dict, the reference's 20-iteration budget
        from lnst.styler_3p import Styler
        p, r, sty = make_scene(wl)

        def mk(it):
            cfg = make_cfg(wl, args.view_mode, conv_math)
            cfg.iter = it
            st = Styler(cfg, weights=synth.vgg_weights(), device=ctx.dev)
            st.style_img = sty
            return st
        walls, final_loss = [], None
        for _ in range(3):                                   # later calls find the first one's device / pinned blocks cached
            st = mk(20)
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            out_run = st.run({'p': p, 'r': r})
            torch.cuda.synchronize()
            walls.append(time.perf_counter() - w0)
            final_loss = float(out_run['l'][0][-1])
            del st, out_run                                  # the result arrays are views of the run's pinned staging buffer
        wall = min(walls)
        extras['e2e_run'] = {'iters': 20, 'wall_s': wall, 'wall_s_each_call': walls, 'value': 20.0 / wall, 'unit': 'iters/s',
                             'what': 'Styler(config).run(params): upload + cell sort + weight maps + workspace + graph capture '
                                     '+ 20 iterations (test_smokegun.py:143-146) + final inference + D2H of the results; '
                                     'best of three calls (the first one pays cudaMalloc / cudaHostAlloc)',
                             'final_loss': final_loss}
        torch.cuda.empty_cache()
        extras['configs'] = other_configs(ctx, conv_math, hbm_peak, tf_peak, src)
    if rank != 0:
        _finish(world, dist)
        return
    roof = roofline_entry(wl, table, ms_step, hbm_peak, tf_peak, src, {
        'launches_timed': int(max(table.values(), key=lambda d: d['ms'])['calls'] * m['prof_steps']),
        'timed_in': 'eager re-issue of %d steps after the timed region (graph replays cannot carry events); '
                    'eager step = %.3f ms' % (m['prof_steps'], m['eager_ms_step'])})
    kl = 0
    for name, t in table.items():
        kl += t['calls'] * KERNELS_PER_CALL.get(name, 1)
    kl += 1  # axpy
    dtype = {'bf16x3': 'f32 (loss-net convolutions on tensor cores as three bf16 passes per product, hi/lo split operands, '
                       'f32 accumulate: fp32-tolerance results)',
             'bf16': 'f32 (loss-net convolutions: bf16 operands, f32 accumulate)', 'fp32': 'f32'}[conv_math]
    out = {
        'metric': 'style-opt iters/sec, 200^3 smoke x9 views', 'value': 1000.0 / ms_step, 'unit': 'iters/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_step,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': dtype,
        'data': 'synthetic (seeded ellipsoid particle cloud, He-normal VGG-19 weights, low-pass noise style image)',
        'config': {'workload': '%s: %s' % (wl, WORKLOADS[wl]['desc']), 'view_mode': args.view_mode,
                   'conv_math': conv_math, 'active_box_fraction': round(m['active_cells'] / float(res[0] ** 3), 4) if m['active_cells'] else 1.0,
                   'views_per_rank': [len(range(k, WORKLOADS[wl]['n_views'], world)) for k in range(world)],
                   'l2': 'per-step working set (8 volumes x %.0f MB + activations) exceeds the 126 MB L2; no flush'
                         % (4e-6 * res[0] ** 3)},
        'e2e': m['e2e'],
        'gpu_launches': int(kl * args.steps), 'abi_calls_per_step': m['calls_per_step'],
        'cuda_graph': m['cuda_graph'],
        'clocks': m['clocks'], 'roofline': roof, 'final_loss': m['loss'],
        'kernel_table_ms_per_step': {k: round(v['ms'], 4) for k, v in sorted(table.items(), key=lambda kv: -kv[1]['ms'])},
        'roofline_all': roofline_all(table, hbm_peak, tf_peak),
    }
    out.update(extras)
    if world == 1 and not args.no_cpu_baseline:
        out['cpu_baseline'] = cpu_baseline(wl, budget_s=args.cpu_budget, view_mode=args.view_mode)
    if ctx.real_stdout is not None:
        sys.stdout.flush()
        os.dup2(ctx.real_stdout, 1)
    print(json.dumps(out), flush=True)
    _finish(world, dist)


# ---------------------------------------------------------------------------------------------
def cpu_baseline(wl, budget_s=25.0, iters=None, warmup_views=0, view_mode='allreduce'):
    """The oracle (CPU restatement of the reference's TF-1.15 graph) on this box's host cores, in the SAME view mode
    as the engine arm: 'allreduce' = gradients of the n_views views summed, one Adam step per iteration;
    'sequential' = one Adam step per view (styler_3p.py:329-340).  Either way an iteration is n_views full-size
    forward + backward passes.  Timed in whole iterations when the budget allows at least one; otherwise the
    sample is the view passes that fit and the iteration time is n_views x their mean (said so in ``sample``)."""
    import oracle.vgg
    from oracle.styler import Oracle3P
    from oracle.adam import TFAdam
    torch.set_num_threads(os.cpu_count() or 1)
    w = WORKLOADS[wl]
    nv = w['n_views'] if w['rotate'] else 1
    cfg = make_cfg(wl, view_mode, 'fp32')
    p, r, sty = make_scene(wl)
    o = Oracle3P(cfg, oracle.vgg.synthetic_weights(), content_weights=content_nodes(wl))
    res = [w['res']] * 3
    sf = o.style_features(sty)
    pt, rt = torch.tensor(p[0]), torch.tensor(r[0])
    var = torch.zeros(pt.shape[0], 2)
    rot = o.views()[0] if w['rotate'] else None
    adam = TFAdam()
    view_times, iter_times = [], []
    t_all = time.perf_counter()
    for _ in range(warmup_views):
        o.loss_and_grad([pt], [rt], [var], res, rot[0:1] if rot is not None else None, sf, None)
    it = 0
    deadline = t_all + budget_s
    while iters is None or it < iters:
        t_it = time.perf_counter()
        gsum, partial = None, False
        for v in range(nv):
            t0 = time.perf_counter()
            rv = rot[v:v + 1] if rot is not None else None
            _, gr = o.loss_and_grad([pt], [rt], [var], res, rv, sf, None)
            if view_mode == 'sequential' or nv == 1:
                var = torch.nan_to_num(adam.step(var, gr[0], cfg.lr))
            else:
                gsum = gr[0] if gsum is None else gsum + gr[0]
            view_times.append(time.perf_counter() - t0)
            if not iter_times and v < nv - 1 and time.perf_counter() > deadline + 2 * budget_s:
                partial = True                            # not even one iteration fits three budgets: extrapolate
                break
        if partial:
            break
        if gsum is not None:
            var = torch.nan_to_num(adam.step(var, gsum / nv, cfg.lr))
        iter_times.append(time.perf_counter() - t_it)
        it += 1
        if time.perf_counter() > deadline:
            break
    if iter_times:
        t_iter = float(np.mean(iter_times))
        sample = ('%d whole iterations of %s at full size (%d views each: forward + backward per view, %s), %s view mode; '
                  'PyTorch-CPU fp32 restatement of the TF-1.15 graph (TF itself cannot run here)'
                  % (len(iter_times), wl, nv, 'one Adam step per view' if view_mode == 'sequential' else 'summed gradient, one Adam step',
                     view_mode))
    else:
        t_iter = float(np.mean(view_times)) * nv
        sample = ('%d view passes of %s at full size (the time budget ended before one whole %d-view iteration); '
                  'iteration time = %d x their mean; PyTorch-CPU fp32 restatement' % (len(view_times), wl, nv, nv))
    return {'value': 1.0 / t_iter, 'unit': 'iters/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'seconds_per_iteration': t_iter, 'seconds_per_view_pass': float(np.mean(view_times)),
            'iterations_timed': len(iter_times), 'view_passes_timed': len(view_times), 'view_mode': view_mode, 'sample': sample}


def run_reference(args):
    """The reference's CPU path (oracle port) on the engine arm's config, metric and unit.  A step = one whole iteration
    (n_views forward + backward passes at full size, ~12 s at C3 on 16 cores), so the run is bounded by a time budget:
    ``steps`` in the line = the iterations actually timed (at least one)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    budget = max(30.0, min(150.0, args.cpu_budget * 6))
    cb = cpu_baseline(args.workload, budget_s=budget, iters=max(1, args.steps), warmup_views=1 if args.warmup else 0,
                      view_mode=args.view_mode)
    wl = args.workload
    n_it = max(cb['iterations_timed'], 1)
    out = {'impl': 'reference', 'metric': 'style-opt iters/sec, 200^3 smoke x9 views', 'value': cb['value'],
           'unit': 'iters/s', 'n_gpus': 0, 'steps': n_it, 'steps_requested': args.steps,
           'warmup': 1 if args.warmup else 0, 'warmup_unit': 'one view pass',
           'ms_per_step': 1000.0 / cb['value'], 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
           'dtype': 'f32', 'data': 'synthetic (same seeded inputs as the engine arm)',
           'config': {'workload': '%s: %s' % (wl, WORKLOADS[wl]['desc']), 'view_mode': args.view_mode,
                      'conv_math': 'fp32'},
           'cpu_baseline': cb,
           'e2e': {'value': cb['value'], 'unit': 'iters/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='engine', choices=['engine', 'reference'])
    ap.add_argument('--workload', default='C3', choices=list(WORKLOADS))
    ap.add_argument('--view-mode', dest='view_mode', default='allreduce', choices=['allreduce', 'sequential'])
    ap.add_argument('--conv-math', dest='conv_math', default=None, choices=['bf16', 'bf16x3', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--quick', action='store_true', help='headline workload only: no sequential-mode, e2e_run or other-config entries')
    ap.add_argument('--cpu-budget', type=float, default=25.0)
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_engine(args)


if __name__ == '__main__':
    main()
