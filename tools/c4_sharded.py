#!/usr/bin/env python
"""BASELINE.json configs[3] with the frames sharded over the ranks: 60-frame liquid sequence, 128^3,
position mode, per-frame Adam, temporal Gaussian sigma 9 -- iterations/s for both exchange variants of the temporal
filter (two all-to-all transposes | one all-gather).  Run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/c4_sharded.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'neural-flow-style_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from helpers import liquid_cfg
    from lnst import synth
    from lnst.styler_3p import Styler
    nf, n = 60, 200000
    p = synth.liquid_particles(n, num_frames=nf)
    sty = synth.style_image(128, 128)
    out = {'workload': 'C4: 60 frames, 128^3, position mode, N = 200000 per frame, temporal Gaussian sigma 9', 'n_gpus': world}
    for mode in ('alltoall', 'allgather'):
        # steady-state rate from CUDA events the run loop records at the start of every iteration (iterations 0 and 1 are
        # the eager pass and the graph captures): 22 replayed iterations, max over ranks
        iters = 25
        cfg = liquid_cfg(res=128, iter=iters, num_frames=nf, window_sigma=9, frames_per_opt=1, lr=0.002, conv_math='bf16x3',
                         style_layer=['conv2_1', 'conv3_1'], w_style_layer=[0.5, 0.5])
        st = Styler(cfg, weights=synth.vgg_weights(), device=torch.device('cuda', local))
        st.style_img = sty
        st.frame_exchange = mode
        st.iter_events = []
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = st.run({'p': p})
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ev = st.iter_events
        ms = torch.tensor([ev[2].elapsed_time(ev[-1]) / (len(ev) - 3)], device=torch.device('cuda', local))
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        per_iter = float(ms.item()) * 1e-3
        del st
        torch.cuda.empty_cache()
        out[mode] = {'iters_per_s': 1.0 / per_iter, 'ms_per_iter': 1e3 * per_iter, 'frame_steps_per_s': nf / per_iter,
                     'run_wall_s_%d_iters' % iters: wall, 'final_loss': float(res['l'][0][-1])}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


if __name__ == '__main__':
    main()
