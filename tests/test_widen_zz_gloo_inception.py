"""World-size-2 gloo run of the widened loss networks: views sharded over ranks (``view_mode='allreduce'``) with the
GraphDef (inception) network carrying the content term and VGG the style term (multi-net loss) must reproduce the
single-process result.  Kernels on the CPU interpreter (test tooling); the collective is the real torch.distributed
path NCCL takes on the GPUs."""
import os
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out_path):
    sys.path[:0] = [ROOT, os.path.join(ROOT, 'neural-flow-style_b200'), os.path.join(ROOT, 'tests'),
                    os.path.join(ROOT, 'tools', 'cpu_emu')]
    import torch.distributed as dist
    import build_emu
    from helpers import smoke_cfg
    from lnst import _lib, synth
    from lnst.styler_3p import Styler
    torch.set_num_threads(1)
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    _lib.set_for_testing(_lib.Lib(build_emu.build(), 'emu'))
    res = 16
    kw = dict(res=res, iter=2, rotate=True, n_views=5, view_mode='allreduce', conv_math='fp32',
              style_layer=['conv1_2'], w_style_layer=[1.0], content_network='tensorflow_inception_graph.pb',
              w_content=20.0, content_layer='mixed3a_1x1_pre_relu', content_channel=2)
    nodes = synth.inception5h_nodes(width_div=16, upto='mixed3a')
    p, r = synth.smoke_particles(500, 2, pad=2)
    sty = synth.style_image(res, res)
    st = Styler(smoke_cfg(**kw), weights=synth.vgg_weights(), content_weights=nodes)
    assert (st.rank, st.world) == (rank, world)
    st.style_img = sty
    out = st.run({'p': p, 'r': r})
    if rank == 0:
        solo = Styler(smoke_cfg(**kw), weights=synth.vgg_weights(), content_weights=nodes)
        solo.set_world(0, 1)                      # same process, no sharding
        solo.style_img = sty
        ref = solo.run({'p': p, 'r': r})
        np.savez(out_path, l=np.array(out['l']), l_ref=np.array(ref['l']), g=out['g_opt'][0], g_ref=ref['g_opt'][0],
                 d=out['d'], d_ref=ref['d'])
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_match_one_multinet(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, 'tools', 'cpu_emu'))
    import build_emu
    build_emu.build()
    out = str(tmp_path / 'res.npz')
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    np.testing.assert_allclose(z['l'], z['l_ref'], rtol=1e-5)
    assert np.linalg.norm(z['g'] - z['g_ref']) <= 1e-3 * np.linalg.norm(z['g_ref'])
    assert np.abs(z['d'] - z['d_ref']).max() <= 1e-4 * np.abs(z['d_ref']).max()
