"""2-GPU NCCL run of the multi-net loss (inception semantic term on a GraphDef network + VGG-19 style on the tcgen05
path; BASELINE configs[4] scaled down): views sharded over the ranks, one all-reduce of (d loss / d var, loss) per Adam
step inside the step's CUDA graph, must reproduce the single-GPU result.  Needs >= 2 GPUs (``gpurun --gpus 2``);
skipped on a single-GPU box.  CPU twin: tests/test_widen_zz_gloo_inception.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out_path):
    sys.path[:0] = [ROOT, os.path.join(ROOT, 'neural-flow-style_b200'), os.path.join(ROOT, 'tests')]
    import torch.distributed as dist
    from helpers import smoke_cfg
    from lnst import synth
    from lnst.styler_3p import Styler
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    res = 24
    kw = dict(res=res, iter=4, rotate=True, n_views=9, view_mode='allreduce', conv_math='bf16',
              style_layer=['conv2_1', 'conv3_1'], w_style_layer=[0.5, 0.5],
              content_network='tensorflow_inception_graph.pb', w_content=20.0,
              content_layer='mixed3b_3x3_bottleneck_pre_relu', content_channel=5)
    nodes = synth.inception5h_nodes(width_div=8, upto='mixed3b')
    p, r = synth.smoke_particles(6000, 2, pad=8)
    sty = synth.style_image(res, res)
    st = Styler(smoke_cfg(**kw), weights=synth.vgg_weights(), content_weights=nodes)
    assert (st.rank, st.world) == (rank, world)
    st.style_img = sty
    out = st.run({'p': p, 'r': r})
    if rank == 0:
        solo = Styler(smoke_cfg(**kw), weights=synth.vgg_weights(), content_weights=nodes)
        solo.set_world(0, 1)
        solo.style_img = sty
        ref = solo.run({'p': p, 'r': r})
        np.savez(out_path, l=np.array(out['l']), l_ref=np.array(ref['l']), g=out['g_opt'][0], g_ref=ref['g_opt'][0],
                 d=out['d'], d_ref=ref['d'])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpus_match_one_multinet(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    out = str(tmp_path / 'res.npz')
    port = 29900 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    np.testing.assert_allclose(z['l'], z['l_ref'], rtol=2e-3)
    assert np.linalg.norm(z['g'] - z['g_ref']) <= 2e-2 * np.linalg.norm(z['g_ref'])
    assert np.abs(z['d'] - z['d_ref']).max() <= 2e-2 * np.abs(z['d_ref']).max()
