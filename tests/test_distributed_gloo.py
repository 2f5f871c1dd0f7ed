"""Multi-rank path on CPU: world_size-2 gloo run of ``view_mode='allreduce'`` (views sharded over
ranks, one all-reduce of d loss/d var per Adam step) must reproduce the single-process result.
The kernels run through the CPU interpreter (test tooling); the collective is the real
torch.distributed code path that NCCL takes on the GPUs."""
import os
import sys

import numpy as np
import pytest
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
    res = 10
    kw = dict(res=res, iter=2, rotate=True, n_views=9, view_mode='allreduce', conv_math='fp32',
              style_layer=['conv1_2'], w_style_layer=[1.0])
    p, r = synth.smoke_particles(500, 2, pad=2)
    sty = synth.style_image(res, res)
    st = Styler(smoke_cfg(**kw), weights=synth.vgg_weights())
    assert (st.rank, st.world) == (rank, world)
    st.style_img = sty
    out = st.run({'p': p, 'r': r})
    if rank == 0:
        solo = Styler(smoke_cfg(**kw), weights=synth.vgg_weights())
        solo.set_world(0, 1)                      # same process, no sharding
        solo.style_img = sty
        ref = solo.run({'p': p, 'r': r})
        np.savez(out_path, l=np.array(out['l']), l_ref=np.array(ref['l']), g=out['g_opt'][0], g_ref=ref['g_opt'][0],
                 d=out['d'], d_ref=ref['d'])
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_match_one(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, 'tools', 'cpu_emu'))
    import build_emu
    build_emu.build()
    out = str(tmp_path / 'res.npz')
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    np.testing.assert_allclose(z['l'], z['l_ref'], rtol=1e-5)
    assert np.linalg.norm(z['g'] - z['g_ref']) <= 1e-3 * np.linalg.norm(z['g_ref'])
    assert np.abs(z['d'] - z['d_ref']).max() <= 1e-4 * np.abs(z['d_ref']).max()


def _frames_worker(rank, world, port, out_path):
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
    res = 10
    # 5 frames in Adam groups of 2 (groups {0,1},{2,3},{4}): rank 0 owns frames 0-3, rank 1 frame 4;
    # temporal filter on, two octaves (intermediate renders are gathered too)
    kw = dict(res=res, iter=2, conv_math='fp32', num_frames=5, window_sigma=1.5, frames_per_opt=2, octave_n=2,
              octave_scale=1.25, style_layer=['conv1_2'], w_style_layer=[1.0])
    p, r = synth.smoke_particles(400, 2, pad=2, num_frames=5)
    from lnst.util import octave_sizes
    st = Styler(smoke_cfg(**kw), weights=synth.vgg_weights())
    st.style_img = synth.style_image(res, res)
    out = st.run({'p': p, 'r': r})
    if rank == 0:
        solo = Styler(smoke_cfg(**kw), weights=synth.vgg_weights())
        solo.set_world(0, 1)
        solo.style_img = synth.style_image(res, res)
        ref = solo.run({'p': p, 'r': r})
        np.savez(out_path, l0=np.array(out['l'][0]), l0_ref=np.array(ref['l'][0]), l1=np.array(out['l'][1]),
                 l1_ref=np.array(ref['l'][1]), g=np.stack(out['g_opt']), g_ref=np.stack(ref['g_opt']), d=out['d'],
                 d_ref=ref['d'], i=out['d_intm'][0], i_ref=ref['d_intm'][0])
    dist.barrier()
    dist.destroy_process_group()


def test_frames_sharded_over_two_ranks_match_one(tmp_path):
    """A sequence sharded by frames (owners = contiguous Adam groups) with the per-iteration gather for
    the temporal filter reproduces the single-process run: losses in the reference's (step, frame)
    order, variables, fields and the intermediate octave renders."""
    sys.path.insert(0, os.path.join(ROOT, 'tools', 'cpu_emu'))
    import build_emu
    build_emu.build()
    out = str(tmp_path / 'res.npz')
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_frames_worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    np.testing.assert_allclose(z['l0'], z['l0_ref'], rtol=1e-5)
    np.testing.assert_allclose(z['l1'], z['l1_ref'], rtol=1e-5)
    assert z['l0'].shape == (10,)
    assert np.linalg.norm(z['g'] - z['g_ref']) <= 1e-4 * np.linalg.norm(z['g_ref'])
    assert np.abs(z['d'] - z['d_ref']).max() <= 1e-5 * np.abs(z['d_ref']).max()
    assert np.abs(z['i'].astype(int) - z['i_ref'].astype(int)).max() <= 1


def test_view_sharding_covers_all_views_once():
    for world in (1, 2, 4, 8):
        got = sorted(v for r in range(world) for v in range(r, 9, world))
        assert got == list(range(9))


def _a2a_worker(rank, world, port, out_path):
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
    res = 10
    # 6 frames, one optimizer per frame: 3 frames per rank -> the all-to-all transposes apply; 401 particles (odd, so the
    # particle axis is padded to a multiple of the world size)
    kw = dict(res=res, iter=2, conv_math='fp32', num_frames=6, window_sigma=1.5, frames_per_opt=1,
              style_layer=['conv1_2'], w_style_layer=[1.0])
    p, r = synth.smoke_particles(401, 2, pad=0, num_frames=6)
    st = Styler(smoke_cfg(**kw), weights=synth.vgg_weights())
    st.style_img = synth.style_image(res, res)
    used = []
    orig = st._filter_frames_alltoall
    st._filter_frames_alltoall = lambda *a: (used.append(1), orig(*a))[1]
    out = st.run({'p': p, 'r': r})
    assert len(used) == 2, 'the all-to-all exchange must run once per iteration'
    if rank == 0:
        solo = Styler(smoke_cfg(**kw), weights=synth.vgg_weights())
        solo.set_world(0, 1)
        solo.style_img = synth.style_image(res, res)
        ref = solo.run({'p': p, 'r': r})
        np.savez(out_path, l=np.array(out['l'][0]), l_ref=np.array(ref['l'][0]), g=np.stack(out['g_opt']),
                 g_ref=np.stack(ref['g_opt']), d=out['d'], d_ref=ref['d'])
    dist.barrier()
    dist.destroy_process_group()


def test_frames_sharded_alltoall_exchange_matches_one_rank(tmp_path):
    """the temporal filter of a frame-sharded sequence through two all-to-all transposes (frame-sharded <->
    particle-sharded) reproduces the single-process run"""
    sys.path.insert(0, os.path.join(ROOT, 'tools', 'cpu_emu'))
    import build_emu
    build_emu.build()
    out = str(tmp_path / 'res.npz')
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_a2a_worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    np.testing.assert_allclose(z['l'], z['l_ref'], rtol=1e-5)
    assert np.linalg.norm(z['g'] - z['g_ref']) <= 1e-4 * np.linalg.norm(z['g_ref'])
    assert np.abs(z['d'] - z['d_ref']).max() <= 1e-5 * np.abs(z['d_ref']).max()
