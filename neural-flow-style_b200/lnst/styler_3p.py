"""3-D particle styler -- drop-in for the reference's ``styler_3p.Styler`` (``styler_3p.py:14-438``).

Same constructor / ``load_img`` / ``run(params)`` / result-dict contract; underneath, the TF graph
and its three ``sess.run`` round trips per (frame, view, iteration) are replaced by a fixed
sequence of sm_100a kernels on one stream.  Particle data, the optimisation variables, Adam
moments and the loss history stay resident in HBM for the whole run; the host sees them
once, at the end.

Per (frame, view-batch) step:   var -> splat -> smooth+ReLU -> [rotate+]ray-march -> /max ->
[resize] -> x255, RGB, -mean -> VGG fwd -> Gram/content/TV losses -> VGG dgrad -> ... -> d loss/d var
-> [all-reduce over ranks] -> Adam.
"""
import numpy as np
import torch

from . import _lib, ops
from .styler_base import StylerBase, f32
from .transform import rot_mat
from .util import octave_sizes


_GRAPH_POOLS = {}       # device index -> (pool handle, keeper graph, its buffer): see Styler._graph_pool
_CAPTURE_STREAMS = {}   # device index -> the capture stream


class _Adam:
    """Slots of one tf.compat.v1.train.AdamOptimizer (styler_3p.py:320-323): m, v per variable
    and the fp32 beta-power accumulators, kept across frames of the group and across octaves.
    The accumulators live on the device (``state`` = {beta1^t, beta2^t, lr_t}) so that a step
    replays from a CUDA graph."""

    def __init__(self):
        self.m = self.v = self.state = None

    def step(self, var, grad, lr, gscale=1.0):
        if self.m is None:
            self.m, self.v = torch.zeros_like(var), torch.zeros_like(var)
            self.state = torch.tensor([0.9, 0.999, 0.0], dtype=f32).to(var.device)
        ops.adam_step_dev(var, grad, self.m, self.v, self.state, lr, gscale)

    def iterate(self, g_opt, grad, lr, gscale, mask, mask_stride, apply):
        """The fused single-step iteration (``lnst_adam_iterate_dev``): returns (var, delta)."""
        if self.m is None:
            self.m, self.v = torch.zeros_like(g_opt), torch.zeros_like(g_opt)
            self.state = torch.tensor([0.9, 0.999, 0.0], dtype=f32).to(g_opt.device)
        return ops.adam_iterate_dev(g_opt, grad, self.m, self.v, self.state, lr, gscale, mask, mask_stride,
                                    torch.empty_like(g_opt), torch.empty_like(g_opt), apply)


class StepRunner:
    """One frame's loop body (``Styler.frame_step``) as a callable.  The first call runs eagerly
    (it fills the per-frame caches and allocates the Adam slots); the second call captures the same
    launch sequence into a CUDA graph, and every later call is a single graph replay -- the
    reference pays three ``sess.run`` round trips here (styler_3p.py:312,331,334).
    Returns (var, loss, delta): static buffers, overwritten by the next call."""

    def __init__(self, styler, fr, g_opt_t, adam, ws, style_grams, lr, use_graph=True):
        self.args = (fr, g_opt_t, adam, ws, style_grams, lr)
        self.styler, self.calls, self.graph, self.out, self.n_abi = styler, 0, None, None, 0
        self.use_graph = bool(use_graph) and styler.device.type == 'cuda'

    def __call__(self):
        st = self.styler
        if not self.use_graph or self.calls < 1:
            self.calls += 1
            out = st.frame_step(*self.args)
        else:
            lib = _lib.get()
            if self.graph is None:
                n0 = lib.launches
                # capture_begin / capture_end directly: the torch.cuda.graph context manager empties the caching
                # allocator (device and pinned host) before every capture, and the cudaFree / cudaMalloc round trips
                # of the blocks this run keeps using cost 10-200 ms per capture -- more than the 20 iterations of a
                # reference-sized run (tools/dev/run_phases.py)
                g = torch.cuda.CUDAGraph()
                pool = st._graph_pool()
                cap = st._capture_stream()
                cap.wait_stream(torch.cuda.current_stream(st.device))
                with torch.cuda.stream(cap):
                    g.capture_begin(pool)
                    try:
                        self.out = st.frame_step(*self.args)
                    finally:
                        g.capture_end()
                torch.cuda.current_stream(st.device).wait_stream(cap)
                self.n_abi, lib.launches, self.graph = lib.launches - n0, n0, g
            self.graph.replay()
            lib.launches += self.n_abi
            out = self.out
        st._advance_views()
        return out


class Styler(StylerBase):
    def __init__(self, self_dict, weights=None, device=None, content_weights=None):
        StylerBase.__init__(self, self_dict, weights=weights, device=device, content_weights=content_weights)
        if self.batch_size != 1:
            if self.rotate:
                raise NotImplementedError('batch_size > 1 with rotate: the reference itself breaks there (styler_3p.py:417 '
                                          'feeds batch_size matrices and rotate() tiles the batch by them)')
            if self.style_mask:
                raise NotImplementedError('batch_size > 1 with style_mask')
        if self.target_field not in ('d', 'p'):
            raise ValueError("styler_3p handles target_field 'd' or 'p'")
        if 'd' in self.target_field and self.num_kernels > 4:
            raise NotImplementedError('num_kernels > 4')
        if self.style_mask:                                        # styler_base.py:165-169 with d_gray = the render
            if 'vgg' not in self.model_path:
                raise NotImplementedError('style_mask with a GraphDef loss network')
            # the masked areas (Gram denominators 2 C * area) stay on the device (lnst_gram_diff_dev, lnst_scale_by_dev):
            # the step replays from a CUDA graph like the unmasked one
        self.rot_mat_, self.views = None, None
        if self.rotate:                                            # styler_3p.py:137-145
            self.rot_mat_, self.views = rot_mat(self.phi0, self.phi1, self.phi_unit, self.theta0, self.theta1,
                                                self.theta_unit, sample_type=self.sample_type, rng=self.rng,
                                                nv=self.n_views)
            if self.n_views is None:
                self.n_views = len(self.views)
            assert self.n_views % self.v_batch == 0
            if self.v_batch != 1:
                # a group of v_batch views per Adam step: joint normalisation, Gram loss on the group's first image
                # only (styler_base.py:98 loops over range(batch_size)), content / TV averaged over the group
                if self.view_mode != 'sequential':
                    raise NotImplementedError("v_batch > 1 is the reference's per-group Adam loop: view_mode='sequential'")
                if self.style_mask:
                    raise NotImplementedError('v_batch > 1 with style_mask')
        self._frame_cache = {}
        self._iv_cache = {}
        self._pool = None
        self._rot_all = self._rot_mine = None
        # what the ranks split: 'views' (mean-gradient mode), 'frames' (sequences) or
        # nothing (replicas); decided per run() by _set_shard
        self.view_rank, self.view_world = (self.rank, self.world) if self.view_mode == 'allreduce' else (0, 1)
        self._eye = self._rot_tensor([np.identity(3)])
        if self.rotate:
            self._upload_views()

    def _graph_pool(self):
        """One graph memory pool per device for the whole process, kept alive by a one-kernel graph: the blocks of a
        finished run's step graphs go back to this pool and the next run's captures take them from there.  With a pool per
        Styler every run paid cudaMalloc for ~1.5 GB of step temporaries inside its capture (20-50 ms at C3, more than
        the run's 20 replayed iterations) and left the old pool's memory parked until the allocator ran dry."""
        if self._pool is None:
            key = self.device.index if self.device.index is not None else torch.cuda.current_device()
            ent = _GRAPH_POOLS.get(key)
            if ent is None:
                handle = torch.cuda.graph_pool_handle()
                keeper = torch.cuda.CUDAGraph()
                cap = self._capture_stream()
                cap.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(cap):
                    keeper.capture_begin(handle)
                    buf = torch.zeros(1, dtype=f32, device=self.device)
                    keeper.capture_end()
                torch.cuda.current_stream(self.device).wait_stream(cap)
                ent = _GRAPH_POOLS[key] = (handle, keeper, buf)
            self._pool = ent[0]
        return self._pool

    def _capture_stream(self):
        """The side stream every capture of this process runs on (the caching allocator only hands a freed block to
        requests of the stream it was allocated on: a stream per Styler would strand the pool's blocks)."""
        key = self.device.index if self.device.index is not None else torch.cuda.current_device()
        if key not in _CAPTURE_STREAMS:
            _CAPTURE_STREAMS[key] = torch.cuda.Stream(self.device)
        return _CAPTURE_STREAMS[key]

    def _upload_views(self):
        """Host view matrices -> the persistent device buffers the kernels (and graphs) read."""
        self._iv_cache = {}                                        # ray intervals belong to the old views
        allv = self._rot_tensor(self.rot_mat_)
        mine = self._rot_tensor([self.rot_mat_[i] for i in range(self.view_rank, self.n_views, self.view_world)]) \
            if self.n_views > self.view_rank else None
        if self._rot_all is None or (mine is not None) != (self._rot_mine is not None) or \
                (mine is not None and mine.shape != self._rot_mine.shape):
            self._rot_all, self._rot_mine = allv, mine
        else:
            self._rot_all.copy_(allv)
            if mine is not None:
                self._rot_mine.copy_(mine)

    def set_world(self, rank, world):
        """Override the (rank, world) taken from torch.distributed at construction (tests)."""
        self.rank, self.world = rank, world
        self.view_rank, self.view_world = (rank, world) if self.view_mode == 'allreduce' else (0, 1)
        if self.rotate:
            self._upload_views()

    def _set_shard(self, nf):
        """Pick what this run splits over the ranks.  Views shard only in the mean-gradient mode (in the
        reference-exact sequential mode each view's Adam step depends on the previous one); frames
        shard whenever there are several (their only coupling is the temporal filter on the updates,
        one all-gather per iteration); otherwise every rank is a replica."""
        mode = getattr(self, 'shard', None)
        if mode is None:
            if self.world == 1:
                mode = 'none'
            elif self.rotate and self.view_mode == 'allreduce':
                mode = 'views'
            else:
                mode = 'frames' if nf > 1 else 'none'
        vr, vw = (self.rank, self.world) if mode == 'views' else (0, 1)
        if (vr, vw) != (self.view_rank, self.view_world):
            self.view_rank, self.view_world = vr, vw
            if self.rotate:
                self._upload_views()
        return mode

    def _frame_owners(self, key_frames):
        """rank owning each key frame: contiguous blocks of Adam groups (frames sharing one optimizer,
        styler_3p.py:315-323, are order-dependent and stay together)."""
        groups = sorted({t // self.frames_per_opt for t in key_frames})
        chunks = np.array_split(np.arange(len(groups)), self.world)
        owner_of_group = {groups[i]: r for r, c in enumerate(chunks) for i in c}
        return {t: owner_of_group[t // self.frames_per_opt] for t in key_frames}

    def _gather_frames(self, local, owners, keys, shape):
        """{t: tensor of ``shape``} held by the owners -> the same for every key frame on every rank
        (one all_gather of a buffer padded to the largest share)."""
        per_rank = [[t for t in keys if owners[t] == r] for r in range(self.world)]
        most = max(len(x) for x in per_rank)
        buf = torch.zeros([most] + [int(v) for v in shape], dtype=f32, device=self.device)
        for j, t in enumerate(per_rank[self.rank]):
            buf[j].copy_(local[t])
        out = [torch.empty_like(buf) for _ in range(self.world)]
        torch.distributed.all_gather(out, buf)
        return {t: out[r][j] for r in range(self.world) for j, t in enumerate(per_rank[r])}

    def _filter_frames_alltoall(self, deltas, owners, key_frames, sigma):
        """The one exchange step of a frame-sharded sequence (styler_3p.py:382-383) as two all-to-all transposes
        (SURVEY.md 8e): frame-sharded updates [T/W, N, c] -> particle-sharded [T, N/W, c], Gaussian along T on the
        local particle slice, and back.  Each rank moves (W-1)/W of its own share twice -- the all-gather variant makes
        every rank receive and filter the whole [T, N, c] stack.  Returns {t: filtered update} for this rank's frames,
        or None when the frames do not split evenly (the caller then all-gathers)."""
        W, rk = self.world, self.rank
        per_rank = [[t for t in key_frames if owners[t] == r] for r in range(W)]
        tm = len(per_rank[0])
        if tm == 0 or any(len(x) != tm for x in per_rank) or [t for x in per_rank for t in x] != list(key_frames):
            return None
        mine = per_rank[rk]
        n, c = deltas[mine[0]].shape
        nw = (n + W - 1) // W
        stack = torch.zeros(tm, W * nw, c, dtype=f32, device=self.device)
        for j, t in enumerate(mine):
            stack[j, :n].copy_(deltas[t])
        send = stack.view(tm, W, nw, c).permute(1, 0, 2, 3).contiguous()          # [W, tm, nw, c]: block w -> rank w
        recv = torch.empty_like(send)
        torch.distributed.all_to_all_single(recv, send)                            # block w = rank w's frames, my particles
        sm = ops.temporal_gauss(recv.view(W * tm, nw, c), sigma).view(W, tm, nw, c)
        back = torch.empty_like(sm)
        torch.distributed.all_to_all_single(back, sm.contiguous())                 # block w = my frames, rank w's particles
        out = back.permute(1, 0, 2, 3).reshape(tm, W * nw, c)
        return {t: out[j, :n].contiguous() for j, t in enumerate(mine)}

    def _advance_views(self):
        """Poisson-disc view sets are re-drawn after every frame pass (styler_3p.py:344-349)."""
        if self.rotate and 'uniform' not in self.sample_type:
            self.rot_mat_, self.views = rot_mat(self.phi0, self.phi1, self.phi_unit, self.theta0, self.theta1,
                                                self.theta_unit, sample_type=self.sample_type, rng=self.rng,
                                                nv=self.n_views)
            self._upload_views()

    def step_runner(self, fr, g_opt_t, adam, ws, style_grams, lr):
        return StepRunner(self, fr, g_opt_t, adam, ws, style_grams, lr,
                          use_graph=getattr(self, 'cuda_graphs', True))

    # ---- geometry ------------------------------------------------------------------------------
    def _grid(self, res):
        return _lib.make_grid(3, res, self.domain, self.nsize, self.clip)

    def _supports(self):
        if 'd' in self.target_field:                               # styler_3p.py:80-81
            return [self.radius * self.support / (self.kernel_scale ** k) for k in range(self.num_kernels)]
        return [self.radius * self.support]

    def _rot_tensor(self, mats):
        return torch.tensor(np.asarray(mats, dtype=np.float64).reshape(-1, 9), dtype=f32, device=self.device)

    # ---- forward pieces ------------------------------------------------------------------------
    def _density(self, fr, var, res, ws):
        """var -> density field d [D,H,W] (styler_3p.py:49-91)."""
        grid = ws['grid']
        if 'd' in self.target_field:
            lists = self._cell_lists(fr, res, grid, ws['d'])
            if lists is not None:                                   # gather over per-cell lists, TMA-stored tiles
                ops.splat_wavg_fwd_gather(lists, fr['r'], var, grid, self._supports(), ws['d'], ws['box'])
                return ws['d']
            wmap = self._wmap(fr, res, grid)
            ops.splat_wavg_fwd(fr['p'], fr['r'], var, grid, self._supports(), wmap, ws['num'], ws['d'], ws['box'])
        else:
            scale = 0.8 * (2 * self.radius) ** 3 * self.rest_density / self.rest_density
            ops.splat_sph_fwd(fr['p'], var, grid, self._supports()[0], scale, out=ws['d'])
        return ws['d']

    def _wmap(self, fr, res, grid):
        """Sum of splat weights per kernel and cell; positions are constants in density mode
        (styler_3p.py:60-76), so it is computed once per (frame, octave)."""
        key = (fr['id'], tuple(res))
        if key not in self._frame_cache:
            self._frame_cache[key] = ops.splat_wavg_wmap(fr['p'], grid, self._supports())
        return self._frame_cache[key]

    def _cell_lists(self, fr, res, grid, out):
        """Per-cell particle lists of a frame (positions are constants in density mode): built once per (frame, octave)
        for the gather splat; None when that kernel does not apply (then the scatter kernel runs).  Opt-in
        (``gather_splat = True``): the first version walks the lists straight from global memory and is latency-bound --
        0.33 ms against the scatter kernel's 0.095 ms at C3 on the B200."""
        if not (getattr(self, 'gather_splat', False) and self.nsize == 1 and not self.clip and ops._tma_ok(out)):
            return None
        key = (fr['id'], tuple(res), 'cells')
        if key not in self._frame_cache:
            self._frame_cache[key] = ops.cell_lists(fr['p'], grid) if fr['p'].shape[0] else None
        return self._frame_cache[key]

    def _coef(self, fr, res, grid):
        """d out/d num per kernel and cell (1/wmap with TF's where/div NaN rule): constant like wmap."""
        key = (fr['id'], tuple(res), 'coef')
        if key not in self._frame_cache:
            self._frame_cache[key] = ops.splat_wavg_coef(self._wmap(fr, res, grid))
        return self._frame_cache[key]

    def _workspace(self, res, frames=None):
        """Per-octave volumes.  With ``frames`` (density mode) the active box is set: the bounding box
        of every cell any frame's particles can reach (sum of weights > 0), grown by one voxel for the
        3x3x3 blur.  Density and the needed gradients are zero / unused outside it, so the volume
        kernels only visit the box; the volumes are zero-initialised and stay zero elsewhere."""
        D, H, W = res
        dev = self.device
        nk = self.num_kernels if 'd' in self.target_field else 1
        self._iv_cache = {}                                        # ray intervals belong to the old workspace's bricks
        ws = {'grid': self._grid(res), 'res': res, 'box': None, 'bricks': None, 'touch': None,
              'num': torch.zeros(nk, D * H * W, dtype=f32, device=dev),
              'd': torch.zeros(D, H, W, dtype=f32, device=dev),
              'ds': torch.zeros(D, H, W, dtype=f32, device=dev),
              'g_ds': torch.zeros(D, H, W, dtype=f32, device=dev),
              'g_d': torch.zeros(D, H, W, dtype=f32, device=dev)}
        if frames is not None and 'd' in self.target_field and getattr(self, 'active_box', True):
            occ = None
            for fr in frames:
                o = (self._wmap(fr, res, ws['grid']) > 0).any(0).reshape(D, H, W)
                occ = o if occ is None else (occ | o)
            lo, hi = [0, 0, 0], [0, 0, 0]
            if occ is not None and bool(occ.any()):
                for a, other in enumerate(((1, 2), (0, 2), (0, 1))):
                    idx = torch.nonzero(occ.any(other[1]).any(other[0])).flatten()
                    lo[a] = max(int(idx.min()) - 1, 0)
                    hi[a] = min(int(idx.max()) + 1, res[a] - 1)
            ws['box'] = _lib.make_box(lo, hi)
            if occ is not None and min(res) >= 8:
                # occupancy bricks for the ray-march (4^3 voxels): active = within one voxel of a reachable
                # cell (the blur); marked = any voxel within two voxels of an active one, then one more brick
                mp = torch.nn.functional.max_pool3d
                o = occ.to(f32)[None, None]
                o = mp(o, 3, 1, 1)                                        # active voxels
                o = mp(o, 5, 1, 2)                                        # + two voxels
                o = mp(o, 4, 4, 0, ceil_mode=True)                        # bricks
                o = mp(o, 3, 1, 1)                                        # + one brick
                ws['bricks'] = (o[0, 0] > 0).to(torch.uint8).contiguous()
                # exact footprint mask for fixed view sets: anchor voxel v is marked when any voxel of {v, v+1}^3 lies
                # within one voxel (the blur) of a reachable cell
                a = mp(occ.to(f32)[None, None], 3, 1, 1)
                a = mp(torch.nn.functional.pad(a, (0, 1, 0, 1, 0, 1)), 2, 1, 0)
                ws['touch'] = (a[0, 0] > 0).to(torch.uint8).contiguous()
            ws['box_cells'] = (hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1)
        return ws

    def _render(self, ds, rot, box=None, bricks=None, net_input=True, joint=False, touch=None, glue=None):
        """ds [D,H,W] -> gray [nv,H,W,1] in [0,1] plus what the backward needs.  ``net_input=False``: the loss
        net starts from the gray image itself (``_gray_path``), d_img / x are not produced."""
        D, H, W = ds.shape
        nv = 1 if rot is None else rot.shape[0]
        dev = self.device
        img = torch.empty(nv, H, W, dtype=f32, device=dev)
        stot = torch.empty(nv, H, W, dtype=f32, device=dev)
        iv = None
        if bricks is not None and rot is not None and min(ds.shape) >= 2:
            # per view set and independent of the density: computed once while the views are fixed
            key = (rot.data_ptr(), tuple(ds.shape), bricks.data_ptr())
            if 'uniform' in self.sample_type and key in self._iv_cache:
                iv = self._iv_cache[key]
            else:
                if 'uniform' in self.sample_type and touch is not None and getattr(self, 'exact_intervals', True):
                    iv = ops.ray_intervals_exact(rot, ds.shape, box, touch)   # fixed views: pay one march, once
                else:
                    iv = ops.ray_intervals(rot, ds.shape, box, bricks)
                if 'uniform' in self.sample_type:
                    self._iv_cache[key] = iv
        fused = glue is not None and not self.render_liquid and not joint
        ops.raymarch_fwd(ds, rot, self.transmit, self.render_liquid, img, stot, box, iv, stats=glue[0] if fused else None)
        st = {'img': img, 'stot': stot, 'rot': rot, 'box': box, 'iv': iv}
        if self.render_liquid:
            gray = img
        elif fused:
            # ``glue`` = (stats [2 nv], dots [nv]), zeroed by the caller: the march reduced the per-view maxima, one pass
            # normalises and counts the ties, and the backward folds the rest into its neighbours (_render_bwd)
            st['stats'], st['dots'], st['joint'] = glue[0], glue[1], False
            gray = ops.normalize_ties_fwd(img, st['stats'], torch.empty_like(img))
        else:                                                     # styler_3p.py:158
            # `d /= tf.reduce_max(d)` is over the whole fed tensor: one maximum per view here, or -- ``joint``, a
            # v_batch group -- one for all views of the call (the views seen as a single nv*H x W image)
            imj = img.reshape(1, nv * H, W) if joint else img
            st['stats'] = ops.image_max(imj, torch.empty(2 * imj.shape[0], dtype=f32, device=dev))
            gray = ops.normalize_fwd(imj, st['stats'], torch.empty_like(imj))
            st['joint'] = joint
        gray = gray.reshape(nv, H, W, 1)
        st['gray0'] = gray.reshape(nv, H, W)                      # self.d_gray (styler_3p.py:161): before any resize
        nh, nw = self._net_hw((H, W))
        if (nh, nw) != (H, W):                                    # styler_base.py:35-38
            gray = ops.resize_bilinear_fwd(gray, nh, nw)
        st.update(gray=gray.reshape(nv, nh, nw), hw=(H, W))
        if net_input:
            d_img = torch.empty(nv, nh, nw, 3, dtype=f32, device=dev)
            x = torch.empty(nv, nh, nw, 3, dtype=f32, device=dev)
            ops.to_net_input_fwd(gray, 255.0, d_img, x)           # styler_base.py:41-45, vgg.py:50-53
            st.update(d_img=d_img, x=x)
        return st

    def _render_bwd(self, st, g_x, ds, g_ds, g_gray0=None):
        """d loss / d x -> accumulated into g_ds (which the caller zeroed).  ``g_gray0`` [nv,H,W]: extra cotangent
        of the un-resized gray render (style mask)."""
        nv = g_x.shape[0]
        H, W = st['hw']
        if g_x.dim() == 3:                                        # already d loss / d gray (gray path)
            g_gray = g_x.reshape(nv, g_x.shape[1], g_x.shape[2], 1)
        else:
            g_gray = ops.to_net_input_bwd(g_x, 1, 255.0, torch.empty(nv, g_x.shape[1], g_x.shape[2], 1, dtype=f32,
                                                                      device=self.device))
        if (g_x.shape[1], g_x.shape[2]) != (H, W):
            g_gray = ops.resize_bilinear_bwd(g_gray, H, W)
        g_gray = g_gray.reshape(nv, H, W)
        if g_gray0 is not None:
            g_gray = ops.axpy(g_gray.contiguous(), g_gray0, 1.0)
        tc = getattr(self.net, 'tc', None)
        if self.render_liquid:
            g_img = g_gray
        elif st.get('dots') is not None and tc is not None and tc.gray_dot_done and g_gray0 is None:
            # sum(g_gray * img) came out of conv1_1's data-gradient kernel; the march applies the normalisation's gradient
            tc.gray_dot_done = False
            ops.raymarch_bwd(ds, st['rot'], self.transmit, False, st['stot'], g_gray.contiguous(), g_ds, st['box'], st['iv'],
                             norm=(st['img'], st['stats'], st['dots']))
            return
        else:
            imj, ggj = st['img'], g_gray.contiguous()
            if st.get('joint'):
                imj, ggj = imj.reshape(1, nv * H, W), ggj.reshape(1, nv * H, W)
            g_img = ops.normalize_bwd(imj, st['stats'], ggj, torch.empty(imj.shape[0], dtype=f32, device=self.device),
                                      torch.empty_like(ggj)).reshape(nv, H, W)
        ops.raymarch_bwd(ds, st['rot'], self.transmit, self.render_liquid, st['stot'], g_img, g_ds, st['box'], st['iv'])

    def _gray_path(self):
        """The render is gray and the tensor-core loss net can take it directly: conv1_1 folds the x255, the RGB
        replication and the mean subtraction into its weights (no TV loss, which reads d_img)."""
        wanted = self._wanted()
        if self.w_tv or not self.net.gray_path() or 'input' in wanted or not getattr(self, 'gray_conv', True) or \
                (self.net2 is not None and self.w_content):       # the second network reads the RGB net input
            return False
        pre = self.net.prefix(wanted)
        return bool(pre) and pre[0] == 'conv1_1'

    # ---- one loss + gradient evaluation (= one sess.run([train_op, total_loss]) without Adam) ----
    def loss_and_grad(self, fr, var, ws, rot, style_grams, group=False, grad_out=None):
        """Sum over the given views of total_loss, and d(sum)/d var.  Returns (loss [nv], grad).  ``group``: the views
        are ONE fed batch of the reference graph (v_batch > 1): joint normalisation and the group loss weights."""
        res = ws['res']
        nvtx = _lib.nvtx
        with nvtx('lnst.splat_fwd'):
            d = self._density(fr, var, res, ws)
        box = ws['box']
        with nvtx('lnst.smooth_fwd'):
            ds = ops.smooth3_relu_fwd(d, ws['ds'], self.k, box)    # styler_3p.py:112-125
        gray_path = self._gray_path()
        nv = 1 if rot is None else rot.shape[0]
        # one zeroed block for the step's small accumulators: loss [nv] | per-view {max, ties} [2 nv] | dots [nv]
        fuse = getattr(self, 'fuse_glue', True) and not self.render_liquid and not group
        scal = ops.zeros(4 * nv if fuse else nv, self.device)
        with nvtx('lnst.render_fwd'):
            st = self._render(ds, rot, box, ws['bricks'], net_input=not gray_path, joint=group, touch=ws['touch'],
                              glue=(scal[nv:3 * nv], scal[3 * nv:]) if fuse else None)
        assert nv == st['gray'].shape[0]
        loss = scal[:nv]
        g_gray0 = None
        tc = getattr(self.net, 'tc', None)
        if tc is not None:
            tc.gray_dot, tc.gray_dot_done = None, False
        if gray_path:
            if st.get('dots') is not None and tc is not None and tuple(st['gray'].shape[1:]) == tuple(st['hw']):
                tc.gray_dot = (st['img'], st['dots'])
            g_x = self.image_loss_and_grad(None, None, style_grams, loss, gray=st['gray'])
        elif self.style_mask and self.w_style and style_grams is not None:
            # mask per style layer = TF-legacy bicubic resize of the render to the feature size; its cotangent comes
            # back through the same resize and joins the render's gradient
            masks, mg, side = self.style_masks_for(st['gray0'], (st['x'].shape[1], st['x'].shape[2]),
                                                   device_areas=getattr(self, 'mask_areas_on_device', True)), {}, None
            if self.style_mask_on_ref:                             # :171-173: the target is masked by the same render
                by_layer, side = self.masked_style_grams(masks, per_image=True)
                style_grams = [by_layer[l] for l in self.style_layer]
            g_x = self.image_loss_and_grad(st['x'], st['d_img'], style_grams, loss, style_masks=masks, mask_grads=mg,
                                           style_side=side)
            H, W = st['hw']
            for l, dm in mg.items():
                gm = ops.resize_bicubic_bwd(dm.reshape(nv, dm.shape[1], dm.shape[2], 1), H, W).reshape(nv, H, W)
                g_gray0 = gm if g_gray0 is None else ops.axpy(g_gray0, gm, 1.0)
        else:
            g_x = self.image_loss_and_grad(st['x'], st['d_img'], style_grams, loss, group=group)
        with nvtx('lnst.render_bwd'):
            # (clearing g_ds on a forked graph branch under the loss network was tried: +5..15 us per step)
            g_ds = ops.fill_box(ws['g_ds'], box, 0.0)
            self._render_bwd(st, g_x, ds, g_ds, g_gray0)
        with nvtx('lnst.smooth_bwd'):
            g_d = ops.smooth3_relu_bwd(g_ds, ds, ws['g_d'], self.k, box)
        n_terms = 1 if group else nv                               # field / variable terms: once per fed batch
        if self.w_pressure > 0 and 'p' in self.target_field:       # styler_3p.py:96-98, styler_base.py:228-230
            ops.pressure_reg(d, 1.0, self.w_pressure, n_terms * self.w_pressure * 2.0 / d.numel(), loss, n_terms, g_d)
        if 'd' in self.target_field:
            grad = grad_out if grad_out is not None else torch.empty_like(var)
            if self.nsize == 1:
                ops.splat_wavg_bwd_coef(fr['p'], var, ws['grid'], self._supports(), self._coef(fr, res, ws['grid']),
                                        g_d, grad)
            else:
                ops.splat_wavg_bwd(fr['p'], var, ws['grid'], self._supports(), self._wmap(fr, res, ws['grid']), g_d,
                                   grad)
            if self.w_density > 0:                                 # styler_base.py:217-223
                ops.density_reg(var, self.w_density, n_terms * self.w_density, loss, n_terms, grad)
        else:
            scale = 0.8 * (2 * self.radius) ** 3 * self.rest_density / self.rest_density
            grad = ops.splat_sph_bwd_pos(fr['p'], var, ws['grid'], self._supports()[0], scale, g_d)
        return loss, grad

    def infer(self, fr, var, ws, identity_view):
        """Forward only: (p_out, d_out [D,H,W], d_img [H',W',3]) -- styler_3p.py:409-431."""
        d = self._density(fr, var, ws['res'], ws)
        ds = ops.smooth3_relu_fwd(d, ws['ds'], self.k, ws['box'])
        rot = self._eye if identity_view else None
        st = self._render(ds, rot, ws['box'], ws['bricks'], touch=ws['touch'])
        p_out = fr['p'] + var if 'p' in self.target_field else fr['p']
        return p_out, ds + 0.0, st['d_img'][0]                     # "+0.0" folds the -0.0 markers

    # ---- batch_size > 1 (styler_3p.py:42, 304-363, 409-431), rotate off ----------------------------------------
    def _render_batch(self, ds_list, ws):
        """B smoothed fields -> net input of the fed batch.  `d /= tf.reduce_max(d)` (:158) is ONE maximum over all B
        renders: the frames of a batch are coupled through it (forward value and, via the ties rule, gradient)."""
        B = len(ds_list)
        D, H, W = ds_list[0].shape
        dev = self.device
        img = torch.empty(B, H, W, dtype=f32, device=dev)
        stot = torch.empty(B, H, W, dtype=f32, device=dev)
        for i, ds in enumerate(ds_list):
            ops.raymarch_fwd(ds, None, self.transmit, self.render_liquid, img[i:i + 1], stot[i:i + 1], ws['box'])
        st = {'img': img, 'stot': stot, 'hw': (H, W)}
        if self.render_liquid:
            gray = img
        else:
            imj = img.reshape(1, B * H, W)
            st['stats'] = ops.image_max(imj, torch.empty(2, dtype=f32, device=dev))
            gray = ops.normalize_fwd(imj, st['stats'], torch.empty_like(imj))
        gray = gray.reshape(B, H, W, 1)
        nh, nw = self._net_hw((H, W))
        if (nh, nw) != (H, W):
            gray = ops.resize_bilinear_fwd(gray, nh, nw)
        d_img = torch.empty(B, nh, nw, 3, dtype=f32, device=dev)
        x = torch.empty(B, nh, nw, 3, dtype=f32, device=dev)
        ops.to_net_input_fwd(gray, 255.0, d_img, x)
        st.update(d_img=d_img, x=x)
        return st

    def _fields_batch(self, frs, var_list, ws):
        """(d_i, smoothed d_i) per frame, each in its own buffer (the workspace volumes are per-call scratch)."""
        d_list, ds_list = [], []
        for fr, var in zip(frs, var_list):
            d = self._density(fr, var, ws['res'], ws).clone()
            ds_list.append(ops.smooth3_relu_fwd(d, torch.zeros_like(d), self.k, ws['box']))
            d_list.append(d)
        return d_list, ds_list

    def loss_and_grad_batch(self, frs, var_list, ws, style_grams):
        """One `sess.run([train_op, total_loss])` of a fed batch: joint loss (scalar tensor) and d loss / d var_i.
        Gram terms are sums over the batch, content / TV / pressure are means over it (styler_base.py:127-230)."""
        B = len(frs)
        d_list, ds_list = self._fields_batch(frs, var_list, ws)
        st = self._render_batch(ds_list, ws)
        loss = torch.zeros(B, dtype=f32, device=self.device)
        g_x = self.image_loss_and_grad(st['x'], st['d_img'], style_grams, loss, share=1.0 / B)
        H, W = st['hw']
        g_gray = ops.to_net_input_bwd(g_x, 1, 255.0, torch.empty(B, g_x.shape[1], g_x.shape[2], 1, dtype=f32,
                                                                 device=self.device))
        if (g_x.shape[1], g_x.shape[2]) != (H, W):
            g_gray = ops.resize_bilinear_bwd(g_gray, H, W)
        g_gray = g_gray.reshape(B, H, W).contiguous()
        if self.render_liquid:
            g_img = g_gray
        else:
            g_img = ops.normalize_bwd(st['img'].reshape(1, B * H, W), st['stats'], g_gray.reshape(1, B * H, W),
                                      torch.empty(1, dtype=f32, device=self.device),
                                      torch.empty(1, B * H, W, dtype=f32, device=self.device)).reshape(B, H, W)
        total = loss.sum()
        extra = torch.zeros(1, dtype=f32, device=self.device)     # regulariser terms, accumulated by their kernels
        grads = []
        for i, (fr, var) in enumerate(zip(frs, var_list)):
            d, ds = d_list[i], ds_list[i]
            g_ds = ops.fill_box(ws['g_ds'], ws['box'], 0.0)
            ops.raymarch_bwd(ds, None, self.transmit, self.render_liquid, st['stot'][i:i + 1], g_img[i:i + 1].contiguous(),
                             g_ds, ws['box'])
            g_d = ops.smooth3_relu_bwd(g_ds, ds, ws['g_d'], self.k, ws['box'])
            if self.w_pressure > 0 and 'p' in self.target_field:   # reduce_mean over the [B,D,H,W,1] pressure tensor
                ops.pressure_reg(d, 1.0, self.w_pressure / B, self.w_pressure * 2.0 / (B * d.numel()), extra, 1, g_d)
            if 'd' in self.target_field:
                grad = torch.empty_like(var)
                if self.nsize == 1:
                    ops.splat_wavg_bwd_coef(fr['p'], var, ws['grid'], self._supports(),
                                            self._coef(fr, ws['res'], ws['grid']), g_d, grad)
                else:
                    ops.splat_wavg_bwd(fr['p'], var, ws['grid'], self._supports(), self._wmap(fr, ws['res'], ws['grid']),
                                       g_d, grad)
                if self.w_density > 0:                             # summed over the batch (styler_base.py:217-223)
                    ops.density_reg(var, self.w_density, self.w_density, extra, 1, grad)
            else:
                scale = 0.8 * (2 * self.radius) ** 3 * self.rest_density / self.rest_density
                grad = ops.splat_sph_bwd_pos(fr['p'], var, ws['grid'], self._supports()[0], scale, g_d)
            grads.append(grad)
        return total + extra[0], grads

    def _infer_batch(self, frs, var_list, ws):
        """forward only for a fed batch: (p_out_i, d_out_i, d_img_i) with the batch's joint normalisation"""
        _, ds_list = self._fields_batch(frs, var_list, ws)
        st = self._render_batch(ds_list, ws)
        return [((fr['p'] + var) if 'p' in self.target_field else fr['p'], ds + 0.0, st['d_img'][i])
                for i, (fr, var, ds) in enumerate(zip(frs, var_list, ds_list))]

    def _run_batched(self, params):
        """`run` for batch_size > 1: the reference's loop with B frames per `sess.run` (one process, eager launches --
        the single-frame path keeps the CUDA-graph / sharded loop)."""
        dev = self.device
        nf, B, itp = self.num_frames, self.batch_size, self.interp
        if any(t + (B - 1) * itp >= nf for t in range(0, nf, B * itp)) or nf % B:
            raise ValueError('the key frames must fill whole batches (the reference feeds p[t+i*interp], styler_3p.py:304-308, '
                             'and p[t+i] in the final pass, :409-412)')
        lr_list = None
        if abs(self.lr_scale - 1) > 1e-7:
            lr_list = [self.lr / self.lr_scale ** i for i in range(self.octave_n)]
        oct_size = octave_sizes(self.resolution, self.octave_n, self.octave_scale)
        frames, inv = self.upload(params)
        width = 3 if 'p' in self.target_field else self.num_kernels
        g_opt = [torch.zeros(fr['p'].shape[0], width, dtype=f32, device=dev) for fr in frames]
        key = list(range(0, nf, itp))
        mask_of = (lambda fr: fr['r']) if 'd' in self.target_field else (lambda fr: None)
        mstride = self.num_kernels if 'd' in self.target_field else 0
        loss_history, d_intm, opt_ = [], [], {}
        for octave in range(self.octave_n):
            res = oct_size[octave]
            self._frame_cache = {}
            ws = self._workspace(res, frames)
            style_grams = None
            if self.w_style and self.style_img is not None:
                style_grams = self._style_feature(self.style_img, res[1:])
            self._content_feat = None
            if self.w_content and self.content_img is not None:
                self._content_feat = self._content_feature(self.content_img, res[1:])
            lr = lr_list[octave] if lr_list is not None else (self.lr[octave] if isinstance(self.lr, list) else self.lr)
            loss_o, intm_o = [], []
            for step in range(self.iter):
                deltas = {}
                for t in range(0, nf, B * itp):
                    idx = [t + i * itp for i in range(B)]
                    frs = [frames[f] for f in idx]
                    var = [g_opt[f].clone() for f in idx]                            # :312
                    total, grads = self.loss_and_grad_batch(frs, var, ws, style_grams)
                    for i in range(B):                                               # one Adam op, a slot per position
                        opt_.setdefault((t // self.frames_per_opt, i), _Adam()).step(var[i], grads[i], lr)
                    loss_o.append(total)
                    for i, f in enumerate(idx):                                      # :359-363
                        deltas[f] = ops.iterate_delta(var[i], 1.0, g_opt[f], mask_of(frames[f]), mstride,
                                                      torch.empty_like(var[i]))
                    if step == self.iter - 1 and octave < self.octave_n - 1:         # :365-370
                        intm_o += [o[2] for o in self._infer_batch(frs, var, ws)]
                if self.window_sigma > 0 and nf > 1:                                 # :382-383
                    sm = ops.temporal_gauss(torch.stack([deltas[f] for f in key], 0), self.window_sigma)
                    for j, f in enumerate(key):
                        deltas[f] = sm[j]
                for f in key:                                                        # :385-386
                    ops.axpy(g_opt[f], deltas[f].contiguous(), 1.0)
            loss_history.append([float(v) for v in torch.stack(loss_o).cpu().tolist()] if loss_o else [])
            if octave < self.octave_n - 1:
                d_intm.append(torch.stack(intm_o, 0).cpu().numpy().astype(np.uint8))
        if itp > 1:                                                                  # :392-397
            w = np.linspace(0, 1, itp + 1)
            for t in range(0, nf - 1, itp):
                for i in range(1, itp):
                    g_opt[t + i] = g_opt[t] * float(1 - w[i]) + g_opt[t + itp] * float(w[i])
        result = {'l': loss_history, 'd_intm': d_intm, 'v': None, 'c': None}
        res = oct_size[-1]
        self._frame_cache = {}
        ws = self._workspace(res, frames)
        p_sty, v_sty, d_sty, r_sty = [], [], [], []
        for t in range(0, nf, B):                                                    # :409-431
            idx = list(range(t, min(t + B, nf)))
            outs = self._infer_batch([frames[f] for f in idx], [g_opt[f] for f in idx], ws)
            for f, (p_out, d_out, d_img) in zip(idx, outs):
                if inv is not None:
                    p_out, g_opt[f] = p_out[inv], g_opt[f][inv]
                p_sty.append(p_out.cpu().numpy())
                v_sty.append(g_opt[f].cpu().numpy())
                d_sty.append(d_out.cpu().numpy()[..., None])
                r_sty.append(d_img.cpu().numpy().astype(np.uint8))
        result['p'] = p_sty
        if 'p' in self.target_field:
            result['v'] = v_sty
        result['d'] = np.array(d_sty)
        result['r'] = np.array(r_sty)
        result['g_opt'] = [g.cpu().numpy() for g in g_opt]
        return result

    # ---- one pass of the loop body for one frame (styler_3p.py:304-363) --------------------------
    def frame_step(self, fr, g_opt_t, adam, ws, style_grams, lr):
        """var <- g_opt[t]; Adam step(s) over the views; returns (var, loss, delta) with
        delta = (nan_to_num(mean iterate) - g_opt[t]) [* r[:,0:1]].  Everything stays on device and the
        launch sequence is fixed, so ``StepRunner`` can capture it into a CUDA graph (the view set is
        re-drawn by ``_advance_views`` after the call, outside the graph)."""
        dev = self.device
        mask = fr['r'] if 'd' in self.target_field else None       # :359-363
        mstride = self.num_kernels if mask is not None else 0
        if self.rotate and self.view_mode == 'sequential':         # :329-340: one Adam step per view
            var = g_opt_t.clone()                                  # :312 (device copy, no H2D)
            n_step_views = self.n_views // self.v_batch
            acc = torch.empty_like(var)
            losses = torch.empty(n_step_views, dtype=f32, device=dev)
            for j, i in enumerate(range(0, self.n_views, self.v_batch)):
                l, grad = self.loss_and_grad(fr, var, ws, self._rot_all[i:i + self.v_batch], style_grams,
                                             group=self.v_batch > 1)
                adam.step(var, grad, lr)
                ops.iterate_accumulate(acc, var, i == 0)
                ops.sum_scale(l, 1.0, out=losses[j:j + 1])         # one total_loss per fed group
            loss_t = ops.sum_scale(losses, 1.0 / n_step_views)[0]  # :342 (library reductions: no framework kernels in the graph)
            delta = ops.iterate_delta(acc, 1.0 / n_step_views, g_opt_t, mask, mstride, torch.empty_like(var))  # :351-352
            return var, loss_t, delta
        # one Adam step per iteration: the forward/backward pass reads g_opt[t] itself (the reference's variable
        # holds exactly that value, :312), then ONE kernel does Adam + the iterate bookkeeping (+ g_opt += delta
        # when no temporal filter runs in between, ``self.fuse_apply``)
        gscale = 1.0
        if self.rotate:                                            # mean view gradient, views sharded over ranks
            # the gradient is written straight into the all-reduce buffer, the loss scalar rides in its last element
            buf = torch.empty(g_opt_t.numel() + 1, dtype=f32, device=dev) if self.view_world > 1 else None
            gview = buf[:-1].view_as(g_opt_t) if buf is not None and 'd' in self.target_field else None
            if self._rot_mine is not None:
                l, grad = self.loss_and_grad(fr, g_opt_t, ws, self._rot_mine, style_grams, grad_out=gview)
                lsum = ops.sum_scale(l, 1.0) if self.view_world > 1 else l      # one rank: summed and scaled in one launch below
            else:
                grad = gview.zero_() if gview is not None else torch.zeros_like(g_opt_t)
                lsum = torch.zeros(1, dtype=f32, device=dev)
            if self.view_world > 1:                                # ONE all-reduce: gradient + loss scalar
                if gview is None:
                    buf[:-1].copy_(grad.reshape(-1))
                buf[-1:].copy_(lsum)
                torch.distributed.all_reduce(buf)
                grad, lsum = buf[:-1].view_as(g_opt_t), buf[-1:]
            gscale = 1.0 / self.n_views
            loss_t = ops.sum_scale(lsum, 1.0 / self.n_views)[0]
        else:                                                      # :354-357
            l, grad = self.loss_and_grad(fr, g_opt_t, ws, None, style_grams)
            loss_t = l[0]
        var, delta = adam.iterate(g_opt_t, grad, lr, gscale, mask, mstride, getattr(self, 'fuse_apply', False))
        return var, loss_t, delta

    # ---- device residency -----------------------------------------------------------------------
    def upload(self, params):
        """Host particle lists -> device frames.  When every frame has the same particle count the
        particles are re-ordered ONCE by the linear index of their cell in frame 0 (same permutation
        for all frames, so per-particle state -- variables, Adam moments, the temporal filter --
        stays aligned); neighbouring threads then splat into / gather from neighbouring cells.
        Returns (frames, inverse permutation or None)."""
        dev = self.device
        nf = self.num_frames
        ps = [torch.as_tensor(np.asarray(params['p'][i]), dtype=f32).to(dev) for i in range(nf)]
        rs = None
        if 'd' in self.target_field:
            rs = [torch.as_tensor(np.asarray(params['r'][i]), dtype=f32).to(dev) for i in range(nf)]
        perm = inv = None
        if getattr(self, 'sort_particles', True) and len({p.shape[0] for p in ps}) == 1 and ps[0].shape[0] > 0:
            res = torch.tensor([float(v) for v in self.resolution], dtype=f32, device=dev)
            c = torch.floor(ps[0] * res).clamp_(min=-1)
            c = torch.minimum(c, res).to(torch.float64)             # out-of-domain / padding rows sort to the ends
            r64 = res.to(torch.float64)
            key = (c[:, 0] * (r64[1] + 2) + c[:, 1]) * (r64[2] + 2) + c[:, 2]
            perm = torch.argsort(key, stable=True)
            inv = torch.empty_like(perm)
            inv[perm] = torch.arange(perm.numel(), device=dev)
        frames = []
        for i in range(nf):
            fr = {'id': i, 'p': (ps[i][perm] if perm is not None else ps[i]).contiguous()}
            if rs is not None:
                fr['r'] = (rs[i][perm] if perm is not None else rs[i]).contiguous()
            frames.append(fr)
        return frames, inv

    # ---- the optimisation loop (styler_3p.py:229-438) ----------------------------------------------
    def run(self, params):
        if self.batch_size > 1:
            return self._run_batched(params)
        dev = self.device
        nf = self.num_frames
        lr_list = None
        if abs(self.lr_scale - 1) > 1e-7:                          # :237-238
            lr_list = [self.lr / self.lr_scale ** i for i in range(self.octave_n)]
        oct_size = octave_sizes(self.resolution, self.octave_n, self.octave_scale)

        frames, inv = self.upload(params)
        width = 3 if 'p' in self.target_field else self.num_kernels
        g_opt = [torch.zeros(fr['p'].shape[0], width, dtype=f32, device=dev) for fr in frames]
        eye = True if self.rotate else False

        key_frames = list(range(0, nf, self.batch_size * self.interp))
        shard = self._set_shard(nf)
        owners = self._frame_owners(key_frames) if shard == 'frames' else {t: self.rank for t in key_frames}
        mine = [t for t in key_frames if owners[t] == self.rank]

        # g_opt += delta happens inside the fused iteration kernel unless the temporal filter sits in between
        self.fuse_apply = not (self.window_sigma > 0 and nf > 1)
        loss_history, d_intm, opt_ = [], [], {}
        for octave in range(self.octave_n):
            res = oct_size[octave]
            self._frame_cache = {}
            ws = self._workspace(res, frames)
            style_grams = None
            if self.w_style and self.style_img is not None:        # :281-286
                style_grams = self._style_feature(self.style_img, res[1:])
            self._content_feat = None
            if self.w_content and self.content_img is not None:    # :276-279
                self._content_feat = self._content_feature(self.content_img, res[1:])
            lr = lr_list[octave] if lr_list is not None else (self.lr[octave] if isinstance(self.lr, list) else self.lr)
            loss_o, intm_o = {t: [] for t in mine}, {}
            runners = {}
            for step in range(self.iter):
                if getattr(self, 'iter_events', None) is not None and self.device.type == 'cuda':
                    ev = torch.cuda.Event(enable_timing=True)    # measurement hook (bench.py, tools/c4_sharded.py):
                    ev.record()                                  # one event at the start of every iteration
                    self.iter_events.append(ev)
                deltas = {}
                for t in mine:
                    fr = frames[t]
                    adam = opt_.setdefault(t // self.frames_per_opt, _Adam())   # :315-323
                    if t not in runners:
                        runners[t] = self.step_runner(fr, g_opt[t], adam, ws, style_grams, lr)
                    var, loss_t, deltas[t] = runners[t]()
                    loss_o[t].append(loss_t.clone())
                    if step == self.iter - 1 and octave < self.octave_n - 1:   # :365-370
                        _, _, d_img = self.infer(fr, var, ws, eye)
                        intm_o[t] = d_img
                if self.window_sigma > 0 and nf > 1:               # :382-383
                    done = None
                    if shard == 'frames' and getattr(self, 'frame_exchange', 'alltoall') == 'alltoall':
                        done = self._filter_frames_alltoall(deltas, owners, key_frames, self.window_sigma)
                    if done is not None:
                        deltas = done
                    else:
                        if shard == 'frames':                      # the one exchange step of a sharded sequence
                            deltas = self._gather_frames(deltas, owners, key_frames, g_opt[key_frames[0]].shape)
                        sm = ops.temporal_gauss(torch.stack([deltas[t] for t in key_frames], 0), self.window_sigma)
                        for j, t in enumerate(key_frames):
                            deltas[t] = sm[j]
                if not self.fuse_apply or (self.rotate and self.view_mode == 'sequential'):
                    for t in mine:                                 # :385-386
                        ops.axpy(g_opt[t], deltas[t].contiguous(), 1.0)
            hist = {t: torch.stack(v) for t, v in loss_o.items() if v}
            if shard == 'frames':
                hist = self._gather_frames(hist, owners, key_frames, (self.iter,)) if self.iter else {}
                if octave < self.octave_n - 1:
                    intm_o = self._gather_frames({t: v.to(f32) for t, v in intm_o.items()}, owners, key_frames,
                                                 tuple(self._net_hw(res[1:])) + (3,))
            # the reference appends one loss per (step, frame) in that order (:342,355,388)
            loss_history.append([float(v) for v in torch.stack([hist[t] for t in key_frames], 1).reshape(-1).cpu().tolist()]
                                if hist else [])
            if octave < self.octave_n - 1:
                d_intm.append(torch.stack([intm_o[t] for t in key_frames], 0).cpu().numpy().astype(np.uint8))
            runners.clear()                                        # graphs hold this octave's workspaces

        if shard == 'frames':                                      # every rank finishes with every frame's variables
            got = self._gather_frames({t: g_opt[t] for t in mine}, owners, key_frames, g_opt[key_frames[0]].shape)
            for t in key_frames:
                g_opt[t] = got[t].contiguous()

        if self.interp > 1:                                        # :392-397
            w = np.linspace(0, 1, self.interp + 1)
            for t in range(0, nf - 1, self.interp):
                for i in range(1, self.interp):
                    g_opt[t + i] = g_opt[t] * float(1 - w[i]) + g_opt[t + self.interp] * float(w[i])

        # final inference (:404-438)
        result = {'l': loss_history, 'd_intm': d_intm, 'v': None, 'c': None}
        res = oct_size[-1]
        if self.octave_n < 1 or self.iter < 1 or tuple(ws['res']) != tuple(res):
            self._frame_cache = {}                                 # else: the last octave's volumes, box and weight maps
            ws = self._workspace(res, frames)
        # results leave through ONE page-locked staging buffer (positions, variables, density and render of every
        # frame, asynchronous copies, one synchronisation); the arrays handed back are views of it
        D, H, W = res
        V = D * H * W
        psz = [int(frames[t]['p'].numel()) for t in range(nf)]
        vsz = [int(g_opt[t].numel()) for t in range(nf)]
        stage, views, img_shape = None, [], None
        for t in range(nf):
            p_out, d_out, d_img = self.infer(frames[t], g_opt[t], ws, eye)
            if inv is not None:                                    # back to the caller's particle order
                p_out, g_opt[t] = p_out[inv], g_opt[t][inv]
            if stage is None:
                img_shape = tuple(d_img.shape)
                isz = int(d_img.numel())
                stage = torch.empty(sum(psz) + sum(vsz) + nf * (V + isz), dtype=f32, pin_memory=(dev.type == 'cuda'))
                o_d, o_r = sum(psz) + sum(vsz), sum(psz) + sum(vsz) + nf * V
            o_p = sum(psz[:t]) + sum(vsz[:t])
            dst = (stage[o_p:o_p + psz[t]], stage[o_p + psz[t]:o_p + psz[t] + vsz[t]],
                   stage[o_d + t * V:o_d + (t + 1) * V], stage[o_r + t * isz:o_r + (t + 1) * isz])
            for x, y in zip(dst, (p_out, g_opt[t], d_out, d_img)):
                x.copy_(y.reshape(-1).to(f32), non_blocking=True)
            views.append((dst[0].numpy().reshape(tuple(p_out.shape)), dst[1].numpy().reshape(tuple(g_opt[t].shape))))
        if dev.type == 'cuda':
            torch.cuda.current_stream(dev).synchronize()
        result['p'] = [v[0] for v in views]
        if 'p' in self.target_field:
            result['v'] = [v[1] for v in views]
        if stage is not None:
            result['d'] = stage[o_d:o_d + nf * V].numpy().reshape(nf, D, H, W, 1)
            result['r'] = stage[o_r:o_r + nf * isz].numpy().reshape((nf,) + img_shape).astype(np.uint8)
        else:
            result['d'], result['r'] = np.array([]), np.array([])
        result['g_opt'] = [v[1] for v in views]                    # engine extra: final variables
        return result
