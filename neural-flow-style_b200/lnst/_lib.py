"""ctypes binding of ``liblnst_b200.so`` (the C-ABI declared in ``include/lnst_b200.h``).

This is the only place the native library is loaded.  There is NO fallback: if the CUDA
library is missing or no CUDA device is present the engine raises.  (Tests that exercise the
kernels through the CPU interpreter under ``tools/cpu_emu`` install it explicitly with
``set_for_testing``; nothing in the package ever does.)
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'liblnst_b200.so')

i32, i64, f32, vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class LnstGrid(C.Structure):
    _fields_ = [('dim', i32), ('res', i32 * 3), ('domain', f32 * 3), ('cell', f32), ('nsize', i32),
                ('clip', i32)]


class LnstBox(C.Structure):
    _fields_ = [('lo', i32 * 3), ('hi', i32 * 3)]


GP = C.POINTER(LnstGrid)
BP = C.POINTER(LnstBox)
FP = C.POINTER(f32)      # host float array
IP = C.POINTER(i32)      # host int array

# name -> argtypes (return type is always int)
SIGNATURES = {
    'lnst_abi_version': [],
    'lnst_splat_sph_fwd': [vp, vp, i64, GP, f32, f32, vp, vp, i32, f32, vp, vp],
    'lnst_splat_sph_bwd_pos': [vp, vp, i64, GP, f32, f32, vp, vp, vp],
    'lnst_splat_sph_bwd_color': [vp, i64, GP, f32, f32, vp, i32, f32, vp, vp, vp],
    'lnst_splat_wavg_wmap': [vp, i64, GP, FP, i32, vp, vp],
    'lnst_splat_wavg_fwd': [vp, vp, vp, i64, GP, FP, i32, vp, vp, vp, vp],
    'lnst_splat_wavg_fwd_box': [vp, vp, vp, i64, GP, FP, i32, vp, vp, vp, BP, vp],
    'lnst_splat_wavg_bwd': [vp, vp, i64, GP, FP, i32, vp, vp, vp, vp],
    'lnst_splat_wavg_coef': [vp, i32, i64, vp, vp],
    'lnst_splat_cells': [vp, i64, GP, vp, vp, vp],
    'lnst_splat_wavg_bwd_coef': [vp, vp, i64, GP, FP, i32, vp, vp, vp, vp],
    'lnst_smooth3_relu_fwd': [vp, vp, i32, i32, i32, i32, vp],
    'lnst_smooth3_relu_bwd': [vp, vp, vp, i32, i32, i32, i32, vp],
    'lnst_smooth3_relu_fwd_box': [vp, vp, i32, i32, i32, i32, BP, vp],
    'lnst_smooth3_relu_bwd_box': [vp, vp, vp, i32, i32, i32, i32, BP, vp],
    'lnst_fill_box': [vp, i32, i32, i32, BP, f32, vp],
    'lnst_rotate_fwd': [vp, vp, i32, i32, i32, i32, vp, vp],
    'lnst_raymarch_fwd': [vp, vp, i32, i32, i32, i32, f32, i32, vp, vp, vp],
    'lnst_raymarch_bwd': [vp, vp, i32, i32, i32, i32, f32, i32, vp, vp, vp, vp],
    'lnst_raymarch_fwd_box': [vp, vp, i32, i32, i32, i32, f32, i32, BP, vp, vp, vp, vp],
    'lnst_raymarch_bwd_box': [vp, vp, i32, i32, i32, i32, f32, i32, BP, vp, vp, vp, vp, vp],
    'lnst_raymarch_fwd_max_box': [vp, vp, i32, i32, i32, i32, f32, i32, BP, vp, vp, vp, vp, vp],
    'lnst_raymarch_bwd_norm_box': [vp, vp, i32, i32, i32, i32, f32, i32, BP, vp, vp, vp, vp, vp, vp, vp, vp],
    'lnst_normalize_ties_fwd': [vp, vp, i32, i64, vp, vp],
    'lnst_ray_intervals': [vp, i32, i32, i32, i32, BP, vp, vp, vp],
    'lnst_ray_intervals_exact': [vp, i32, i32, i32, i32, BP, vp, vp, vp],
    'lnst_set_raymarch_merge': [i32],
    'lnst_image_max': [vp, i32, i64, vp, vp],
    'lnst_normalize_fwd': [vp, vp, i32, i64, vp, vp],
    'lnst_normalize_bwd': [vp, vp, vp, i32, i64, vp, vp, vp],
    'lnst_resize_bilinear_fwd': [vp, i32, i32, i32, i32, i32, i32, vp, vp],
    'lnst_resize_bilinear_bwd': [vp, i32, i32, i32, i32, i32, i32, vp, vp],
    'lnst_resize_bicubic_fwd': [vp, i32, i32, i32, i32, i32, i32, vp, vp],
    'lnst_resize_bicubic_bwd': [vp, i32, i32, i32, i32, i32, i32, vp, vp],
    'lnst_rowdot': [vp, vp, i32, i64, vp, f32, i32, vp, vp],
    'lnst_to_net_input_fwd': [vp, i32, i64, i32, f32, vp, vp, vp],
    'lnst_to_net_input_bwd': [vp, i32, i64, i32, f32, vp, vp],
    'lnst_conv3x3_f32': [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp],
    'lnst_avgpool2_fwd': [vp, vp, i32, i32, i32, i32, vp],
    'lnst_avgpool2_bwd': [vp, vp, vp, i32, i32, i32, i32, vp],
    'lnst_gram_diff': [vp, i64, i32, f32, vp, f32, vp, vp, vp],
    'lnst_gram_diff_dev': [vp, i64, i32, f32, vp, f32, vp, f32, vp, vp, vp],
    'lnst_scale_by_dev': [vp, i64, f32, vp, f32, vp, vp],
    'lnst_gram_bwd': [vp, vp, i64, i32, f32, f32, i32, vp, vp],
    'lnst_content_loss': [vp, i64, i32, i32, f32, vp, vp, f32, i32, vp],
    'lnst_content_mse': [vp, vp, i64, f32, f32, vp, vp, f32, i32, vp],
    'lnst_tv_loss': [vp, i32, i32, i32, f32, vp, vp, vp],
    'lnst_adam_step': [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, vp],
    'lnst_adam_step_dev': [vp, vp, vp, vp, i64, vp, f32, f32, f32, f32, f32, vp],
    'lnst_adam_iterate_dev': [vp, vp, vp, vp, i64, vp, f32, f32, f32, f32, f32, vp, i32, i32, vp, vp, i32, vp],
    'lnst_iterate_accumulate': [vp, vp, i64, i32, vp],
    'lnst_iterate_delta': [vp, f32, vp, vp, i32, i32, i64, vp, vp],
    'lnst_masked_accumulate': [vp, vp, vp, i32, i32, f32, vp, i64, vp],
    'lnst_temporal_gauss': [vp, vp, i32, i64, f32, vp],
    'lnst_axpy': [vp, vp, f32, i64, vp],
    'lnst_sum_scale': [vp, i32, f32, vp, vp],
    'lnst_zero': [vp, i64, vp],
    'lnst_clip_fwd': [vp, f32, f32, vp, i64, vp],
    'lnst_clip_bwd': [vp, vp, f32, f32, f32, vp, i64, vp],
    'lnst_mul_bcast': [vp, vp, i32, vp, i64, vp],
    'lnst_advect': [vp, vp, i32, IP, i32, vp, vp],
    'lnst_g2p': [vp, i32, IP, i32, vp, vp, i64, i32, vp, vp],
    'lnst_rk4_advect': [vp, i32, IP, vp, i64, f32, i32, vp, vp, vp],
    'lnst_pressure_loss': [vp, i64, f32, f32, vp, vp, vp],
    'lnst_sub_fliph': [vp, vp, vp, i32, i32, i32, vp],
    'lnst_pressure_reg': [vp, i64, f32, f32, f32, vp, i32, vp, vp],
    'lnst_density_reg': [vp, i64, f32, f32, vp, vp, i32, vp, vp],
    'lnst_rotate_bwd': [vp, vp, i32, i32, i32, i32, vp, vp],
    'lnst_conv2d_f32': [vp, vp, vp, vp] + [i32] * 14 + [vp],
    'lnst_conv2d_bwd_data_f32': [vp, vp, i32, vp, vp] + [i32] * 13 + [vp],
    'lnst_relu_fwd': [vp, vp, i64, vp],
    'lnst_relu_bwd': [vp, vp, vp, i64, i32, vp],
    'lnst_maxpool_fwd': [vp, vp] + [i32] * 10 + [vp],
    'lnst_maxpool_bwd': [vp, vp, vp] + [i32] * 11 + [vp],
    'lnst_avgpool_fwd': [vp, vp] + [i32] * 10 + [vp],
    'lnst_avgpool_bwd': [vp, vp] + [i32] * 11 + [vp],
    'lnst_lrn_fwd': [vp, vp, i64, i32, i32, f32, f32, f32, vp],
    'lnst_lrn_bwd': [vp, vp, vp, i64, i32, i32, f32, f32, f32, i32, vp],
    'lnst_copy_channels': [vp, i32, vp, i32, i32, i64, i32, vp],
}
# entry points that only exist in the CUDA build (tcgen05 / TMA); filled in by conv_tc.cu
CUDA_ONLY = {
    'lnst_tc_supported': [],
    'lnst_set_conv_persistent': [i32],
    'lnst_set_conv_halo': [i32],
    'lnst_set_conv_first_col': [i32],
    'lnst_set_conv_first_mma': [i32],
    'lnst_set_gram_split3': [i32],
    'lnst_conv_first_bwd_tc': [vp, vp, vp, i32, i32, i32, vp],
    'lnst_conv_first_fwd_gray': [vp, vp, vp, vp, vp, i32, i32, i32, vp],
    'lnst_conv_first_bwd_gray_tc': [vp, vp, vp, i32, i32, i32, vp],
    'lnst_umma_probe': [vp, vp, vp, i32, i32, i32, i32, vp],
    'lnst_conv3x3_bf16_tc': [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp],
    'lnst_conv3x3_mixed': [vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
    'lnst_gram_diff_bf16_tc': [vp, i32, i64, i32, f32, vp, f32, vp, vp, vp, vp],
    'lnst_gram_bwd_bf16_tc': [vp, vp, f32, vp, i32, vp, i32, i32, i32, i32, vp],
    'lnst_conv_first_fwd': [vp, vp, vp, vp, i32, i32, i32, vp],
    'lnst_conv_first_bwd': [vp, vp, vp, i32, i32, i32, vp],
    'lnst_avgpool2_bf16_fwd': [vp, vp, i32, i32, i32, i32, vp],
    'lnst_avgpool2_bf16_bwd': [vp, vp, vp, i32, i32, i32, i32, vp],
    'lnst_conv3x3_bf16x3_tc': [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp],
    'lnst_conv3x3_pool_bf16x3_tc': [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp],
    'lnst_conv3x3_gram_bf16x3_tc': [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
    'lnst_conv3x3_unpool_bf16x3_tc': [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
    'lnst_gram_diff_scaled_bf16x3_tc': [vp, i32, i64, i32, f32, vp, f32, f32, vp, vp, vp, vp, vp],
    'lnst_gram_diff_bf16x3_tc': [vp, i32, i64, i32, f32, vp, f32, vp, vp, vp, vp, vp],
    'lnst_gram_bwd_bf16x3_tc': [vp, vp, f32, vp, i32, vp, i32, i32, i32, i32, vp],
    'lnst_conv_first_fwd_x3': [vp, vp, vp, vp, i32, i32, i32, vp],
    'lnst_conv_first_fwd_gray_x3': [vp, vp, vp, vp, vp, i32, i32, i32, vp],
    'lnst_conv_first_bwd_x3_tc': [vp, vp, vp, i32, i32, i32, vp],
    'lnst_conv_first_bwd_gray_x3_tc': [vp, vp, vp, i32, i32, i32, vp],
    'lnst_avgpool2_bf16x3_fwd': [vp, vp, i32, i32, i32, i32, vp],
    'lnst_avgpool2_bf16x3_bwd': [vp, vp, vp, i32, i32, i32, i32, vp],
    'lnst_f32_to_bf16x3': [vp, vp, i64, i32, vp],
    'lnst_bf16x3_to_f32': [vp, vp, i64, i32, vp],
    'lnst_f32_to_bf16': [vp, vp, i64, vp],
    # TMA-tiled volume kernels (csrc/tiles_tma.cu)
    'lnst_tma_supported': [],
    'lnst_conv_first_bwd_gray_direct': [vp, i32, vp, vp, i32, i32, i32, vp],
    'lnst_advect3_tma': [vp, vp, i32, i32, i32, i32, vp, vp],
    'lnst_splat_wavg_fwd_gather': [vp, vp, vp, vp, vp, GP, FP, i32, vp, BP, vp],
    'lnst_set_raymarch_slab': [i32],
    'lnst_smooth3_relu_fwd_tma': [vp, vp, i32, i32, i32, i32, BP, vp],
    'lnst_smooth3_relu_bwd_tma': [vp, vp, vp, i32, i32, i32, i32, BP, vp],
    'lnst_raymarch_fwd_tma': [vp, vp, i32, i32, i32, i32, f32, i32, BP, vp, vp, vp, vp],
    'lnst_raymarch_fwd_max_tma': [vp, vp, i32, i32, i32, i32, f32, i32, BP, vp, vp, vp, vp, vp],
    'lnst_conv_first_bwd_gray_dot_tc': [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp],
    'lnst_raymarch_bwd_tma': [vp, vp, i32, i32, i32, i32, f32, BP, vp, vp, vp, vp, vp],
    'lnst_bf16_to_f32': [vp, vp, i64, vp],
}


class LnstError(RuntimeError):
    pass


class Lib:
    """A loaded native library with typed entry points; ``kind`` is 'cuda' or 'emu'."""

    def __init__(self, path, kind='cuda'):
        if not os.path.exists(path):
            raise LnstError('native library not found: %s (run __graft_entry__.build())' % path)
        self.path, self.kind = path, kind
        self.dll = C.CDLL(path)
        self.launches = 0
        for name, args in SIGNATURES.items():
            fn = getattr(self.dll, name)
            fn.argtypes = args
            fn.restype = C.c_int
        self.dll.lnst_version.restype = C.c_char_p
        self.dll.lnst_version.argtypes = []
        self.dll.lnst_workspace_bytes.restype = C.c_int64
        self.dll.lnst_workspace_bytes.argtypes = [C.c_char_p, C.POINTER(C.c_int64), i32]
        self.has_tc = self.has_tma = False
        if kind == 'cuda' and hasattr(self.dll, 'lnst_tc_supported'):
            for name, args in CUDA_ONLY.items():
                fn = getattr(self.dll, name)
                fn.argtypes = args
                fn.restype = C.c_int
            self.has_tc = True
            self.has_tma = bool(self.dll.lnst_tma_supported())
        if self.dll.lnst_abi_version() != 1:
            raise LnstError('ABI version mismatch in %s' % path)

    def version(self):
        return self.dll.lnst_version().decode()

    def workspace_bytes(self, op, dims=()):
        arr = (C.c_int64 * max(len(dims), 1))(*[int(d) for d in dims])
        return int(self.dll.lnst_workspace_bytes(op.encode(), arr, len(dims)))

    def call(self, name, *args):
        rc = getattr(self.dll, name)(*args)
        self.launches += 1
        if rc != 0:
            what = 'bad argument' if rc < 0 else 'cudaError %d' % rc
            raise LnstError('%s failed: %s' % (name, what))


_lib = None


def get():
    """The CUDA library (loaded on first use).  Raises if it or a CUDA device is missing."""
    global _lib
    if _lib is None:
        if not torch.cuda.is_available():
            raise LnstError('lnst needs a CUDA device (sm_100a); there is no CPU fallback')
        _lib = Lib(LIB_PATH, 'cuda')
    return _lib


def set_for_testing(lib):
    """Install an explicit library handle (used by tests with the CPU interpreter)."""
    global _lib
    _lib = lib


def stream_ptr(device):
    if device.type != 'cuda':
        return None
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Device pointer of a contiguous fp32/bf16 tensor (None -> NULL)."""
    if t is None:
        return None
    lib = get()
    if lib.kind == 'cuda' and not t.is_cuda:
        raise LnstError('CPU tensor passed to the CUDA library')
    if lib.kind == 'emu' and t.is_cuda:
        raise LnstError('CUDA tensor passed to the CPU interpreter')
    if not t.is_contiguous():
        raise LnstError('non-contiguous tensor')
    return C.c_void_p(t.data_ptr())


def make_box(lo, hi):
    """LnstBox from inclusive (z,y,x) voxel bounds."""
    b = LnstBox()
    for a in range(3):
        b.lo[a] = int(lo[a])
        b.hi[a] = int(hi[a])
    return b


def make_grid(dim, res, domain, nsize, clip):
    """LnstGrid from the reference's (domain, res) lists ((z,)y,x order, transform.py:1316-1330)."""
    import numpy as np
    g = LnstGrid()
    g.dim = dim
    res = [int(r) for r in res]
    dom = [float(d) for d in domain]
    if dim == 2:
        res3, dom3 = [1] + res, [1.0] + dom
    else:
        res3, dom3 = res, dom
    for a in range(3):
        g.res[a] = res3[a]
        g.domain[a] = dom3[a]
    g.cell = float(np.float32(dom[0]) / np.float32(res[0]))
    g.nsize = int(nsize)
    g.clip = 1 if clip else 0
    return g


class _Range:
    """NVTX range around one stage of the step (tracing).  Off unless LNST_NVTX=1 or
    ``enable_nvtx(True)``: under CUDA-graph replay the ranges would only mark the capture."""
    enabled = os.environ.get('LNST_NVTX', '') not in ('', '0')

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _Range.enabled and torch.cuda.is_available():
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if _Range.enabled and torch.cuda.is_available():
            torch.cuda.nvtx.range_pop()
        return False


def nvtx(name):
    return _Range(name)


def enable_nvtx(on=True):
    _Range.enabled = bool(on)
