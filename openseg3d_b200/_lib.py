"""ctypes binding of libos3d.so (the C ABI declared in include/os3d.h).

There is no CPU fallback and no alternative backend: if the shared library is missing or a call fails the
error is raised immediately (the north star's "fail loudly").  Tensors are passed as raw device pointers,
the stream is torch's current CUDA stream, scratch comes from torch's caching allocator.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libos3d.so')

I32, I64, F32, PTR = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p


class WindowCfg(ctypes.Structure):
    """os3d_window_cfg_t"""
    _fields_ = [(n, ctypes.c_int) for n in
                ['sparse_x', 'sparse_y', 'sparse_z', 'win_x', 'win_y', 'win_z', 'nwin_x', 'nwin_y', 'nwin_z',
                 'shift_x', 'shift_y', 'shift_z', 'n_levels']] + \
               [('lvl_lo', ctypes.c_int * 4), ('lvl_hi', ctypes.c_int * 4), ('lvl_tokens', ctypes.c_int * 4)]


class MlpLayer(ctypes.Structure):
    """os3d_mlp_layer (include/os3d.h)."""
    _fields_ = [('w', ctypes.c_void_p), ('bias', ctypes.c_void_p), ('k', ctypes.c_int), ('n', ctypes.c_int),
                ('act', ctypes.c_int)]


# name -> argtypes (the trailing stream argument included); every function returns int
SIGNATURES = {
    'os3d_voxelize_scratch': [I64, ctypes.POINTER(I64), ctypes.POINTER(I64)],
    'os3d_spconv_bf16_packed_elems': [I32, I32, ctypes.POINTER(I64)],
    'os3d_voxelize': [PTR, I64, I32, I32, F32, F32, F32, F32, F32, F32, I32, I32, I32, PTR, I64, PTR, PTR, I64, PTR, PTR,
                      PTR, PTR],
    'os3d_cart2polar_rows': [PTR, I64, I32, I32, PTR, PTR],
    'os3d_scatter_max_f32': [PTR, PTR, I64, I32, PTR, I64, I32, PTR],
    'os3d_scatter_max_bf16': [PTR, PTR, I32, I64, I32, PTR, PTR, I64, I32, PTR],
    'os3d_scatter_mean_small_bf16': [PTR, PTR, I32, I64, I32, PTR, PTR, I64, PTR],
    'os3d_scatter_max_sorted_scratch': [I64, I64, ctypes.POINTER(ctypes.c_int64)],
    'os3d_scatter_max_sorted_bf16': [PTR, PTR, I32, I64, I32, PTR, PTR, PTR, PTR, PTR, I64, PTR, I64, I32, PTR],
    'os3d_scatter_mean_f32': [PTR, PTR, I64, I32, PTR, PTR, PTR, I64, PTR],
    'os3d_scatter_max_bwd_f32': [PTR, PTR, PTR, PTR, I64, I32, I64, PTR, PTR, PTR],
    'os3d_scatter_mean_bwd_f32': [PTR, PTR, PTR, I64, I32, I64, PTR, PTR],
    'os3d_gather_rows': [PTR, PTR, I64, I32, I32, PTR, PTR],
    'os3d_scatter_add_rows_f32': [PTR, PTR, I64, I32, PTR, I64, PTR],
    'os3d_hash_build': [PTR, I64, I32, I32, I32, PTR, I64, PTR],
    'os3d_subm_table': [PTR, I64, I32, I32, I32, PTR, I64, PTR, PTR, PTR],
    'os3d_strided_sites': [PTR, I64, I32, I32, I32, I32, PTR, I64, PTR, PTR, I64, PTR, I64, PTR, PTR],
    'os3d_strided_tables': [PTR, I64, I32, I32, I32, PTR, I64, PTR, I64, I32, I32, I32, PTR, PTR, PTR, PTR, PTR, PTR],
    'os3d_spconv_fwd_f32': [PTR, PTR, I64, I32, I32, PTR, PTR, PTR, PTR],
    'os3d_kernel_map_order_scratch': [I64, ctypes.POINTER(I64)],
    'os3d_kernel_map_order': [PTR, I64, PTR, PTR, PTR, PTR, PTR, I64, PTR],
    'os3d_kernel_map_tiles': [PTR, I64, PTR, PTR, PTR, PTR],
    'os3d_spconv_fwd_bf16': [PTR, I64, PTR, PTR, PTR, I64, I32, I32, PTR, PTR, PTR, PTR, I32, PTR, PTR],
    'os3d_spconv_fwd_bf16_ld': [PTR, I64, PTR, PTR, PTR, I64, I32, I32, PTR, PTR, PTR, PTR, I32, PTR, I64, PTR],
    'os3d_pack_weight_f32': [PTR, I32, I32, PTR, PTR],
    'os3d_linear_bf16_packed_elems': [I32, I32, ctypes.POINTER(I64)],
    'os3d_pack_linear_bf16': [PTR, I32, I32, PTR, PTR],
    'os3d_linear_bf16': [PTR, I64, I32, I32, PTR, PTR, I32, PTR, PTR, PTR, F32, PTR, PTR, I32, PTR, I64, PTR],
    'os3d_linear_tc_bf16': [PTR, I64, I32, I32, PTR, PTR, I32, PTR, PTR, PTR, F32, PTR, PTR, I32, PTR, I64, PTR],
    'os3d_linear_tc_fits': [I32, I32],
    'os3d_swformer_mlp_fits': [I32, I32],
    'os3d_swformer_mlp_bf16': [PTR, I64, I32, I32, PTR, PTR, PTR, PTR, PTR, PTR, F32, PTR, PTR],
    'os3d_mlp_chain_fits': [ctypes.POINTER(MlpLayer), I32, I32],
    'os3d_mlp_chain_bf16': [PTR, I64, I64, PTR, I64, I32, PTR, PTR, I32, ctypes.POINTER(MlpLayer), I32, PTR, I64, PTR, PTR, F32,
                            PTR, I64, I32, I32, PTR],
    'os3d_pack_weight_bf16': [PTR, I32, I32, PTR, PTR],
    'os3d_wide_linear_plan': [I32, I32, I32, I32, ctypes.POINTER(ctypes.c_int)],
    'os3d_wide_linear_bf16': [PTR, I64, I32, I32, I32, PTR, PTR, PTR, PTR, I64, I32, I32, I32, PTR, I64, PTR],
    'os3d_window_partition': [PTR, I64, I32, ctypes.POINTER(WindowCfg), PTR, PTR, PTR, I64, PTR, PTR, PTR, PTR, PTR, PTR,
                              PTR, PTR, PTR, PTR, PTR, PTR],
    'os3d_group_partition': [PTR, I64, I64, ctypes.POINTER(WindowCfg), PTR, PTR, PTR, I64, PTR, PTR, PTR, PTR, PTR, PTR, PTR,
                             PTR, PTR],
    'os3d_window_attention_bf16_tc': [PTR, PTR, PTR, I64, I64, I64, I32, I32, PTR, PTR, PTR, PTR, F32, PTR, I64, PTR],
    'os3d_window_attention_bf16_tc_prenorm': [PTR, PTR, PTR, I64, I64, I64, I32, I32, PTR, PTR, PTR, PTR, F32, PTR, I64, PTR],
    'os3d_window_attention_bf16_tc_drop': [PTR, PTR, PTR, I64, I64, I64, I32, I32, PTR, PTR, PTR, PTR, F32, F32, ctypes.c_uint64, PTR, I64, PTR],
    'os3d_window_attention_bf16_v2': [PTR, PTR, PTR, I64, I64, I64, I32, I32, PTR, PTR, PTR, PTR, F32, PTR, I64, PTR],
    'os3d_pos_embed': [PTR, I64, I32, I32, I32, I32, F32, I32, PTR, PTR],
    'os3d_voxel_majority_labels': [PTR, PTR, I64, I64, I32, PTR, PTR, PTR, PTR],
    'os3d_argmax_rows': [PTR, I64, I32, I32, PTR, PTR],
    'os3d_knn_query': [PTR, PTR, I64, I32, PTR, PTR, I32, PTR, PTR, PTR],
    'os3d_add_table_rows': [PTR, PTR, PTR, I64, I32, I32, PTR, PTR],
    'os3d_gelu_bf16': [PTR, I64, PTR, PTR],
    'os3d_scale_rows_by_table': [PTR, PTR, PTR, I64, I32, F32, I32, PTR, PTR],
    'os3d_layernorm_residual': [PTR, PTR, PTR, PTR, I64, I32, F32, I32, PTR, PTR],
    'os3d_qk_normalize': [PTR, PTR, I64, I64, I32, I32, I32, PTR],
    'os3d_window_attention': [PTR, PTR, PTR, I64, I64, I64, I32, I32, PTR, PTR, PTR, PTR, ctypes.POINTER(ctypes.c_int * 4), PTR,
                              F32, F32, ctypes.c_uint64, I32, PTR, PTR],
    'os3d_window_attention_bwd': [PTR, PTR, PTR, PTR, PTR, I64, I32, I32, PTR, PTR, PTR, PTR,
                                  ctypes.POINTER(ctypes.c_int * 4), PTR, F32, F32, ctypes.c_uint64, I32, PTR, PTR, PTR, PTR,
                                  PTR, PTR],
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f'{LIB_PATH} is missing: build it with `python openseg3d_b200/csrc/build.py` '
                              '(nvcc, sm_100a). openseg3d_b200 has no CPU or PyTorch fallback.')
        L = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here = header / library mismatch: fail loudly
            fn.argtypes = argtypes
            fn.restype = ctypes.c_int
        L.os3d_error_string.restype = ctypes.c_char_p
        L.os3d_error_string.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def exported_symbols():
    return sorted(SIGNATURES) + ['os3d_error_string', 'os3d_version']


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor must be contiguous."""
    if t is None:
        return None
    assert t.is_contiguous(), 'os3d kernels take contiguous tensors'
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


# CUDA kernels launched by each entry point (memsets not counted) -- bench.py's gpu_launches evidence
KERNELS_PER_CALL = {
    'os3d_voxelize': 5, 'os3d_cart2polar_rows': 1, 'os3d_scatter_max_f32': 2, 'os3d_scatter_max_bf16': 3, 'os3d_scatter_mean_small_bf16': 2, 'os3d_scatter_max_sorted_bf16': 6, 'os3d_scatter_mean_f32': 2,
    'os3d_scatter_max_bwd_f32': 2, 'os3d_scatter_mean_bwd_f32': 1, 'os3d_gather_rows': 1, 'os3d_scatter_add_rows_f32': 1,
    'os3d_hash_build': 1, 'os3d_subm_table': 1, 'os3d_strided_sites': 4, 'os3d_strided_tables': 2,
    'os3d_spconv_fwd_f32': 1, 'os3d_spconv_fwd_bf16': 1, 'os3d_spconv_fwd_bf16_ld': 1, 'os3d_pack_weight_f32': 1, 'os3d_pack_weight_bf16': 1, 'os3d_kernel_map_tiles': 1, 'os3d_kernel_map_order': 2, 'os3d_linear_bf16': 1, 'os3d_linear_tc_bf16': 1, 'os3d_pack_linear_bf16': 1, 'os3d_mlp_chain_bf16': 1, 'os3d_swformer_mlp_bf16': 1,
    'os3d_window_partition': 7, 'os3d_group_partition': 7, 'os3d_pos_embed': 1, 'os3d_qk_normalize': 1,
    'os3d_window_attention': 1, 'os3d_window_attention_bwd': 2, 'os3d_window_attention_bf16_tc': 1, 'os3d_window_attention_bf16_v2': 1, 'os3d_window_attention_bf16_tc_prenorm': 1, 'os3d_window_attention_bf16_tc_drop': 1, 'os3d_wide_linear_bf16': 1, 'os3d_voxel_majority_labels': 2, 'os3d_argmax_rows': 1, 'os3d_layernorm_residual': 1,
}
_launches = 0
PROFILE = None      # bench.py sets this to a list to collect (name, start_event, end_event, work) per call


class Work(float):
    """Algorithmic FLOPs of a call that also carries its algorithmic HBM bytes (``.bytes``) for the bench's roofline table."""

    def __new__(cls, flops, nbytes=0.0, detail=None):
        w = super().__new__(cls, flops)
        w.bytes = float(nbytes)
        w.detail = detail            # e.g. dict(cin=, cout=, m_out=, pairs=) of a sparse conv: bench.py's per-layer table
        return w


class _Raw(object):
    """A raw device address as a call() argument (a strided view's base pointer; the pitch travels separately)."""

    def __init__(self, ptr):
        self.ptr = ptr


def launches():
    """Number of libos3d CUDA kernels launched so far (bench.py reports the delta over its timed region)."""
    return _launches


def call(name, *args, work=None):
    """Invoke an entry point on torch's current stream; raise RuntimeError on a non-zero status.
    ``work``: optional callable returning the algorithmic FLOPs / bytes of this call (evaluated only when profiling)."""
    global _launches
    L = lib()
    conv = [ptr(a) if isinstance(a, torch.Tensor) else (a.ptr if hasattr(a, 'ptr') else a) for a in args]
    prof = PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(L, name)(*conv, stream())
    if prof is not None:
        e1.record()
        w = work() if work is not None else 0.0
        prof.append((name, e0, e1, w if isinstance(w, float) else float(w)))
    _launches += KERNELS_PER_CALL.get(name, 1)
    if rc != 0:
        raise RuntimeError(f'{name} failed: {L.os3d_error_string(rc).decode()} (code {rc})')


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('openseg3d_b200 ops take CUDA tensors (there is no CPU path)')
