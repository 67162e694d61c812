"""Point -> voxel pooling (stage 1b).

API mirrors seg3d/ops/voxel_pooling/voxel_pooling.py:10-79 (``voxel_avg_pooling(feats, coords, counts)``,
``voxel_max_pooling(feats, coords)``) and the torch_scatter call inside VFE.forward
(seg3d/models/voxel_encoders/vfe.py:24-25).  fp32 atomics in L2; the max is bit-exact and order independent.
"""
import ctypes
import os

import torch
from torch.autograd import Function

from .. import _lib


_MAX_IMPL = os.environ.get('OS3D_SCATTER_MAX', 'sorted')      # 'sorted' | 'atomic' (bf16 inference path)


def _prep(feats, ids):
    _lib.require_cuda(feats, ids)
    if feats.dim() != 2:
        raise RuntimeError('feats must be [N, C]')
    return feats.float().contiguous(), ids.long().contiguous()


class _ScatterMax(Function):
    @staticmethod
    def forward(ctx, feats, ids, m, fix_empty):
        f, i = _prep(feats, ids)
        out = torch.empty((m, f.shape[1]), dtype=torch.float32, device=f.device)
        _lib.call('os3d_scatter_max_f32', f, i, f.shape[0], f.shape[1], out, m, int(fix_empty),
                  work=lambda: (f.shape[0] + m) * f.shape[1] * 4 + f.shape[0] * 8)
        ctx.save_for_backward(f, i, out)
        ctx.in_dtype = feats.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        f, i, out = ctx.saved_tensors
        gi = torch.empty_like(f)
        arg = torch.empty(out.shape, dtype=torch.int32, device=f.device)          # argmax row per (voxel, channel)
        _lib.call('os3d_scatter_max_bwd_f32', g.float().contiguous(), f, out, i, f.shape[0], f.shape[1], out.shape[0], arg, gi)
        return gi.to(ctx.in_dtype), None, None, None


class _ScatterMean(Function):
    @staticmethod
    def forward(ctx, feats, ids, m):
        f, i = _prep(feats, ids)
        out = torch.empty((m, f.shape[1]), dtype=torch.float32, device=f.device)
        counts = torch.empty(m, dtype=torch.int32, device=f.device)
        _lib.call('os3d_scatter_mean_f32', f, i, f.shape[0], f.shape[1], out, counts, None, m,
                  work=lambda: (f.shape[0] + m) * f.shape[1] * 4 + f.shape[0] * 8)
        ctx.save_for_backward(i, counts)
        ctx.n, ctx.in_dtype = f.shape[0], feats.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        i, counts = ctx.saved_tensors
        gi = torch.empty((ctx.n, g.shape[1]), dtype=torch.float32, device=g.device)
        _lib.call('os3d_scatter_mean_bwd_f32', g.float().contiguous(), i, counts, ctx.n, g.shape[1], counts.shape[0], gi)
        return gi.to(ctx.in_dtype), None, None


def _num_rows(ids, m):
    if m is None:                      # torch_scatter sizes the output by index.max() + 1 (one host sync)
        m = int(ids.max().item()) + 1 if ids.numel() else 0
    return m


def scatter_max(feats, ids, m=None, fix_empty=True):
    """scatter(feats[ids != -1], ids[ids != -1], dim=0, reduce='max'); rows nothing maps to are 0.  fp32 out, except for
    bf16 features outside autograd (bf16 inference): those are reduced as they are and come back as bf16 (exact -- every
    maximum is one of the inputs), without an fp32 copy of the point features (os3d_scatter_max_bf16)."""
    m = _num_rows(ids, m)
    if (feats.dtype == torch.bfloat16 and feats.dim() == 2 and feats.shape[1] % 8 == 0
            and ids.dtype in (torch.int32, torch.int64) and not (torch.is_grad_enabled() and feats.requires_grad)):
        _lib.require_cuda(feats, ids)
        feats, ids = feats.contiguous(), ids.contiguous()
        n, c = feats.shape
        out = torch.empty((m, c), dtype=torch.bfloat16, device=feats.device)
        if _MAX_IMPL == 'sorted' and c <= 2048:
            # sort the point rows by voxel id, reduce each run once: no atomics (a voxel holds ~1.5 points)
            tb = ctypes.c_int64(0)
            _lib.lib().os3d_scatter_max_sorted_scratch(n, m, ctypes.byref(tb))
            keys = torch.empty((2, max(n, 1)), dtype=torch.int32, device=feats.device)
            rows = torch.empty((2, max(n, 1)), dtype=torch.int32, device=feats.device)
            temp = torch.empty(max(tb.value, 1), dtype=torch.uint8, device=feats.device)
            _lib.call('os3d_scatter_max_sorted_bf16', feats, ids, int(ids.dtype == torch.int64), n, c, keys[0], keys[1], rows[0],
                      rows[1], temp, tb.value, out, m, int(fix_empty),
                      work=lambda: n * c * 2 + n * ids.element_size() + m * c * 2)
            return out
        acc = torch.empty((m, c), dtype=torch.float32, device=feats.device)
        _lib.call('os3d_scatter_max_bf16', feats, ids, int(ids.dtype == torch.int64), n, c, acc, out, m, int(fix_empty),
                  work=lambda: n * c * 2 + n * ids.element_size() + m * c * 2)
        return out
    return _ScatterMax.apply(feats, ids, m, fix_empty)


def scatter_mean(feats, ids, m=None):
    """scatter(..., reduce='mean'); fp32 out.  bf16 rows pooled into <= 64 rows outside autograd (the SE layer's per-frame
    mean in bf16 inference) are read as they are (os3d_scatter_mean_small_bf16) -- no fp32 copy, no int64 index copy."""
    m = _num_rows(ids, m)
    if (feats.dtype == torch.bfloat16 and feats.dim() == 2 and feats.shape[1] % 8 == 0 and feats.shape[1] <= 2048 and m <= 64
            and ids.dtype in (torch.int32, torch.int64) and not (torch.is_grad_enabled() and feats.requires_grad)):
        _lib.require_cuda(feats, ids)
        feats, ids = feats.contiguous(), ids.contiguous()
        n, c = feats.shape
        out = torch.empty((m, c), dtype=torch.float32, device=feats.device)
        counts = torch.empty(max(m, 1), dtype=torch.int32, device=feats.device)
        _lib.call('os3d_scatter_mean_small_bf16', feats, ids, int(ids.dtype == torch.int64), n, c, out, counts, m,
                  work=lambda: n * c * 2 + n * ids.element_size() + m * c * 4)
        return out
    return _ScatterMean.apply(feats, ids, m)


class VoxelAvgPoolingFunction(Function):
    """voxel_pooling.py:10-60: out[pos] += feats[i] / counts[pos], ids outside [0, M) skipped."""

    @staticmethod
    def forward(ctx, feats, coords, counts):
        f, i = _prep(feats, coords)
        counts = counts.int().contiguous()
        m = counts.shape[0]
        out = torch.empty((m, f.shape[1]), dtype=torch.float32, device=f.device)
        _lib.call('os3d_scatter_mean_f32', f, i, f.shape[0], f.shape[1], out, None, counts, m)
        ctx.save_for_backward(i, counts)
        ctx.n, ctx.in_dtype = f.shape[0], feats.dtype
        return out.to(feats.dtype)

    @staticmethod
    def backward(ctx, g):
        i, counts = ctx.saved_tensors
        gi = torch.empty((ctx.n, g.shape[1]), dtype=torch.float32, device=g.device)
        _lib.call('os3d_scatter_mean_bwd_f32', g.float().contiguous(), i, counts, ctx.n, g.shape[1], counts.shape[0], gi)
        return gi.to(ctx.in_dtype), None, None


class VoxelMaxPooling(object):
    def __call__(self, feats, coords):
        return scatter_max(feats, coords)

    def __repr__(self):
        return f'{self.__class__.__name__}()'


voxel_avg_pooling = VoxelAvgPoolingFunction.apply
voxel_max_pooling = VoxelMaxPooling()
