"""Point -> voxel pooling (stage 1b).

API mirrors seg3d/ops/voxel_pooling/voxel_pooling.py:10-79 (``voxel_avg_pooling(feats, coords, counts)``,
``voxel_max_pooling(feats, coords)``) and the torch_scatter call inside VFE.forward
(seg3d/models/voxel_encoders/vfe.py:24-25).  fp32 atomics in L2; the max is bit-exact and order independent.
"""
import torch
from torch.autograd import Function

from .. import _lib


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
    """scatter(feats[ids != -1], ids[ids != -1], dim=0, reduce='max'); rows nothing maps to are 0."""
    return _ScatterMax.apply(feats, ids, _num_rows(ids, m), fix_empty)


def scatter_mean(feats, ids, m=None):
    """scatter(..., reduce='mean')."""
    return _ScatterMean.apply(feats, ids, _num_rows(ids, m))


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
