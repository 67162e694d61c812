"""Voxel -> point gather (stage 1c); same callable object as seg3d/ops/voxel_to_point/voxel_to_point.py:5-20."""
import torch
from torch.autograd import Function

from .. import _lib


class _Gather(Function):
    @staticmethod
    def forward(ctx, feats, coords):
        _lib.require_cuda(feats, coords)
        if feats.dtype not in (torch.float32, torch.bfloat16):
            raise RuntimeError('voxel_to_point takes float32 or bfloat16 features')
        feats = feats.contiguous()
        ids = coords.long().contiguous()
        out = torch.empty((ids.shape[0], feats.shape[-1]), dtype=feats.dtype, device=feats.device)
        _lib.call('os3d_gather_rows', feats, ids, ids.shape[0], feats.shape[-1], feats.element_size(), out,
                  work=lambda: (ids.shape[0] + feats.shape[0]) * feats.shape[-1] * feats.element_size() + ids.shape[0] * 8)
        ctx.save_for_backward(ids)
        ctx.m = feats.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        ids, = ctx.saved_tensors
        gf = torch.empty((ctx.m, g.shape[1]), dtype=torch.float32, device=g.device)
        _lib.call('os3d_scatter_add_rows_f32', g.float().contiguous(), ids, ids.shape[0], g.shape[1], gf, ctx.m)
        return gf.to(g.dtype), None


class VoxelToPoint(object):
    def __call__(self, feats, coords):
        """feats (num_voxels, C), coords (N) -> (N, C); rows whose id is -1 are zero."""
        return _Gather.apply(feats, coords)

    def __repr__(self):
        return f'{self.__class__.__name__}()'


voxel_to_point = VoxelToPoint()
