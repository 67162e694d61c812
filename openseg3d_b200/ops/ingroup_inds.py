"""get_inner_win_inds (seg3d/ops/ingroup_inds/ingroup_inds.py:7-20): rank of each element inside its group.

The reference kernel returns an arrival-order rank (atomicAdd race, ingroup_inds_cuda.cu:23).  This op returns the
deterministic member of that set: the number of earlier elements with the same group id.  Like the reference it
needs one host read (the largest group id) to size its counters; the window-partition layer does not go through
this op and needs none.
"""
import torch

from .. import _lib
from .._lib import WindowCfg


def group_partition(group_inds, n_groups):
    """Counting sort of elements by group id.  Returns dict(order, seg_start, seg_len, inner, group_rank, level_info)."""
    _lib.require_cuda(group_inds)
    g = group_inds.long().contiguous()
    n, dev = g.shape[0], g.device
    cfg = WindowCfg()
    cfg.n_levels = 1
    cfg.lvl_lo[0], cfg.lvl_hi[0], cfg.lvl_tokens[0] = 1, 2 ** 31 - 1, 2 ** 31 - 1
    nb = (max(n_groups, 1) + 1023) // 1024
    i32 = dict(dtype=torch.int32, device=dev)
    out = dict(level=torch.empty(n, **i32), group_rank=torch.empty(n, **i32), inner=torch.empty(n, **i32),
               order=torch.empty(n, **i32), seg_start=torch.empty(n + 1, **i32), seg_len=torch.empty(n + 1, **i32),
               pos_seg=torch.empty((max(n, 1), 2), **i32),
               level_info=torch.empty(16, **i32))
    count = torch.empty(max(n_groups, 1), **i32)
    meta = torch.empty(max(n_groups, 1) * 3, **i32)
    block_sums = torch.empty((nb + 1) * 5, **i32)
    import ctypes
    _lib.call('os3d_group_partition', g, n, n_groups, ctypes.byref(cfg), count, meta, block_sums, nb, out['level'],
              out['group_rank'], out['inner'], out['order'], out['seg_start'], out['seg_len'], out['pos_seg'],
              out['level_info'])
    return out


class IngroupIndicesFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, group_inds):
        if group_inds.numel() == 0:
            return torch.zeros_like(group_inds)
        n_groups = int(group_inds.max().item()) + 1
        out = group_partition(group_inds, n_groups)['inner'].to(group_inds.dtype)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, g):
        return None


get_inner_win_inds = IngroupIndicesFunction.apply
