"""knn_query (seg3d/ops/knn_query/knn_query.py:8-29): same call and return values as the reference op."""
import torch

from .. import _lib


def knn_query(nsample, xyz, new_xyz, offset, new_offset):
    """xyz [n, 3], new_xyz [m, 3] (None = xyz), offset / new_offset [b] int32 cumulative segment ends
    -> (idx int32 [m, nsample], dist float32 [m, nsample], the Euclidean distances)."""
    if new_xyz is None:
        new_xyz = xyz
    _lib.require_cuda(xyz, new_xyz, offset, new_offset)
    if not (xyz.is_contiguous() and new_xyz.is_contiguous()):
        raise RuntimeError('knn_query takes contiguous tensors')             # the reference asserts (knn_query.py:17)
    if xyz.dtype != torch.float32 or new_xyz.dtype != torch.float32:
        raise RuntimeError('knn_query takes float32 coordinates')
    m = new_xyz.shape[0]
    idx = torch.zeros((m, nsample), dtype=torch.int32, device=xyz.device)
    dist2 = torch.zeros((m, nsample), dtype=torch.float32, device=xyz.device)
    _lib.call('os3d_knn_query', xyz, new_xyz, m, int(nsample), offset.int().contiguous(), new_offset.int().contiguous(),
              int(offset.numel()), idx, dist2)
    return idx, torch.sqrt(dist2)
