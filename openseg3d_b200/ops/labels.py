"""Loss-side label preparation on the GPU.

``voxel_majority_labels`` replaces WaymoDataset.prepare_voxel_labels (seg3d/datasets/waymo_dataset.py:213-246: a Python
dict of counters per voxel, run per frame in DataLoader workers); ``aux_voxel_labels`` restates the label transfer of
tools/train.py:86-104: every voxel of the coarsest level (stride 8) takes the label of the level-1 voxel whose centre is
nearest to its own centre -- the reference's 1-NN ``knn_query`` with its tie rule (equal distances keep the lower
index), here through os3d_knn_query."""
import torch

from .. import _lib
from .knn_query import knn_query


def voxel_majority_labels(point_voxel_ids, point_labels, num_voxels, ignore_index=255, check=True):
    """point_voxel_ids [N] (int64, -1 = outside the range; for multi-sweep frames pass the current sweep's rows as the
    reference does), point_labels [N] uint8-valued -> voxel_labels [num_voxels] uint8.  ``check``: one host read that
    raises if a label outside [0, 30] / ignore_index was seen (the kernel orders labels by a 32-bin histogram)."""
    _lib.require_cuda(point_voxel_ids, point_labels)
    ids = point_voxel_ids.long().contiguous()
    labels = point_labels.to(torch.uint8).contiguous()
    if ids.shape[0] != labels.shape[0]:
        raise RuntimeError('voxel_majority_labels: one label per point')
    m = int(num_voxels)
    out = torch.empty(m, dtype=torch.uint8, device=ids.device)
    hist = torch.empty(max(m, 1) * 32, dtype=torch.int32, device=ids.device)
    bad = torch.empty(1, dtype=torch.int32, device=ids.device)
    _lib.call('os3d_voxel_majority_labels', ids, labels, ids.shape[0], m, int(ignore_index), hist, bad, out)
    if check and m and int(bad.item()):
        raise RuntimeError('voxel_majority_labels: labels must lie in [0, 30] or equal ignore_index')
    return out


def get_voxel_centers(voxel_coords_zyx, downsample_scale, voxel_size, point_cloud_range):
    """seg3d/utils/pointops_utils.py:14-22 -- (coord + 0.5) * voxel_size * scale + range minimum, xyz order, float32."""
    centers = voxel_coords_zyx[:, [2, 1, 0]].float()
    vs = torch.tensor(voxel_size, device=centers.device).float() * downsample_scale
    lo = torch.tensor(point_cloud_range[0:3], device=centers.device).float()
    return (centers + 0.5) * vs + lo


def aux_voxel_labels(voxel_labels, voxel_coords, aux_voxel_coords, batch_size, voxel_size, point_cloud_range,
                     aux_scale=8.0):
    """Labels of the auxiliary (coarsest-level) voxels, tools/train.py:86-104: label of the nearest level-1 voxel centre
    inside the same frame.  voxel_coords / aux_voxel_coords: [M, 4] / [M4, 4] (b, z, y, x)."""
    centers = get_voxel_centers(voxel_coords[:, 1:], 1.0, voxel_size, point_cloud_range).contiguous()
    aux_centers = get_voxel_centers(aux_voxel_coords[:, 1:], aux_scale, voxel_size, point_cloud_range).contiguous()
    b = torch.arange(batch_size, device=voxel_coords.device)
    offset = (voxel_coords[:, 0].long()[None, :] <= b[:, None]).sum(dim=1).int()          # cumulative frame ends
    aux_offset = (aux_voxel_coords[:, 0].long()[None, :] <= b[:, None]).sum(dim=1).int()
    idx, _ = knn_query(1, centers, aux_centers, offset, aux_offset)
    return voxel_labels[idx.reshape(-1).long()]


def predict_labels(point_out):
    """tools/test.py:58 -- ``torch.argmax(point_out, dim=1)`` as uint8 labels [N] (ties -> the lowest class), one pass over
    the logits (os3d_argmax_rows) instead of torch's generic reduce + an int64 -> uint8 cast."""
    _lib.require_cuda(point_out)
    if point_out.dim() != 2 or point_out.dtype not in (torch.bfloat16, torch.float32):
        raise RuntimeError('predict_labels takes [N, classes] bf16 / fp32 logits')
    x = point_out.contiguous()
    n, c = x.shape
    out = torch.empty(n, dtype=torch.uint8, device=x.device)
    _lib.call('os3d_argmax_rows', x, n, c, x.element_size(), out, work=lambda: n * c * x.element_size() + n)
    return out
