"""Stage 1a on the GPU: dynamic voxelization behind the reference's VoxelGenerator interface.

Mirrors seg3d/core/voxel/voxel_generator.py:5-52 (same constructor, properties and ``generate`` contract) and
adds the batched device entry point ``voxelize_batch`` that produces exactly what ``collate_batch`` +
``load_data_to_gpu`` hand the model today (seg3d/datasets/waymo_dataset.py:339-376, seg3d/utils/data_utils.py:6-15),
so voxelization becomes the first kernel of the forward instead of a 0.25 s numba loop per frame in a DataLoader
worker.
"""
import numpy as np
import torch

from .. import _lib


def _geometry(voxel_size, point_cloud_range):
    pcr = np.array(point_cloud_range, dtype=np.float32)
    vs = np.array(voxel_size, dtype=np.float32)
    grid = np.round((pcr[3:] - pcr[:3]) / vs).astype(np.int64)      # voxel_generator.py:17-18
    return vs, pcr, grid


def voxelize_batch(points, voxel_size, point_cloud_range, has_batch=True):
    """points: CUDA float32 [N, (1+)D], frames contiguous and ascending in column 0.
    Returns (voxel_coords int32 [M, 4] (b, z, y, x) in first-occurrence order, point_voxel_ids int64 [N], -1 = dropped).
    One device->host read (the voxel count) sizes the result."""
    _lib.require_cuda(points)
    if points.dtype != torch.float32:
        raise RuntimeError('voxelize_batch takes float32 points')
    points = points.contiguous()
    vs, pcr, grid = _geometry(voxel_size, point_cloud_range)
    n, stride = points.shape
    dev = points.device
    import ctypes
    cap, nb = ctypes.c_int64(0), ctypes.c_int64(0)
    _lib.lib().os3d_voxelize_scratch(n, ctypes.byref(cap), ctypes.byref(nb))
    table = torch.empty(cap.value * 2, dtype=torch.int64, device=dev)          # 16-byte slots
    slot_of = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    block_sums = torch.empty(nb.value + 1, dtype=torch.int32, device=dev)
    coors = torch.empty((max(n, 1), 4), dtype=torch.int32, device=dev)
    pvid = torch.empty(n, dtype=torch.int64, device=dev)
    num = torch.empty(1, dtype=torch.int32, device=dev)                # written by the entry point
    _lib.call('os3d_voxelize', points, n, stride, int(has_batch), float(pcr[0]), float(pcr[1]), float(pcr[2]),
              float(vs[0]), float(vs[1]), float(vs[2]), int(grid[0]), int(grid[1]), int(grid[2]), table, cap.value,
              slot_of, block_sums, nb.value, coors, pvid, num,
              work=lambda: n * (stride * 4 + 8 + 16))        # points in, ids out, <= one coordinate row per point
    m = int(num.item())
    return coors[:m], pvid


def cart2polar_rows(points, has_batch=True):
    """Device version of cart2polar + concatenate (pointops_utils.py:8-11, waymo_dataset.py:270-273):
    (b,) x, y, z, rest -> (b,) rho, phi, z, x, y, rest.  atan2f differs from numpy in the last ulp, so the
    bit-exact parity tests feed host-computed polar rows instead."""
    _lib.require_cuda(points)
    points = points.contiguous()
    out = torch.empty((points.shape[0], points.shape[1] + 2), dtype=torch.float32, device=points.device)
    _lib.call('os3d_cart2polar_rows', points, points.shape[0], points.shape[1], int(has_batch), out)
    return out


class VoxelGenerator(object):
    """Same interface as the reference class (voxel_generator.py:5-52); the work runs on cuda:current."""

    def __init__(self, voxel_size, point_cloud_range):
        self._voxel_size, self._point_cloud_range, self._grid_size = _geometry(voxel_size, point_cloud_range)

    def generate(self, points):
        """points [N, >=3] (numpy or CUDA tensor, xyz first) -> (coors int32 [M, 3] zyx, point_voxel_ids int32 [N]).
        numpy in -> numpy out, like the reference; tensor in -> tensors out."""
        as_numpy = isinstance(points, np.ndarray)
        pts = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32)).cuda() if as_numpy else points
        coors, pvid = voxelize_batch(pts, self._voxel_size, self._point_cloud_range, has_batch=False)
        coors, pvid = coors[:, 1:], pvid.int()
        if as_numpy:
            return coors.cpu().numpy(), pvid.cpu().numpy()
        return coors, pvid

    @property
    def voxel_size(self):
        return self._voxel_size

    @property
    def point_cloud_range(self):
        return self._point_cloud_range

    @property
    def grid_size(self):
        return self._grid_size

    def __repr__(self):
        return (f'{self.__class__.__name__}(voxel_size={self._voxel_size}, '
                f'point_cloud_range={self._point_cloud_range.tolist()}, grid_size={self._grid_size.tolist()})')
