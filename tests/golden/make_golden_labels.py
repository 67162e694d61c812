"""Golden vectors of the loss-side label preparation, produced by the REFERENCE's own code (build container only):
    python tests/golden/make_golden_labels.py
  * WaymoDataset.prepare_voxel_labels (seg3d/datasets/waymo_dataset.py:213-246), called unbound with ignore_index = 255,
    on a small synthetic frame voxelized by the reference's numba voxelizer  -> voxel_labels.npz
  * get_voxel_centers (seg3d/utils/pointops_utils.py:14-22) for the level-1 and the stride-8 voxels -> same file
(the reference's knn_query is a GPU-only extension; the 1-NN over those centres is the oracle's restatement.)"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, '..', '..')
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

import openseg3d_b200.compat as compat  # noqa: E402

compat.install(layers=False)            # stubs for the CUDA-only / absent modules the dataset module imports
from seg3d.datasets.waymo_dataset import WaymoDataset  # noqa: E402
from seg3d.core import VoxelGenerator  # noqa: E402
from seg3d.utils.pointops_utils import get_voxel_centers  # noqa: E402

from openseg3d_b200 import synthetic  # noqa: E402


def main():
    rng = np.random.default_rng(5)
    vs, pcr = [0.1, 0.1, 0.1], [-72, -72, -2, 72, 72, 4.4]
    pts, _ = synthetic.make_frame(3, 1, 16, 300)
    # thicken the frame so that many voxels hold several points (ties and majorities both occur)
    pts = np.concatenate([pts, pts + rng.normal(0, 0.03, pts.shape).astype(np.float32), pts[:500] * np.float32(50.0)])
    gen = VoxelGenerator(voxel_size=vs, point_cloud_range=pcr)
    coors, pvid = gen.generate(pts)
    labels = rng.integers(0, 22, pts.shape[0]).astype(np.uint8)
    labels[rng.random(pts.shape[0]) < 0.15] = 255
    data = {'point_voxel_ids': pvid, 'point_labels': labels, 'voxel_coords': coors}
    WaymoDataset.prepare_voxel_labels(SimpleNamespace(ignore_index=255), data)
    # stride-8 voxels of the same frame (what the backbone's aux branch sees) and both sets of centres
    aux = np.unique(coors // 8, axis=0)
    c1 = get_voxel_centers(torch.from_numpy(coors.astype(np.int64)), 1.0, vs, pcr).numpy()
    c8 = get_voxel_centers(torch.from_numpy(aux.astype(np.int64)), 8.0, vs, pcr).numpy()
    np.savez_compressed(os.path.join(HERE, 'voxel_labels.npz'), point_voxel_ids=pvid.astype(np.int64), point_labels=labels,
                        voxel_coords=coors.astype(np.int32), voxel_labels=data['voxel_labels'], aux_coords=aux.astype(np.int32),
                        centers=c1, aux_centers=c8)
    vl = data['voxel_labels']
    print('points', pts.shape[0], 'voxels', coors.shape[0], 'outside', int((pvid < 0).sum()), 'ignored voxels', int((vl == 255).sum()),
          'aux voxels', aux.shape[0])


if __name__ == '__main__':
    main()
