import torch.nn as nn

from ..ops.pooling import scatter_max, scatter_mean


class VFE(nn.Module):
    """Voxel feature encoder, same interface as seg3d/models/voxel_encoders/vfe.py:6-27: a masked
    scatter(mean | max) of point features over point_voxel_ids (-1 = point outside the range)."""

    def __init__(self, voxel_feature_channel, reduce='mean'):
        super().__init__()
        if reduce not in ('mean', 'max'):
            raise ValueError(reduce)
        self._voxel_feature_channel = voxel_feature_channel
        self.reduce = reduce

    @property
    def voxel_feature_channel(self):
        return self._voxel_feature_channel

    def forward(self, features, index, num_voxels=None):
        """features (N, C), index (N) -> (num_voxels, C).  ``num_voxels`` (known to the caller that just voxelized)
        saves the index.max() host read torch_scatter does; every voxel of the voxelizer owns >= 1 point, so the
        empty-row fix-up pass is skipped in that case."""
        if self.reduce == 'max':
            return scatter_max(features, index, num_voxels, fix_empty=num_voxels is None)
        return scatter_mean(features, index, num_voxels)
