"""Operator namespace mirroring seg3d/ops/__init__.py:1-6."""
from .pooling import voxel_avg_pooling, voxel_max_pooling, scatter_max, scatter_mean
from .voxel_to_point import voxel_to_point
from .knn_query import knn_query
from .ingroup_inds import get_inner_win_inds
from .labels import voxel_majority_labels, aux_voxel_labels, get_voxel_centers, predict_labels

__all__ = ['predict_labels', 'voxel_avg_pooling', 'voxel_max_pooling', 'voxel_to_point', 'knn_query', 'get_inner_win_inds', 'scatter_max',
           'scatter_mean', 'voxel_majority_labels', 'aux_voxel_labels', 'get_voxel_centers']
