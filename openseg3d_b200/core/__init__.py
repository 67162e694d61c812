from .voxel import VoxelGenerator, voxelize_batch, cart2polar_rows

__all__ = ['VoxelGenerator', 'voxelize_batch', 'cart2polar_rows']
