from .voxel_encoders import VFE
from .layers import (SparseWindowPartitionLayer, WindowAttention, CosineMultiheadAttention, SWFormerBlock, EncoderLayer,
                     MLP, DropPath, FlattenSELayer)
from .backbones import PointTransformer, SparseBasicBlock, UpBlock, ConvModule
from .segmentors import Segformer, build_segformer

__all__ = ['VFE', 'SparseWindowPartitionLayer', 'WindowAttention', 'CosineMultiheadAttention', 'SWFormerBlock',
           'EncoderLayer', 'MLP', 'DropPath', 'FlattenSELayer', 'PointTransformer', 'SparseBasicBlock', 'UpBlock',
           'ConvModule', 'Segformer', 'build_segformer']
