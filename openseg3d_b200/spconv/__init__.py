"""A spconv.pytorch-shaped namespace over libos3d (stages 2-3).

Same names and keyword arguments the reference uses (seg3d/utils/spconv_utils.py:13-32,
seg3d/models/backbones/pointtransformer.py:13,26-32,69,132-136,184-189): SparseConvTensor, SubMConv3d, SparseConv3d,
SparseInverseConv3d, SparseSequential, SparseModule.  Only what that model exercises is implemented: 3x3x3 kernels,
SubM stride 1 / pad 1, strided conv stride 2 / pad 1, dilation 1.
"""
from .tensor import SparseConvTensor
from .modules import SparseModule, SparseSequential, SubMConv3d, SparseConv3d, SparseInverseConv3d
from .rulebook import SubmRulebook, StridedRulebook, build_subm_rulebook, build_strided_rulebook

__all__ = ['SparseConvTensor', 'SparseModule', 'SparseSequential', 'SubMConv3d', 'SparseConv3d', 'SparseInverseConv3d',
           'SubmRulebook', 'StridedRulebook', 'build_subm_rulebook', 'build_strided_rulebook']
