"""PointTransformer backbone: a sparse UNet (48/96/192/384) whose encoder stages end in SWFormer blocks.

Stays in Python / PyTorch as the north star says; module tree and state_dict keys follow
seg3d/models/backbones/pointtransformer.py:13-219 and seg3d/utils/spconv_utils.py:13-32 so that reference
checkpoints load.  Everything heavy it calls is libos3d: kernel maps, sparse convolutions (BatchNorm + ReLU + residual
folded into the conv epilogue at inference), window partition and window attention.
"""
from functools import partial

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import spconv
from ..spconv.modules import bn_scale_shift
from .layers import SparseWindowPartitionLayer, SWFormerBlock, linear_in


def replace_feature(out, new_features):
    return out.replace_feature(new_features)


def ConvModule(in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, conv_type='subm', norm_fn=None,
               act_fn=None, indice_key=None):
    """conv(bias=False) + norm + act as a SparseSequential (spconv_utils.py:13-32)."""
    if conv_type == 'subm':
        conv = spconv.SubMConv3d(in_channels, out_channels, kernel_size, padding=padding, dilation=dilation, bias=False,
                                 indice_key=indice_key)
    elif conv_type == 'spconv':
        conv = spconv.SparseConv3d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                   dilation=dilation, bias=False, indice_key=indice_key)
    elif conv_type == 'inverseconv':
        conv = spconv.SparseInverseConv3d(in_channels, out_channels, kernel_size, bias=False, indice_key=indice_key)
    else:
        raise NotImplementedError(conv_type)
    return spconv.SparseSequential(conv, norm_fn(out_channels), act_fn)


class SparseBasicBlock(spconv.SparseModule):
    """Two SubM convs with a residual (pointtransformer.py:13-66).  with_se / with_sa are never enabled by the
    reference configs and are not built."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, with_se=False, with_sa=False, norm_fn=None, act_fn=None,
                 indice_key=None):
        super().__init__()
        assert norm_fn is not None
        if with_se or with_sa:
            raise NotImplementedError('with_se / with_sa are unused by the reference model')
        self.conv1 = spconv.SubMConv3d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=True,
                                       indice_key=indice_key)
        self.bn1 = norm_fn(planes)
        self.act = act_fn
        self.conv2 = spconv.SubMConv3d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=True,
                                       indice_key=indice_key)
        self.bn2 = norm_fn(planes)

    def forward(self, x, out=None):
        """``out`` (inference only): destination view for the block's output features."""
        if not self.training and not torch.is_grad_enabled():   # inference: BN, residual add and ReLU live in the conv epilogues
            s1, b1 = bn_scale_shift(self.bn1, self.conv1.bias)
            s2, b2 = bn_scale_shift(self.bn2, self.conv2.bias)
            mid = self.conv1(x, scale=s1, shift=b1, relu=True)
            return self.conv2(mid, scale=s2, shift=b2, residual=x.features, relu=True, out=out)
        out = self.conv1(x)
        out = replace_feature(out, self.act(self.bn1(out.features)))
        out = self.conv2(out)
        out = replace_feature(out, self.bn2(out.features))
        return replace_feature(out, self.act(out.features + x.features))


class UpBlock(spconv.SparseModule):
    """Decoder block (pointtransformer.py:69-113): transform(lateral) -> cat with bottom -> bottleneck ->
    + channel_reduction(cat) -> out (inverse conv to the finer level, or SubM at level 1)."""

    def __init__(self, inplanes, planes, norm_fn, act_fn, conv_type, layer_id):
        super().__init__()
        self.transform = SparseBasicBlock(inplanes, inplanes, norm_fn=norm_fn, act_fn=act_fn,
                                          indice_key='subm' + str(layer_id))
        self.bottleneck = ConvModule(2 * inplanes, inplanes, 3, padding=1, norm_fn=norm_fn, act_fn=act_fn,
                                     indice_key='subm' + str(layer_id))
        if conv_type == 'inverseconv':
            self.out = ConvModule(inplanes, planes, 3, norm_fn=norm_fn, act_fn=act_fn, conv_type=conv_type,
                                  indice_key='spconv' + str(layer_id))
        elif conv_type == 'subm':
            self.out = ConvModule(inplanes, planes, 3, padding=1, norm_fn=norm_fn, act_fn=act_fn, conv_type=conv_type,
                                  indice_key='subm' + str(layer_id))
        else:
            raise NotImplementedError(conv_type)

    @staticmethod
    def channel_reduction(x, out_channels):
        features = x.features
        n, in_channels = features.shape
        assert (in_channels % out_channels == 0) and (in_channels >= out_channels)
        return replace_feature(x, features.view(n, out_channels, -1).sum(dim=2))

    feeds_upblock = False      # set by PointTransformer: the next decoder block concatenates this block's output

    def forward(self, x_bottom, x_lateral):
        if not self.training and not torch.is_grad_enabled():
            # inference: no torch.cat -- x_bottom already sits in the left half of a double-width buffer (written there by
            # the previous block's `out` conv; copied once at level 4, where it is the encoder output) and the transform
            # block writes its result into the right half; BN + ReLU + channel_reduction(cat) + add run in the
            # bottleneck's epilogue
            c = x_bottom.features.shape[1]
            wide = getattr(x_bottom, '_os3d_wide', None)
            if wide is None or wide.shape[0] != x_lateral.features.shape[0] or wide.dtype != x_lateral.features.dtype:
                wide = torch.empty((x_bottom.features.shape[0], 2 * c), dtype=x_lateral.features.dtype,
                                   device=x_bottom.features.device)
                wide[:, :c] = x_bottom.features
            x_trans = self.transform(x_lateral, out=wide[:, c:])
            x = replace_feature(x_trans, wide)
            conv, bn = self.bottleneck[0], self.bottleneck[1]
            scale, shift = bn_scale_shift(bn, conv.bias)
            return self.out(conv(x, scale=scale, shift=shift, residual=x.features, relu=3), wide_out=self.feeds_upblock)
        x_trans = self.transform(x_lateral)
        x = replace_feature(x_trans, torch.cat([x_bottom.features, x_trans.features], dim=1))
        x_m = self.bottleneck(x)
        x = self.channel_reduction(x, x_m.features.shape[1])
        x = replace_feature(x, x_m.features + x.features)
        return self.out(x)


class PointTransformer(nn.Module):
    def __init__(self, input_channels, output_channels, grid_size, voxel_size, point_cloud_range, batching_info,
                 window_shape, drop_path_rate, depths, num_classes):
        super().__init__()
        self.sparse_shape = np.asarray(grid_size)[::-1]
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.batching_info = batching_info
        self.depths = depths
        self.window_shape = window_shape
        self.drop_path_rate = drop_path_rate
        self.norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        self.act_fn = nn.ReLU(inplace=True)

        self.conv_input = spconv.SparseSequential(
            spconv.SubMConv3d(input_channels, 48, 3, padding=1, bias=False, indice_key='subm1'),
            self.norm_fn(48), self.act_fn)

        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]     # stochastic depth decay
        chans = [48, 96, 192, 384]
        for i, c in enumerate(chans):
            block = nn.Sequential(
                SparseWindowPartitionLayer(batching_info[i], window_shape, self.sparse_shape[::-1] / (2 ** i)),
                SWFormerBlock(c, 8, depth=depths[i], drop_path=dpr[sum(depths[:i]):sum(depths[:i + 1])]))
            setattr(self, f'swformer_block{i + 1}', block)

        down = dict(norm_fn=self.norm_fn, act_fn=self.act_fn, stride=2, padding=1, conv_type='spconv')
        self.conv_down1 = ConvModule(48, 96, 3, indice_key='spconv2', **down)       # [1440,1440,64] -> [720,720,32]
        self.conv_down2 = ConvModule(96, 192, 3, indice_key='spconv3', **down)      # -> [360,360,16]
        self.conv_down3 = ConvModule(192, 384, 3, indice_key='spconv4', **down)     # -> [180,180,8]

        self.up4 = UpBlock(384, 192, self.norm_fn, self.act_fn, conv_type='inverseconv', layer_id=4)
        self.up3 = UpBlock(192, 96, self.norm_fn, self.act_fn, conv_type='inverseconv', layer_id=3)
        self.up2 = UpBlock(96, 48, self.norm_fn, self.act_fn, conv_type='inverseconv', layer_id=2)
        self.up1 = UpBlock(48, output_channels, self.norm_fn, self.act_fn, conv_type='subm', layer_id=1)
        self.up4.feeds_upblock = self.up3.feeds_upblock = self.up2.feeds_upblock = True

        self.aux_voxel_classifier = nn.Sequential(nn.Linear(384, num_classes, bias=False))
        self.voxel_classifier = nn.Sequential(nn.Linear(output_channels, num_classes, bias=False))

    def _stage(self, block, x):
        info = block[0](x)
        return replace_feature(x, block[1](info))

    def forward(self, batch_dict):
        voxel_features, voxel_coords = batch_dict['voxel_features'], batch_dict['voxel_coords']
        x = spconv.SparseConvTensor(features=voxel_features, indices=voxel_coords.int(),
                                    spatial_shape=self.sparse_shape, batch_size=batch_dict['batch_size'])
        # encoder
        x_conv1 = self._stage(self.swformer_block1, self.conv_input(x))
        x_conv2 = self._stage(self.swformer_block2, self.conv_down1(x_conv1))
        x_conv3 = self._stage(self.swformer_block3, self.conv_down2(x_conv2))
        x_conv4 = self._stage(self.swformer_block4, self.conv_down3(x_conv3))
        # auxiliary branch
        batch_dict['aux_voxel_out'] = linear_in(self.aux_voxel_classifier[0], x_conv4.features)
        batch_dict['aux_voxel_coords'] = x_conv4.indices
        # decoder
        x_conv4 = self.up4(x_conv4, x_conv4)
        x_conv3 = self.up3(x_conv4, x_conv3)
        x_conv2 = self.up2(x_conv3, x_conv2)
        x_conv1 = self.up1(x_conv2, x_conv1)

        batch_dict['voxel_features'] = x_conv1.features
        batch_dict['voxel_coords'] = x_conv1.indices
        batch_dict['voxel_out'] = linear_in(self.voxel_classifier[0], x_conv1.features)
        return batch_dict
