"""Segformer segmentor (seg3d/models/segmentors/segformer.py:12-146): point MLP -> VFE scatter -> PointTransformer ->
voxel_to_point gather -> fusion MLP -> SE -> classifier.  Same module tree / state_dict keys as the reference.

Two additions for the B200 path, both optional so that reference-shaped batch dicts keep working:
  * if the batch has no 'point_voxel_ids' the raw points are voxelized on the GPU as the first step of forward
     -- the reference does this per frame in DataLoader workers;
  * ``compute_dtype=torch.bfloat16`` runs the backbone in bf16 (tcgen05 sparse conv, bf16 attention) and the point
    MLPs under bf16 autocast; voxelization and pooling always stay fp32.
"""
import os
from collections import OrderedDict
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

from ..core.voxel import voxelize_batch, _geometry
from ..ops import voxel_to_point
from ..ops.mlp_chain import NONE, RELU, MlpChain
from .backbones import PointTransformer
from .layers import FlattenSELayer
from .voxel_encoders import VFE


_USE_MLP_CHAIN = os.environ.get('OS3D_MLP_CHAIN', '1') != '0'


class FoldedMLP(object):
    """Inference form of a point-wise MLP written as nn.Sequential([BN], (Linear, [BN], [ReLU], [Dropout])*)
    (segformer.py:21-32,58-76): eval-mode BatchNorm1d is affine, so it folds into the neighbouring Linear
    (y = s*(Wx + b) + t  ->  W' = s*W, b' = s*b + t; a leading BN folds into the first Linear's columns), and the ReLU
    runs in the GEMM epilogue (cuBLASLt bias + ReLU).  The module tree / state_dict stay the reference's; this is
    only how forward() executes them when not training.  Rebuilt when any parameter or buffer changes."""

    def __init__(self, seq):
        self.seq, self._tag, self._layers = seq, None, None

    def _build(self):
        mods = list(self.seq)
        tag = tuple((t.data_ptr(), t._version) for m in mods for t in list(m.parameters()) + list(m.buffers()))
        if tag == self._tag:
            return self._layers

        def affine(bn):
            s = bn.weight.detach().float() * torch.rsqrt(bn.running_var.float() + bn.eps)
            return s, bn.bias.detach().float() - bn.running_mean.float() * s

        layers, pre, i = [], None, 0
        if isinstance(mods[0], nn.BatchNorm1d):
            pre, i = affine(mods[0]), 1
        while i < len(mods):
            lin = mods[i]
            if not isinstance(lin, nn.Linear):
                raise NotImplementedError(f'FoldedMLP: unexpected {type(lin).__name__} at position {i}')
            w = lin.weight.detach().float()
            b = lin.bias.detach().float() if lin.bias is not None else w.new_zeros(w.shape[0])
            if pre is not None:                       # Linear(s*x + t) = (W*s) x + (W t + b)
                b = b + w @ pre[1]
                w = w * pre[0][None, :]
                pre = None
            i += 1
            relu = False
            while i < len(mods) and not isinstance(mods[i], nn.Linear):
                m = mods[i]
                if isinstance(m, nn.BatchNorm1d):
                    if relu:
                        raise NotImplementedError('FoldedMLP: BatchNorm after ReLU')
                    sc, sh = affine(m)
                    w, b = w * sc[:, None], b * sc + sh
                elif isinstance(m, nn.ReLU):
                    relu = True
                elif not isinstance(m, nn.Dropout):
                    raise NotImplementedError(f'FoldedMLP: unexpected {type(m).__name__}')
                i += 1
            layers.append((w.contiguous(), b.contiguous(), relu, {}))
        self._tag, self._layers = tag, layers
        return layers

    def _chain(self, x):
        """The folded layers as one os3d_mlp_chain_bf16 launch (activations on chip), or None when the chain does not fit
        the kernel.  An fp32 input goes through the kernel's fp32 front layer (raw metric coordinates do not survive a
        bf16 cast)."""
        layers = self._build()
        key = (self._tag, x.dtype)
        hit = self.__dict__.get('_chain_hit')
        if hit is None or hit[0] != key:
            chain = None
            spec = [(w, b, RELU if relu else NONE) for w, b, relu, _ in layers]
            front = None
            if x.dtype == torch.float32:
                front, spec = spec[0], spec[1:]
                if front[0].shape[0] != 64 or front[0].shape[1] > 16:
                    spec = []
            if 2 <= len(spec) <= 4 and MlpChain.fits([tuple(w.shape) for w, _, _ in spec], front is not None):
                chain = MlpChain(spec, front=front)
            hit = self.__dict__['_chain_hit'] = (key, chain)
        return hit[1]

    def __call__(self, x, dtype, out=None):
        """x: [N, C] fp32 or bf16.  The first layer of an fp32 input runs in fp32, the rest in ``dtype``.  ``out``:
        optional preallocated destination (a column slice of a wider buffer) for the bf16 chain kernel."""
        if dtype == torch.bfloat16 and _USE_MLP_CHAIN and x.dtype in (torch.float32, torch.bfloat16):
            chain = self._chain(x)
            if chain is not None:
                return chain(x, out=out)
        y = self._layerwise(x, dtype)
        if out is not None:
            out.copy_(y)
            return out
        return y

    def _layerwise(self, x, dtype):
        for li, (w, b, relu, cast) in enumerate(self._build()):
            dt = torch.float32 if li == 0 and x.dtype == torch.float32 else dtype
            if dt not in cast:
                cast[dt] = (w.to(dt).t().contiguous(), b.to(dt))
            wt, bb = cast[dt]
            x = x.to(dt)
            x = torch._addmm_activation(bb, x, wt) if relu else torch.addmm(bb, x, wt)
        return x.to(dtype)


class Segformer(nn.Module):
    def __init__(self, dataset, batching_info, window_shape, depths, drop_path_rate, compute_dtype=torch.float32):
        super().__init__()
        dim_point = dataset.dim_point + (2 if dataset.use_cylinder else 0)
        self.compute_dtype = compute_dtype
        self.voxel_size, self.point_cloud_range = dataset.voxel_size, dataset.point_cloud_range

        self.point_feature_channel = 64
        self.point_encoder = nn.Sequential(
            nn.BatchNorm1d(dim_point),
            nn.Linear(dim_point, 64, bias=False), nn.BatchNorm1d(64), nn.ReLU(inplace=True),
            nn.Linear(64, 128, bias=False), nn.BatchNorm1d(128), nn.ReLU(inplace=True),
            nn.Linear(128, 256, bias=False), nn.BatchNorm1d(256), nn.ReLU(inplace=True),
            nn.Linear(256, self.point_feature_channel))

        self.use_multi_sweeps = dataset.use_multi_sweeps
        self.vfe = VFE(dim_point, reduce='mean') if self.use_multi_sweeps else VFE(self.point_feature_channel, reduce='max')
        self.scatter = VFE(3, reduce='mean')          # present (parameter-free) in the reference too, never called

        self.voxel_in_feature_channel = self.vfe.voxel_feature_channel
        self.voxel_feature_channel = 32
        self.point_transformer = PointTransformer(self.voxel_in_feature_channel, self.voxel_feature_channel,
                                                  dataset.grid_size, dataset.voxel_size, dataset.point_cloud_range,
                                                  batching_info=batching_info, window_shape=window_shape, depths=depths,
                                                  drop_path_rate=drop_path_rate, num_classes=dataset.num_classes)
        if dataset.use_image_feature:
            raise NotImplementedError('image-feature fusion is off in every reference config (config.py:17)')
        self.use_image_feature = False
        self.image_feature_channel = 0

        self.fusion_feature_channel = 64
        self.fusion_encoder = nn.Sequential(
            nn.Linear(self.point_feature_channel + self.voxel_feature_channel, 256, bias=False), nn.BatchNorm1d(256),
            nn.ReLU(inplace=True),
            nn.Linear(256, 128, bias=False), nn.BatchNorm1d(128), nn.ReLU(inplace=True),
            nn.Linear(128, self.fusion_feature_channel, bias=False), nn.BatchNorm1d(self.fusion_feature_channel),
            nn.ReLU(inplace=True))
        self.se = FlattenSELayer(self.fusion_feature_channel)
        self.classifier = nn.Sequential(nn.Linear(self.fusion_feature_channel, 64, bias=False), nn.BatchNorm1d(64),
                                        nn.ReLU(True), nn.Dropout(0.3),
                                        nn.Linear(64, dataset.num_classes, bias=False))
        self.weight_initialization()

    def _folded(self, name):
        cache = self.__dict__.setdefault('_folded_mlps', {})
        if name not in cache:
            cache[name] = FoldedMLP(getattr(self, name))
        return cache[name]

    def weight_initialization(self):
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv2d)):
                nn.init.kaiming_normal_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
                nn.init.constant_(m.weight, 1)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.weight, 1.0)
                nn.init.constant_(m.bias, 0)

    def forward(self, batch_dict):
        raw = batch_dict['points']
        bf16 = self.compute_dtype == torch.bfloat16
        if 'point_voxel_ids' not in batch_dict:            # voxelize in the forward
            coords, pvid = voxelize_batch(raw, self.voxel_size, self.point_cloud_range, has_batch=True)
            batch_dict['voxel_coords'], batch_dict['point_voxel_ids'] = coords, pvid
        points = raw[:, 1:]
        point_voxel_ids = batch_dict['point_voxel_ids']
        num_voxels = batch_dict['voxel_coords'].shape[0]
        if self.use_multi_sweeps:
            cur_point_indices = points[:, 3] == 0
            cur_points = points[cur_point_indices]
        else:
            cur_points = points
        fold = not self.training and not torch.is_grad_enabled()   # inference: BatchNorm folded, ReLU in the GEMM epilogue
        if fold:
            point_per_features = self._folded('point_encoder')(cur_points, self.compute_dtype)
        else:
            with torch.autocast('cuda', dtype=torch.bfloat16, enabled=bf16):
                point_per_features = self.point_encoder(cur_points)

        # encode voxel features (fp32 pooling)
        if self.use_multi_sweeps:
            voxel_features = self.vfe(points, point_voxel_ids, num_voxels)
        else:
            voxel_features = self.vfe(point_per_features, point_voxel_ids, num_voxels)
        batch_dict['voxel_features'] = voxel_features.to(self.compute_dtype)
        batch_dict = self.point_transformer(batch_dict)

        # point features from the encoded voxel features
        ids = point_voxel_ids[cur_point_indices] if self.use_multi_sweeps else point_voxel_ids
        point_voxel_features = voxel_to_point(batch_dict['voxel_features'], ids)
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=bf16 and not fold):
            fused = torch.cat([point_per_features.to(point_voxel_features.dtype), point_voxel_features], dim=1)
            fused = self._folded('fusion_encoder')(fused, self.compute_dtype) if fold else self.fusion_encoder(fused)
            batch_idx = raw[:, 0][cur_point_indices] if self.use_multi_sweeps else raw[:, 0]
            fused = self.se.residual_forward(fused, batch_idx, batch_dict['batch_size'])     # fused + se(fused)
            point_out = self._folded('classifier')(fused, self.compute_dtype) if fold else self.classifier(fused)

        result = OrderedDict()
        result['point_out'] = point_out
        result['voxel_out'] = batch_dict['voxel_out']
        result['aux_voxel_out'] = batch_dict['aux_voxel_out']
        result['voxel_coords'] = batch_dict['voxel_coords']
        result['aux_voxel_coords'] = batch_dict['aux_voxel_coords']
        return result


def default_batching_info():
    """cfg.MODEL.BATCHING_INFO with integer level keys, as build_segmentor converts it
    (seg3d/utils/config.py:42-67, seg3d/models/builder.py:10-15)."""
    spec = [[(16, 0, 16), (64, 16, 64), (256, 64, 256), (800, 256, 100000)],
            [(32, 0, 32), (128, 32, 128), (512, 128, 512), (800, 512, 100000)],
            [(64, 0, 64), (160, 64, 160), (384, 160, 384), (800, 384, 100000)],
            [(128, 0, 128), (256, 128, 256), (512, 256, 512), (800, 512, 100000)]]
    return [{i: {'max_tokens': t, 'batching_range': [lo, hi]} for i, (t, lo, hi) in enumerate(level)} for level in spec]


DATASET_CONFIGS = {
    # configs/waymo_one_sweep.yaml + seg3d/utils/config.py defaults
    'waymo_one_sweep': dict(voxel_size=[0.1, 0.1, 0.1], point_cloud_range=[-72, -72, -2, 72, 72, 4.4],
                            use_cylinder=False, use_multi_sweeps=False, num_sweeps=1),
    # configs/waymo_one_sweep_cylinder.yaml:2-4
    'waymo_one_sweep_cylinder': dict(voxel_size=[0.05, 0.012, 0.1], point_cloud_range=[0, -3.1415926, -2, 75.2, 3.1415926, 5.2],
                                     use_cylinder=True, use_multi_sweeps=False, num_sweeps=1),
    # configs/waymo_multi_sweeps.yaml:2-4
    'waymo_multi_sweeps': dict(voxel_size=[0.1, 0.1, 0.1], point_cloud_range=[-72, -72, -2, 72, 72, 4.4],
                               use_cylinder=False, use_multi_sweeps=True, num_sweeps=3),
}


def dataset_spec(name):
    """The handful of dataset properties Segformer reads (waymo_dataset.py:51-77)."""
    c = DATASET_CONFIGS[name]
    vs, pcr, grid = _geometry(c['voxel_size'], c['point_cloud_range'])
    return SimpleNamespace(name=name, dim_point=6, use_cylinder=c['use_cylinder'], use_multi_sweeps=c['use_multi_sweeps'],
                           num_sweeps=c['num_sweeps'], use_image_feature=False, dim_image_feature=28, num_classes=22,
                           grid_size=grid, voxel_size=vs, point_cloud_range=pcr)


def build_segformer(config='waymo_one_sweep', compute_dtype=torch.float32, depths=(3, 4, 8, 3), window_shape=(10, 10, 8),
                    drop_path_rate=0.3, batching_info=None, seed=0):
    """build_segmentor for MODEL.SEGMENTOR == 'segformer' (builder.py:8-17) with random-init weights."""
    torch.manual_seed(seed)
    ds = dataset_spec(config)
    return Segformer(ds, batching_info or default_batching_info(), list(window_shape), list(depths), drop_path_rate,
                     compute_dtype=compute_dtype)
