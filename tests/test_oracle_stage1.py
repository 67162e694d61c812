"""Pins the oracle's stage-1 restatement against the reference's own code (golden vectors made by
tests/golden/make_golden.py from seg3d/core/voxel/voxel_generator.py and seg3d/ops/voxel_to_point)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from openseg3d_b200 import synthetic
from oracle import oracle


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize('name', ['cart_small', 'cyl_small', 'multi_small'])
def test_voxelize_matches_reference_numba(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f'voxelize_{name}.npz'))
    coors, ids = oracle.voxelize(g['points'], g['voxel_size'], g['pc_range'], has_batch=True)
    assert np.array_equal(oracle.grid_size(g['voxel_size'], g['pc_range']), g['grid_size'])
    assert coors.dtype == np.int32 and ids.dtype == np.int64
    assert np.array_equal(coors, g['coors'])
    assert np.array_equal(ids, g['point_voxel_ids'])
    assert (ids == -1).sum() >= 16          # the fixture carries out-of-range points


@pytest.mark.parametrize('name,sweeps,cyl', [('cart_full', 1, False), ('cyl_full', 1, True), ('multi_full', 3, False)])
def test_voxelize_full_frame_checksums(golden_dir, name, sweeps, cyl):
    sums = json.load(open(os.path.join(golden_dir, 'voxelize_full_checksums.json')))[name]
    from tests.golden_cfg import CFG
    cfg = CFG['cyl' if cyl else 'cart']
    pts, _ = synthetic.make_batch([0], sweeps, cyl)
    assert sha(pts) == sums['points'], 'synthetic generator drifted from the one the goldens were made with'
    coors, ids = oracle.voxelize(pts, cfg['voxel_size'], cfg['pc_range'], has_batch=True)
    assert coors.shape[0] == sums['m'] and pts.shape[0] == sums['n']
    assert sha(coors) == sums['coors']
    assert sha(ids) == sums['ids']


def test_grid_sizes():
    from tests.golden_cfg import CFG
    assert oracle.grid_size(**CFG['cart']).tolist() == [1440, 1440, 64]
    assert oracle.grid_size(**CFG['cyl']).tolist() == [1504, 524, 72]


def test_voxel_to_point_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, 'stage1_gather.npz'))
    out = oracle.voxel_to_point(torch.from_numpy(g['feats']), torch.from_numpy(g['ids']))
    assert torch.equal(out, torch.from_numpy(g['out']))


def test_scatter_semantics():
    feats = torch.tensor([[1., -2.], [3., 4.], [-5., -6.], [7., 8.]])
    idx = torch.tensor([1, -1, 1, 0])
    assert torch.equal(oracle.scatter_reduce(feats, idx, 'max'), torch.tensor([[7., 8.], [1., -2.]]))
    assert torch.equal(oracle.scatter_reduce(feats, idx, 'mean'), torch.tensor([[7., 8.], [-2., -4.]]))
    # empty rows: max -> 0 (torch_scatter fills rows nothing was scattered to with 0)
    idx2 = torch.tensor([2, -1, 2, 0])
    assert torch.equal(oracle.scatter_reduce(feats, idx2, 'max')[1], torch.zeros(2))
    counts = torch.tensor([1, 2])
    assert torch.allclose(oracle.voxel_avg_pooling(feats, torch.tensor([1, 5, 1, 0]), counts),
                          torch.tensor([[7., 8.], [-2., -4.]]))


@pytest.mark.parametrize('name', ['small', 'ragged', 'heavy'])
def test_voxel_avg_pooling_matches_reference_cpu_function(golden_dir, name):
    """oracle.voxel_avg_pooling against the output of the reference's own voxel_pooling_forward_cpu
    (seg3d/ops/voxel_pooling/src/voxel_pooling.cpp:5-23, compiled from /root/reference by oracle/build_ref.py;
    vectors made by tests/golden/make_golden_pooling.py): ids outside [0, M) skipped, an empty voxel, a voxel holding
    80 % of the points.  Same divide-then-accumulate order -> bit-exact."""
    g = np.load(os.path.join(golden_dir, 'voxel_avg_pooling.npz'))
    out = oracle.voxel_avg_pooling(torch.from_numpy(g[f'{name}_feats']), torch.from_numpy(g[f'{name}_ids']),
                                   torch.from_numpy(g[f'{name}_counts']))
    assert torch.equal(out, torch.from_numpy(g[f'{name}_out']))
    # the mean reduction of VFE (scatter 'mean', vfe.py:24-25) is the same quantity up to summation order
    ids = torch.from_numpy(g[f'{name}_ids']).long()
    ids = torch.where((ids >= 0) & (ids < g[f'{name}_counts'].shape[0]), ids, torch.full_like(ids, -1))
    mean = oracle.scatter_reduce(torch.from_numpy(g[f'{name}_feats']), ids, 'mean')
    assert torch.allclose(mean[:out.shape[0]], out[:mean.shape[0]], rtol=1e-5, atol=1e-5)


def test_voxel_avg_pooling_against_live_reference_build():
    """Where the reference sources exist (the build container), the compiled reference function itself is the checker."""
    from oracle import build_ref
    if not os.path.exists(build_ref.REF_SRC) and not os.path.exists(build_ref.out_path()):
        pytest.skip('no /root/reference and no prebuilt oracle/_ref here')
    ext = build_ref.load()
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(777, 24, generator=g)
    ids = torch.randint(-1, 41, (777,), generator=g, dtype=torch.int32)
    counts = torch.bincount(ids[(ids >= 0) & (ids < 40)].long(), minlength=40).int()
    ref = ext.voxel_pooling_forward_cpu(feats, ids, counts)
    assert torch.equal(oracle.voxel_avg_pooling(feats, ids, counts), ref)
    # backward (voxel_pooling.cpp:25-43): autograd of the restatement == the reference's hand-written CPU backward
    x = feats.clone().requires_grad_()
    top = torch.randn(ref.shape, generator=g)
    oracle.voxel_avg_pooling(x, ids, counts).backward(top)
    assert torch.equal(x.grad, ext.voxel_pooling_backward_cpu(top, ids, counts, feats.shape[0]))


def test_folded_point_mlp_equals_sequential():
    """Inference-time BatchNorm folding of the point-wise MLPs (segformer.py:21-32,58-76) is exact up to fp32
    rounding: same nn.Sequential, same state_dict, evaluated both ways on CPU."""
    import torch
    import torch.nn as nn
    from openseg3d_b200.models.segmentors import FoldedMLP
    torch.manual_seed(0)
    seq = nn.Sequential(nn.BatchNorm1d(6), nn.Linear(6, 64, bias=False), nn.BatchNorm1d(64), nn.ReLU(True),
                        nn.Linear(64, 128, bias=False), nn.BatchNorm1d(128), nn.ReLU(True), nn.Dropout(0.3),
                        nn.Linear(128, 22)).eval()
    for m in seq:
        if isinstance(m, nn.BatchNorm1d):
            m.running_mean.normal_()
            m.running_var.uniform_(0.5, 2)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_()
    x = torch.randn(1000, 6) * 30
    folded = FoldedMLP(seq)
    with torch.no_grad():
        ref = seq(x)
        assert (ref - folded(x, torch.float32)).abs().max() < 1e-4 * ref.abs().max()
        seq[2].running_mean.add_(1.0)                 # buffers changed -> the folded copy is rebuilt
        assert (seq(x) - folded(x, torch.float32)).abs().max() < 1e-4 * ref.abs().max()
