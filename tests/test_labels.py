"""Loss-side label preparation: per-voxel majority labels and the coarse-level label
transfer.  The oracle is pinned against the REFERENCE's own WaymoDataset.prepare_voxel_labels / get_voxel_centers
(tests/golden/voxel_labels.npz, produced by tests/golden/make_golden_labels.py); the CUDA ops must match bit for bit."""
import os

import numpy as np
import pytest
import torch


@pytest.fixture(scope='module')
def g(golden_dir):
    return np.load(os.path.join(golden_dir, 'voxel_labels.npz'))


def test_oracle_majority_labels_match_reference(g):
    from oracle import oracle
    out = oracle.voxel_majority_labels(g['point_voxel_ids'], g['point_labels'], g['voxel_coords'].shape[0])
    assert out.dtype == np.uint8 and np.array_equal(out, g['voxel_labels'])
    assert (out == 255).any() and (g['point_voxel_ids'] == -1).any()          # both skip rules are exercised


def test_oracle_voxel_centers_match_reference(g):
    from oracle import oracle
    vs, pcr = [0.1, 0.1, 0.1], [-72, -72, -2, 72, 72, 4.4]
    assert np.array_equal(oracle.voxel_centers(g['voxel_coords'], 1.0, vs, pcr), g['centers'])
    assert np.array_equal(oracle.voxel_centers(g['aux_coords'], 8.0, vs, pcr), g['aux_centers'])


def test_oracle_majority_ties_take_the_lowest_label():
    from oracle import oracle
    ids = np.array([0, 0, 0, 0, 1, 1, 2, -1, 3, 3])
    lab = np.array([5, 2, 5, 2, 255, 7, 255, 1, 255, 255])
    assert oracle.voxel_majority_labels(ids, lab, 5).tolist() == [2, 7, 255, 255, 255]


@pytest.mark.gpu
def test_gpu_majority_labels_match_reference_and_oracle(g):
    from openseg3d_b200.ops import voxel_majority_labels
    from oracle import oracle
    m = g['voxel_coords'].shape[0]
    out = voxel_majority_labels(torch.from_numpy(g['point_voxel_ids']).cuda(), torch.from_numpy(g['point_labels']).cuda(), m)
    assert out.dtype == torch.uint8 and np.array_equal(out.cpu().numpy(), g['voxel_labels'])
    # BASELINE size: 8 frames, random labels with a 10 % ignore share, against the oracle
    from openseg3d_b200 import synthetic
    from openseg3d_b200.core import voxelize_batch
    pts, _ = synthetic.make_batch(list(range(8)), 1, False)
    coors, pvid = voxelize_batch(torch.from_numpy(pts).cuda(), [0.1, 0.1, 0.1], [-72, -72, -2, 72, 72, 4.4])
    rng = np.random.default_rng(0)
    lab = rng.integers(0, 22, pts.shape[0]).astype(np.uint8)
    lab[rng.random(pts.shape[0]) < 0.1] = 255
    out = voxel_majority_labels(pvid, torch.from_numpy(lab).cuda(), coors.shape[0])
    assert np.array_equal(out.cpu().numpy(), oracle.voxel_majority_labels(pvid.cpu().numpy(), lab, coors.shape[0]))
    # edge cases: no points, labels the 32-bin layout cannot order
    assert voxel_majority_labels(pvid[:0], torch.from_numpy(lab[:0]).cuda(), 3).tolist() == [255, 255, 255]
    with pytest.raises(RuntimeError):
        voxel_majority_labels(pvid[:4], torch.tensor([1, 2, 40, 3], dtype=torch.uint8).cuda(), coors.shape[0])


@pytest.mark.gpu
def test_gpu_aux_voxel_labels_match_oracle(g):
    """tools/train.py:86-104 on two frames: nearest level-1 centre per stride-8 voxel, inside the same frame."""
    from openseg3d_b200.ops import aux_voxel_labels
    from oracle import oracle
    vs, pcr = [0.1, 0.1, 0.1], [-72, -72, -2, 72, 72, 4.4]
    c1 = np.concatenate([np.pad(g['voxel_coords'], ((0, 0), (1, 0)), constant_values=0),
                         np.pad(g['voxel_coords'][::2], ((0, 0), (1, 0)), constant_values=1)]).astype(np.int32)
    c8 = np.concatenate([np.pad(g['aux_coords'], ((0, 0), (1, 0)), constant_values=0),
                         np.pad(g['aux_coords'][::3], ((0, 0), (1, 0)), constant_values=1)]).astype(np.int32)
    labels = np.concatenate([g['voxel_labels'], g['voxel_labels'][::2]])
    ref = oracle.aux_voxel_labels(labels, c1, c8, 2, vs, pcr)
    out = aux_voxel_labels(torch.from_numpy(labels).cuda(), torch.from_numpy(c1).cuda(), torch.from_numpy(c8).cuda(), 2, vs, pcr)
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', ['bf16', 'f32'])
@pytest.mark.parametrize('n,c', [(0, 23), (1, 23), (255, 23), (256, 23), (100003, 23), (5000, 22), (777, 7), (4097, 64)])
def test_predict_labels_matches_first_argmax(n, c, dtype):
    """tools/test.py:58: argmax over the class logits of each point; ties resolve to the lowest class index."""
    import torch
    from openseg3d_b200.ops import predict_labels
    torch.manual_seed(n + c)
    x = torch.randn(n, c)
    if dtype == 'bf16':
        x = x.bfloat16()                    # 8-bit mantissa: ties are common
    if n > 3:
        x[1] = x[1, 0]                      # a whole row of ties -> class 0
        x[2, c - 1] = x[2].max()            # a tie between an inner class and the last one -> the inner one
    xf = x.float()
    is_max = xf == xf.max(dim=1, keepdim=True).values
    expect = torch.where(is_max, torch.arange(c)[None].expand(n, c), c).min(dim=1).values.to(torch.uint8) if n else torch.empty(0, dtype=torch.uint8)
    got = predict_labels(x.cuda())
    assert got.dtype == torch.uint8 and got.shape == (n,)
    assert torch.equal(got.cpu(), expect)
