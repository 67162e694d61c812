"""knn_query: the oracle's serial restatement against numpy brute force (CPU), and the CUDA op against the oracle
(bit-exact indices and squared distances, ties included)."""
import numpy as np
import pytest
import torch


def _case(rng, sizes_xyz, sizes_q, lattice):
    pts = [rng.integers(0, 12, (n, 3)).astype(np.float32) * 0.5 if lattice else rng.normal(size=(n, 3)).astype(np.float32) * 5
           for n in sizes_xyz]
    qs = [rng.integers(0, 12, (n, 3)).astype(np.float32) * 0.5 if lattice else rng.normal(size=(n, 3)).astype(np.float32) * 5
          for n in sizes_q]
    return (np.concatenate(pts), np.concatenate(qs), np.cumsum(sizes_xyz).astype(np.int32), np.cumsum(sizes_q).astype(np.int32))


def test_oracle_knn_matches_brute_force():
    from oracle import oracle
    rng = np.random.default_rng(0)
    xyz, q, off, noff = _case(rng, [300, 150], [40, 25], lattice=False)
    idx, d2 = oracle.knn_query(4, xyz, q, off, noff)
    for i in range(q.shape[0]):
        b = 0 if i < noff[0] else 1
        s, e = (0, off[0]) if b == 0 else (off[0], off[1])
        ref = np.sort(((q[i] - xyz[s:e]).astype(np.float64) ** 2).sum(1))[:4]
        assert np.allclose(d2[i], ref, rtol=1e-5)
        assert np.all((idx[i] >= s) & (idx[i] < e)) and len(set(idx[i].tolist())) == 4
        assert np.all(np.diff(d2[i]) >= 0)


@pytest.mark.gpu
@pytest.mark.parametrize('nsample,lattice', [(1, False), (1, True), (3, False), (8, True), (16, False)])
def test_knn_query_matches_oracle(nsample, lattice):
    """lattice=True: many exactly tied distances -- the heap's tie behaviour must match the reference's algorithm."""
    from openseg3d_b200.ops import knn_query
    from oracle import oracle
    rng = np.random.default_rng(nsample)
    xyz, q, off, noff = _case(rng, [2500, 40, 3100], [300, 7, 260], lattice)
    idx, dist = knn_query(nsample, torch.from_numpy(xyz).cuda(), torch.from_numpy(q).cuda(), torch.from_numpy(off).cuda(),
                          torch.from_numpy(noff).cuda())
    ridx, rd2 = oracle.knn_query(nsample, xyz, q, off, noff)
    assert idx.dtype == torch.int32 and dist.dtype == torch.float32
    assert np.array_equal(idx.cpu().numpy(), ridx)
    assert np.array_equal(dist.cpu().numpy(), np.sqrt(rd2))


@pytest.mark.gpu
def test_knn_query_coarse_to_fine_label_transfer_shape():
    """The reference's only use (tools/train.py:86-104): nearest fine voxel centre of every coarse voxel centre."""
    from openseg3d_b200.ops import knn_query
    rng = np.random.default_rng(5)
    fine = torch.from_numpy(rng.uniform(-70, 70, (50000, 3)).astype(np.float32)).cuda()
    coarse = fine[::8].contiguous() + 0.01
    idx, dist = knn_query(1, fine, coarse, torch.tensor([30000, 50000], dtype=torch.int32).cuda(),
                          torch.tensor([3750, 6250], dtype=torch.int32).cuda())
    assert bool((idx.squeeze(1).long() == torch.arange(0, 50000, 8, device='cuda')).all())
    assert float(dist.max()) < 0.02
