"""spconv is absent (parity unpinned against it): pin the oracle's kernel-map + sparse-conv semantics
 against torch's dense conv3d / conv_transpose3d on a densified grid."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle


def _random_sites(rng, batch, shape, n):
    out = []
    for b in range(batch):
        c = np.unique(rng.integers(0, shape, (n, 3)), axis=0)
        c = c[rng.permutation(len(c))]
        out.append(np.pad(c, ((0, 0), (1, 0)), constant_values=b))
    return np.concatenate(out).astype(np.int32)


def _densify(idx, feats, batch, shape):
    d = torch.zeros(batch, feats.shape[1], *shape, dtype=feats.dtype)
    d[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]] = feats
    return d


def _w_dense(w):                     # spconv [Cout,kz,ky,kx,Cin] -> torch conv3d [Cout,Cin,kz,ky,kx]
    return w.permute(0, 4, 1, 2, 3).contiguous()


@pytest.mark.parametrize('shape', [(8, 12, 10), (5, 7, 9)])
def test_subm_matches_dense_conv3d(shape):
    rng = np.random.default_rng(0)
    torch.manual_seed(0)
    idx = _random_sites(rng, 2, shape, 150)
    cin, cout = 5, 7
    x = torch.randn(idx.shape[0], cin, dtype=torch.float64)
    w = torch.randn(cout, 3, 3, 3, cin, dtype=torch.float64)
    b = torch.randn(cout, dtype=torch.float64)
    nbr, pairs = oracle.subm_map(idx, shape)
    assert pairs == (nbr >= 0).sum()
    assert np.array_equal(nbr[:, 13], np.arange(idx.shape[0]))          # centre offset = identity
    y = oracle.sparse_conv(x, nbr, w, b)
    li = torch.from_numpy(idx).long()
    dense = F.conv3d(_densify(li, x, 2, shape), _w_dense(w), b, padding=1)
    ref = dense[li[:, 0], :, li[:, 1], li[:, 2], li[:, 3]]
    torch.testing.assert_close(y, ref, rtol=1e-10, atol=1e-10)
    # C loop (f32 in, f64 accumulate) agrees with the torch restatement
    yc = oracle.sparse_conv_c(x.float().numpy(), nbr, w.float().numpy(), b.float().numpy())
    np.testing.assert_allclose(yc, y.float().numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('shape', [(8, 12, 10), (7, 9, 11)])
def test_strided_and_inverse_match_dense(shape):
    rng = np.random.default_rng(1)
    torch.manual_seed(1)
    idx = _random_sites(rng, 2, shape, 120)
    cin, cout = 4, 6
    x = torch.randn(idx.shape[0], cin, dtype=torch.float64)
    w = torch.randn(cout, 3, 3, 3, cin, dtype=torch.float64)
    out_idx, out_shape, fwd, inv, pairs = oracle.strided_map(idx, shape)
    assert out_shape.tolist() == [(s + 2 - 3) // 2 + 1 for s in shape]
    assert pairs == (fwd >= 0).sum() == (inv >= 0).sum()
    # canonical order: ascending linear index
    lin = ((out_idx[:, 0].astype(np.int64) * out_shape[0] + out_idx[:, 1]) * out_shape[1] + out_idx[:, 2]) * out_shape[2] + out_idx[:, 3]
    assert (np.diff(lin) > 0).all()
    # the two tables hold the same (k, in, out) pairs
    a = oracle.pairs_from_table(fwd)                     # (k, in_row, out_row)
    b_ = oracle.pairs_from_table(inv)[:, [0, 2, 1]]      # (k, out_row, in_row) -> (k, in, out)
    assert np.array_equal(a, b_[np.lexsort((b_[:, 2], b_[:, 1], b_[:, 0]))])
    # forward strided conv == dense conv3d(stride 2, pad 1) on active output sites; active = any input in field
    li, lo = torch.from_numpy(idx).long(), torch.from_numpy(out_idx).long()
    dx = _densify(li, x, 2, shape)
    dense = F.conv3d(dx, _w_dense(w), None, stride=2, padding=1)
    occ = F.conv3d(_densify(li, torch.ones(idx.shape[0], 1, dtype=torch.float64), 2, shape),
                   torch.ones(1, 1, 3, 3, 3, dtype=torch.float64), None, stride=2, padding=1)
    act = torch.nonzero(occ[:, 0] > 0)
    assert torch.equal(act, lo)                           # same sites, same (ascending) order
    y = oracle.sparse_conv(x, fwd, w)
    torch.testing.assert_close(y, dense[lo[:, 0], :, lo[:, 1], lo[:, 2], lo[:, 3]], rtol=1e-10, atol=1e-10)
    # inverse conv == conv_transpose3d masked to the fine sites
    c2 = 3
    winv = torch.randn(c2, 3, 3, 3, cout, dtype=torch.float64)        # [Cout'=c2, k, Cin'=cout]
    z = oracle.sparse_conv(y, inv, winv)
    dy = _densify(lo, y, 2, tuple(out_shape.tolist()))
    # conv_transpose3d weight [Cin', Cout', kz,ky,kx]
    wt = winv.permute(4, 0, 1, 2, 3).contiguous()
    opad = [shape[d] - ((out_shape[d] - 1) * 2 - 2 + 3) for d in range(3)]
    dz = F.conv_transpose3d(dy, wt, None, stride=2, padding=1, output_padding=opad)
    torch.testing.assert_close(z, dz[li[:, 0], :, li[:, 1], li[:, 2], li[:, 3]], rtol=1e-10, atol=1e-10)


def test_ingroup_rank():
    g = np.array([3, 1, 3, 3, 0, 1], np.int64)
    assert oracle.ingroup_rank(g).tolist() == [0, 0, 1, 2, 0, 1]
