"""GPU parity, stage 1: voxelize (bit-exact vs the reference's numba output and the oracle), scatter max (bit-exact),
scatter mean / avg pooling (fp32 rel 1e-4: atomic summation order), gather (bit-exact)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from openseg3d_b200 import synthetic
from tests.golden_cfg import CFG

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize('name', ['cart_small', 'cyl_small', 'multi_small'])
def test_voxelize_matches_reference_golden(golden_dir, name):
    from openseg3d_b200.core import voxelize_batch
    g = np.load(os.path.join(golden_dir, f'voxelize_{name}.npz'))
    coors, ids = voxelize_batch(torch.from_numpy(g['points']).cuda(), g['voxel_size'], g['pc_range'])
    assert coors.dtype == torch.int32 and ids.dtype == torch.int64
    assert np.array_equal(coors.cpu().numpy(), g['coors'])
    assert np.array_equal(ids.cpu().numpy(), g['point_voxel_ids'])


@pytest.mark.parametrize('name,sweeps,cyl', [('cart_full', 1, False), ('cyl_full', 1, True), ('multi_full', 3, False)])
def test_voxelize_full_frames_checksum_and_oracle(golden_dir, name, sweeps, cyl):
    from openseg3d_b200.core import voxelize_batch, VoxelGenerator
    from oracle import oracle
    cfg = CFG['cyl' if cyl else 'cart']
    sums = json.load(open(os.path.join(golden_dir, 'voxelize_full_checksums.json')))[name]
    pts, _ = synthetic.make_batch([0], sweeps, cyl)
    coors, ids = voxelize_batch(torch.from_numpy(pts).cuda(), cfg['voxel_size'], cfg['pc_range'])
    assert sha(coors.cpu().numpy()) == sums['coors'] and sha(ids.cpu().numpy()) == sums['ids']
    # reference-shaped API: numpy in, numpy out, zyx coords, int32 ids
    gen = VoxelGenerator(cfg['voxel_size'], cfg['pc_range'])
    c3, i32 = gen.generate(pts[:, 1:])
    assert c3.dtype == np.int32 and i32.dtype == np.int32
    assert np.array_equal(c3, coors.cpu().numpy()[:, 1:]) and np.array_equal(i32, ids.cpu().numpy())
    # batch of 3 frames against the oracle (first-occurrence order + cumulative offsets)
    pts3, _ = synthetic.make_batch([3, 4, 5], sweeps, cyl)
    co, io = oracle.voxelize(pts3, cfg['voxel_size'], cfg['pc_range'])
    cg, ig = voxelize_batch(torch.from_numpy(pts3).cuda(), cfg['voxel_size'], cfg['pc_range'])
    assert np.array_equal(cg.cpu().numpy(), co) and np.array_equal(ig.cpu().numpy(), io)


def test_voxelize_edge_cases():
    from openseg3d_b200.core import voxelize_batch
    cfg = CFG['cart']
    empty = torch.zeros((0, 7), dtype=torch.float32, device='cuda')
    c, i = voxelize_batch(empty, cfg['voxel_size'], cfg['pc_range'])
    assert c.shape == (0, 4) and i.shape == (0,)
    # all points outside, and all points in ONE voxel (maximal collision)
    out = torch.tensor([[0, 100., 0, 0, 0, 0, 0], [0, 0, 0, 9., 0, 0, 0]], device='cuda')
    c, i = voxelize_batch(out, cfg['voxel_size'], cfg['pc_range'])
    assert c.shape[0] == 0 and i.tolist() == [-1, -1]
    same = torch.zeros((5000, 7), device='cuda')
    same[:, 1:4] = torch.tensor([1.234, -5.678, 0.5])
    c, i = voxelize_batch(same, cfg['voxel_size'], cfg['pc_range'])
    assert c.shape[0] == 1 and bool((i == 0).all())


def test_cart2polar_device_flip_rate():
    """atan2f vs numpy differs in the last ulp: report the voxel flip rate, bound it."""
    from openseg3d_b200.core import voxelize_batch, cart2polar_rows
    cfg = CFG['cyl']
    raw, _ = synthetic.make_batch([0], 1, False)
    host_rows, _ = synthetic.make_batch([0], 1, True)
    dev_rows = cart2polar_rows(torch.from_numpy(raw).cuda())
    np.testing.assert_allclose(dev_rows.cpu().numpy(), host_rows, rtol=1e-6, atol=1e-6)
    _, ids_h = voxelize_batch(torch.from_numpy(host_rows).cuda(), cfg['voxel_size'], cfg['pc_range'])
    ch, _ = voxelize_batch(torch.from_numpy(host_rows).cuda(), cfg['voxel_size'], cfg['pc_range'])
    cd, ids_d = voxelize_batch(dev_rows, cfg['voxel_size'], cfg['pc_range'])
    # compare per-point voxel coordinates
    kh = ch[ids_h.clamp(min=0)].cpu().numpy()
    kd = cd[ids_d.clamp(min=0)].cpu().numpy()
    flips = float((kh != kd).any(axis=1).mean())
    assert flips < 1e-3, flips


@pytest.mark.parametrize('c', [64, 6, 3])
def test_scatter_max_bit_exact_and_mean(c):
    from openseg3d_b200.ops import scatter_max, scatter_mean
    from oracle import oracle
    torch.manual_seed(c)
    n, m = 50000, 7000
    feats = torch.randn(n, c)
    ids = torch.randint(-1, m, (n,))
    ids[:m] = torch.arange(m)                      # every voxel owns a point
    mx = scatter_max(feats.cuda(), ids.cuda(), m, fix_empty=False).cpu()
    assert torch.equal(mx, oracle.scatter_reduce(feats, ids, 'max'))
    mean = scatter_mean(feats.cuda(), ids.cuda(), m).cpu()
    torch.testing.assert_close(mean, oracle.scatter_reduce(feats, ids, 'mean'), rtol=1e-4, atol=1e-6)
    # empty rows -> 0 (torch_scatter semantics), output sized by max id + 1
    ids2 = ids.clone()
    ids2[ids2 == 5] = -1
    mx2 = scatter_max(feats.cuda(), ids2.cuda()).cpu()
    assert torch.equal(mx2, oracle.scatter_reduce(feats, ids2, 'max')) and bool((mx2[5] == 0).all())


@pytest.mark.parametrize('impl', ['sorted', 'atomic'])
@pytest.mark.parametrize('c,idt', [(64, torch.int32), (64, torch.int64), (8, torch.int32), (128, torch.int64)])
def test_scatter_max_bf16_inputs_exact(c, idt, impl, monkeypatch):
    """bf16 inference path of VFE (vfe.py:24-25): bf16 point features reduced as they are, bf16 voxel features out --
    exactly the maximum of the inputs, empty rows 0 (or -inf without fix_empty), ids -1 skipped."""
    from openseg3d_b200.ops import scatter_max, pooling
    from oracle import oracle
    monkeypatch.setattr(pooling, '_MAX_IMPL', impl)
    torch.manual_seed(c)
    n, m = 60000, 9000
    feats = torch.randn(n, c).bfloat16()
    ids = torch.randint(-1, m, (n,))
    ids[ids == 7] = -1                              # voxel 7 stays empty
    ids[0] = m - 1                                  # the oracle sizes its output by the largest id
    ref = oracle.scatter_reduce(feats.float(), ids, 'max')
    got = scatter_max(feats.cuda(), ids.to(idt).cuda(), m)
    assert got.dtype == torch.bfloat16
    assert torch.equal(got.float().cpu(), ref) and bool((got[7] == 0).all())
    raw = scatter_max(feats.cuda(), ids.to(idt).cuda(), m, fix_empty=False)
    seen = torch.zeros(m, dtype=torch.bool).index_fill_(0, ids[ids >= 0], True)
    assert not bool(seen[7]) and bool(torch.isneginf(raw[~seen.cuda()].float()).all())
    assert torch.equal(raw.float().cpu()[seen], ref[seen])
    assert scatter_max(feats[:0].cuda(), ids[:0].to(idt).cuda(), 5).abs().sum().item() == 0


@pytest.mark.parametrize('c,idt,m', [(64, torch.int64, 8), (128, torch.int32, 3), (8, torch.int64, 64)])
def test_scatter_mean_bf16_rows_into_few(c, idt, m):
    """SE layer pooling in bf16 inference (se_layer.py:17-23): per-frame mean of bf16 point features, fp32 out."""
    from openseg3d_b200.ops import scatter_mean
    from oracle import oracle
    torch.manual_seed(c + m)
    n = 70000
    feats = torch.randn(n, c).bfloat16()
    ids = torch.sort(torch.randint(0, m, (n,))).values            # frames arrive in order
    ids[::97] = -1
    ids[-1] = m - 1
    ref = oracle.scatter_reduce(feats.float(), ids, 'mean')
    got = scatter_mean(feats.cuda(), ids.to(idt).cuda(), m)
    assert got.dtype == torch.float32
    torch.testing.assert_close(got.cpu(), ref, rtol=1e-4, atol=1e-6)
    shuffled = torch.randperm(n)                                  # any order is correct, sorted runs are just faster
    got2 = scatter_mean(feats[shuffled].cuda(), ids[shuffled].to(idt).cuda(), m)
    torch.testing.assert_close(got2.cpu(), ref, rtol=1e-4, atol=1e-6)


def test_vfe_and_pooling_api():
    from openseg3d_b200.models import VFE
    from openseg3d_b200.ops import voxel_avg_pooling, voxel_max_pooling
    from oracle import oracle
    torch.manual_seed(0)
    feats = torch.randn(2000, 16)
    ids = torch.randint(-1, 300, (2000,))
    ids[:300] = torch.arange(300)
    vfe = VFE(16, 'max')
    assert vfe.voxel_feature_channel == 16
    assert torch.equal(vfe(feats.cuda(), ids.cuda()).cpu(), oracle.scatter_reduce(feats, ids, 'max'))
    assert torch.equal(voxel_max_pooling(feats.cuda(), ids.cuda()).cpu(), oracle.scatter_reduce(feats, ids, 'max'))
    counts = torch.bincount(ids[ids >= 0], minlength=300).int()
    avg = voxel_avg_pooling(feats.cuda(), ids.int().cuda(), counts.cuda()).cpu()
    torch.testing.assert_close(avg, oracle.voxel_avg_pooling(feats, ids, counts), rtol=1e-4, atol=1e-6)


def test_scatter_backward_matches_autograd():
    from openseg3d_b200.ops import scatter_max, scatter_mean, voxel_to_point
    torch.manual_seed(1)
    feats = torch.randn(3000, 8)
    ids = torch.randint(-1, 400, (3000,))
    ids[:400] = torch.arange(400)
    for fn, red in ((scatter_max, 'amax'), (scatter_mean, 'mean')):
        a = feats.clone().cuda().requires_grad_(True)
        out = fn(a, ids.cuda(), 400)
        w = torch.randn_like(out)
        (out * w).sum().backward()
        b = feats.clone().requires_grad_(True)
        mask = ids >= 0
        ref = torch.zeros(400, 8).scatter_reduce(0, ids[mask][:, None].expand(-1, 8), b[mask], red, include_self=False)
        (ref * w.cpu()).sum().backward()
        torch.testing.assert_close(a.grad.cpu(), b.grad, rtol=1e-4, atol=1e-6)
    v = torch.randn(400, 8, device='cuda', requires_grad=True)
    o = voxel_to_point(v, ids.cuda())
    o.sum().backward()
    cnt = torch.bincount(ids[ids >= 0], minlength=400).float()
    torch.testing.assert_close(v.grad.cpu(), cnt[:, None].expand(-1, 8))


def test_voxel_to_point_matches_reference_golden(golden_dir):
    from openseg3d_b200.ops import voxel_to_point
    g = np.load(os.path.join(golden_dir, 'stage1_gather.npz'))
    out = voxel_to_point(torch.from_numpy(g['feats']).cuda(), torch.from_numpy(g['ids']).cuda())
    assert torch.equal(out.cpu(), torch.from_numpy(g['out']))
    bf = voxel_to_point(torch.from_numpy(g['feats']).cuda().bfloat16(), torch.from_numpy(g['ids']).cuda())
    assert torch.equal(bf.cpu(), torch.from_numpy(g['out']).bfloat16())


def test_scatter_max_backward_with_ties_routes_to_one_row():
    """torch_scatter semantics (VFE: vfe.py:24-25): among tied maxima ONE row receives the gradient (here the lowest
    point index), so the input gradient of every (voxel, channel) sums to the output gradient -- ties must not multiply it."""
    from openseg3d_b200.ops import scatter_max
    torch.manual_seed(2)
    n, m, c = 4000, 300, 8
    feats = torch.randn(n, c).bfloat16().float()                       # 8-bit mantissa: many collisions
    feats[0:3999:3] = feats[1:4000:3]                                  # and exact duplicates (1333 rows each)
    ids = torch.randint(0, m, (n,))
    a = feats.clone().cuda().requires_grad_(True)
    out = scatter_max(a, ids.cuda(), m)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    g = a.grad.cpu()
    # per (voxel, channel): exactly one non-zero entry, equal to w, sitting on the first row that attains the maximum
    tot = torch.zeros(m, c).index_add_(0, ids, g)
    seen = torch.zeros(m, dtype=torch.bool).index_fill_(0, ids, True)
    torch.testing.assert_close(tot[seen], w.cpu()[seen])
    ref_out = torch.full((m, c), -float('inf')).scatter_reduce(0, ids[:, None].expand(-1, c), feats, 'amax')
    is_max = feats == ref_out[ids]
    first = torch.full((m, c), n, dtype=torch.long).scatter_reduce(0, ids[:, None].expand(-1, c),
                                                                   torch.where(is_max, torch.arange(n)[:, None].expand(-1, c), n), 'amin')
    expect = torch.where(first[ids] == torch.arange(n)[:, None], w.cpu()[ids], torch.zeros(()))
    torch.testing.assert_close(g, expect)
