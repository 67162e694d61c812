"""GPU parity, stages 2-3: kernel maps bit-exact vs the oracle (same canonical order, so the tables compare
element-wise), sparse convolution fp32 within rel 1e-4, modules (SubM / strided / inverse, fused BN+ReLU+residual)."""
import numpy as np
import pytest
import torch

from openseg3d_b200 import synthetic
from tests.golden_cfg import CFG

pytestmark = pytest.mark.gpu


def _sites(rng, batch, shape, n):
    out = []
    for b in range(batch):
        c = np.unique(rng.integers(0, shape, (n, 3)), axis=0)
        c = c[rng.permutation(len(c))]
        out.append(np.pad(c, ((0, 0), (1, 0)), constant_values=b))
    return np.concatenate(out).astype(np.int32)


def _tensor(idx, shape, batch, feats=None):
    from openseg3d_b200 import spconv
    f = feats if feats is not None else torch.zeros(idx.shape[0], 1)
    return spconv.SparseConvTensor(f.cuda(), torch.from_numpy(idx).cuda(), list(shape), batch)


@pytest.mark.parametrize('shape,n', [((8, 12, 10), 200), ((5, 41, 37), 3000), ((64, 90, 90), 20000)])
def test_kernel_maps_bit_exact_small(shape, n):
    from openseg3d_b200 import spconv
    from oracle import oracle
    rng = np.random.default_rng(n)
    idx = _sites(rng, 3, shape, n)
    x = _tensor(idx, shape, 3)
    rb = spconv.build_subm_rulebook(x)
    nbr, pairs = oracle.subm_map(idx, shape)
    assert np.array_equal(rb.nbr.cpu().numpy(), nbr) and rb.num_pairs == pairs
    sb = spconv.build_strided_rulebook(x)
    o_idx, o_shape, fwd, inv, p2 = oracle.strided_map(idx, shape)
    assert sb.out_shape == o_shape.tolist()
    assert np.array_equal(sb.out_indices.cpu().numpy(), o_idx)
    assert np.array_equal(sb.fwd_nbr.cpu().numpy(), fwd)
    assert np.array_equal(sb.inv_nbr.cpu().numpy(), inv)
    assert sb.num_pairs == p2
    # canonical (k, in, out) triples agree too -- the form the north star states the bit-exactness in
    assert np.array_equal(oracle.pairs_from_table(sb.fwd_nbr.cpu().numpy()), oracle.pairs_from_table(fwd))


def test_kernel_maps_full_frame_all_levels():
    """Waymo-shape frame(s): 3 strided levels + 4 submanifold maps, every table equal to the oracle's."""
    from openseg3d_b200 import spconv
    from openseg3d_b200.core import voxelize_batch
    from oracle import oracle
    cfg = CFG['cart']
    pts, _ = synthetic.make_batch([0, 1], 1, False)
    coors, _ = voxelize_batch(torch.from_numpy(pts).cuda(), cfg['voxel_size'], cfg['pc_range'])
    idx = coors.cpu().numpy()
    shape = [64, 1440, 1440]
    x = spconv.SparseConvTensor(torch.zeros(idx.shape[0], 1).cuda(), coors, shape, 2)
    for level in range(4):
        rb = spconv.build_subm_rulebook(x)
        nbr, pairs = oracle.subm_map(idx, shape)
        assert np.array_equal(rb.nbr.cpu().numpy(), nbr) and rb.num_pairs == pairs
        if level == 3:
            break
        sb = spconv.build_strided_rulebook(x)
        o_idx, o_shape, fwd, inv, p2 = oracle.strided_map(idx, shape)
        assert np.array_equal(sb.out_indices.cpu().numpy(), o_idx)
        assert np.array_equal(sb.fwd_nbr.cpu().numpy(), fwd) and np.array_equal(sb.inv_nbr.cpu().numpy(), inv)
        idx, shape = o_idx, o_shape.tolist()
        x = spconv.SparseConvTensor(torch.zeros(idx.shape[0], 1).cuda(), sb.out_indices, shape, 2)


def test_kernel_map_edge_cases():
    from openseg3d_b200 import spconv
    # single voxel in a corner; full dense block (every neighbour present)
    one = np.array([[0, 0, 0, 0]], np.int32)
    rb = spconv.build_subm_rulebook(_tensor(one, (4, 4, 4), 1))
    t = rb.nbr.cpu().numpy()
    assert t[0, 13] == 0 and (np.delete(t[0], 13) == -1).all()
    zz, yy, xx = np.meshgrid(np.arange(4), np.arange(5), np.arange(6), indexing='ij')
    dense = np.stack([np.zeros(120, int), zz.ravel(), yy.ravel(), xx.ravel()], 1).astype(np.int32)
    rb = spconv.build_subm_rulebook(_tensor(dense, (4, 5, 6), 1))
    inner = (dense[:, 1] % 3 == 1) & (dense[:, 2] > 0) & (dense[:, 2] < 4) & (dense[:, 3] > 0) & (dense[:, 3] < 5) & (dense[:, 1] > 0) & (dense[:, 1] < 3)
    assert (rb.nbr.cpu().numpy()[inner] >= 0).all()
    sb = spconv.build_strided_rulebook(_tensor(dense, (4, 5, 6), 1))
    assert sb.out_shape == [2, 3, 3] and sb.out_indices.shape[0] == 18


@pytest.mark.parametrize('cin,cout', [(64, 48), (48, 96), (6, 48), (96, 32), (20, 70)])
def test_spconv_f32_matches_oracle(cin, cout):
    from openseg3d_b200 import spconv
    from oracle import oracle
    rng = np.random.default_rng(cin * 100 + cout)
    torch.manual_seed(cin + cout)
    shape = (16, 60, 60)
    idx = _sites(rng, 2, shape, 4000)
    feats = torch.randn(idx.shape[0], cin)
    x = _tensor(idx, shape, 2, feats)
    conv = spconv.SubMConv3d(cin, cout, 3, padding=1, bias=True, indice_key='k').cuda()
    with torch.no_grad():
        conv.bias.normal_()
        y = conv(x)
    nbr, _ = oracle.subm_map(idx, shape)
    ref = oracle.sparse_conv(feats.double(), nbr, conv.weight.detach().cpu().double(), conv.bias.detach().cpu().double())
    torch.testing.assert_close(y.features.cpu().double(), ref, rtol=1e-4, atol=1e-5)
    assert y.indices is x.indices and 'k' in x.indice_dict


def test_strided_inverse_and_fused_epilogue_f32():
    from openseg3d_b200 import spconv
    from openseg3d_b200.models.backbones import ConvModule, SparseBasicBlock
    from functools import partial
    import torch.nn as nn
    from oracle import oracle
    rng = np.random.default_rng(7)
    torch.manual_seed(7)
    shape = (16, 50, 50)
    idx = _sites(rng, 2, shape, 5000)
    feats = torch.randn(idx.shape[0], 16)
    norm = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
    down = ConvModule(16, 32, 3, stride=2, padding=1, conv_type='spconv', norm_fn=norm, act_fn=nn.ReLU(True), indice_key='spconv2').cuda().eval()
    block = SparseBasicBlock(32, 32, norm_fn=norm, act_fn=nn.ReLU(True), indice_key='subm2').cuda().eval()
    up = ConvModule(32, 8, 3, conv_type='inverseconv', norm_fn=norm, act_fn=nn.ReLU(True), indice_key='spconv2').cuda().eval()
    with torch.no_grad():
        for mod in (down[1], block.bn1, block.bn2, up[1]):
            mod.running_mean.normal_(); mod.running_var.uniform_(0.5, 2.0); mod.weight.normal_(1, 0.2); mod.bias.normal_()
        x = _tensor(idx, shape, 2, feats)
        y = down(x)
        z = block(y)
        w = up(z)
    assert w.indices is x.indices and w.spatial_shape == list(shape)
    # oracle
    o_idx, o_shape, fwd, inv, _ = oracle.strided_map(idx, shape)
    sd = {k: v.detach().cpu() for k, v in down.state_dict().items()}
    ref_y = oracle._conv_bn_relu(feats, fwd, {('d.' + k): v for k, v in sd.items()}, 'd.')
    torch.testing.assert_close(y.features.cpu(), ref_y, rtol=1e-4, atol=1e-5)
    nb2, _ = oracle.subm_map(o_idx, o_shape)
    ref_z = oracle._basic_block(ref_y, nb2, {('b.' + k): v.detach().cpu() for k, v in block.state_dict().items()}, 'b.')
    torch.testing.assert_close(z.features.cpu(), ref_z, rtol=1e-4, atol=1e-5)
    ref_w = oracle._conv_bn_relu(ref_z, inv, {('u.' + k): v.detach().cpu() for k, v in up.state_dict().items()}, 'u.')
    torch.testing.assert_close(w.features.cpu(), ref_w, rtol=1e-4, atol=1e-5)
    # train-mode (unfused) module path gives the same numbers when BN uses running stats... checked via eval of the
    # unfused branch: call the conv without epilogue and apply BN in torch
    with torch.no_grad():
        raw = down[0](x)
        unf = torch.relu(down[1](raw.features))
    torch.testing.assert_close(unf, y.features, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('cin,cout,epi', [(64, 48, 'plain'), (48, 96, 'bn_relu'), (96, 96, 'residual'), (192, 192, 'bn_relu'),
                                          (384, 384, 'residual'), (768, 384, 'bn_relu'), (6, 48, 'bn_relu'),
                                          (48, 32, 'plain'), (24, 16, 'plain')])
def test_spconv_bf16_tensor_core_matches_oracle(cin, cout, epi):
    """tcgen05 path: bf16 operands, fp32 accumulate in TMEM, fused scale/shift/residual/ReLU epilogue.  The oracle gets
    the same bf16-rounded operands, so the only differences are fp32 accumulation order and the final bf16 rounding
    (tolerance: rel 2e-2 of the output scale, as the north star states for bf16)."""
    from openseg3d_b200 import spconv
    from openseg3d_b200.spconv.modules import sparse_conv_forward, _PackedWeights
    from oracle import oracle
    rng = np.random.default_rng(cin * 1000 + cout)
    torch.manual_seed(cin + 7 * cout)
    shape = (12, 40, 40)
    idx = _sites(rng, 2, shape, 3000)                      # ~2 x 2700 rows: 43 CTAs incl. a ragged last tile
    m = idx.shape[0]
    feats = torch.randn(m, cin).bfloat16()
    w = (torch.randn(cout, 3, 3, 3, cin) / np.sqrt(27 * cin) * 3).bfloat16().float()
    bias = torch.randn(cout)
    x = _tensor(idx, shape, 2, feats)
    rb = spconv.build_subm_rulebook(x)
    scale = shift = residual = None
    relu = False
    if epi != 'plain':
        scale, shift, relu = torch.rand(cout) + 0.5, torch.randn(cout), True
    if epi == 'residual':
        residual = torch.randn(m, cout).bfloat16()
    cuda = lambda t: None if t is None else t.cuda()
    y = sparse_conv_forward(feats.cuda(), rb.nbr, w.cuda(), cuda(bias) if epi == 'plain' else None, _PackedWeights(),
                            cuda(scale), cuda(shift), cuda(residual), relu)
    assert y.dtype == torch.bfloat16 and y.shape == (m, cout)
    nbr, _ = oracle.subm_map(idx, shape)
    ref = oracle.sparse_conv(feats.double(), nbr, w.double(), bias.double() if epi == 'plain' else None)
    if scale is not None:
        ref = ref * scale.double() + shift.double()
    if residual is not None:
        ref = ref + residual.double()
    if relu:
        ref = ref.clamp(min=0)
    err = (y.float().cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-2, err
    # and much tighter on average (bf16 output rounding only)
    mean_err = (y.float().cpu().double() - ref).abs().mean().item() / ref.abs().mean().item()
    assert mean_err < 4e-3, mean_err


def test_spconv_bf16_strided_and_inverse_modules():
    from functools import partial
    import torch.nn as nn
    from openseg3d_b200.models.backbones import ConvModule
    from oracle import oracle
    rng = np.random.default_rng(3)
    torch.manual_seed(3)
    shape = (16, 50, 50)
    idx = _sites(rng, 2, shape, 5000)
    feats = torch.randn(idx.shape[0], 48).bfloat16()
    norm = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
    down = ConvModule(48, 96, 3, stride=2, padding=1, conv_type='spconv', norm_fn=norm, act_fn=nn.ReLU(True), indice_key='spconv2').cuda().eval()
    up = ConvModule(96, 48, 3, conv_type='inverseconv', norm_fn=norm, act_fn=nn.ReLU(True), indice_key='spconv2').cuda().eval()
    with torch.no_grad():
        for mod in (down, up):
            mod[0].weight.copy_(mod[0].weight.bfloat16().float())
            mod[1].running_mean.normal_(); mod[1].running_var.uniform_(0.5, 2.0)
        x = _tensor(idx, shape, 2, feats)
        y = down(x)
        z = up(y)
    o_idx, o_shape, fwd, inv, _ = oracle.strided_map(idx, shape)
    ref_y = oracle._conv_bn_relu(feats.float(), fwd, {('d.' + k): v.detach().cpu() for k, v in down.state_dict().items()}, 'd.')
    e1 = (y.features.float().cpu() - ref_y).abs().max().item() / ref_y.abs().max().item()
    ref_z = oracle._conv_bn_relu(y.features.float().cpu(), inv, {('u.' + k): v.detach().cpu() for k, v in up.state_dict().items()}, 'u.')
    e2 = (z.features.float().cpu() - ref_z).abs().max().item() / ref_z.abs().max().item()
    assert e1 < 2e-2 and e2 < 2e-2, (e1, e2)
    assert z.indices is x.indices


def test_full_size_conv_is_independent_of_tile_order(monkeypatch):
    """BASELINE size (8 synthetic frames, ~0.93 M voxels at level 1 / ~1.0 M at level 2): the mask-sorted tile order
    (os3d_kernel_map_order) only changes WHICH tile computes a row, never the arithmetic of the row, so the bf16
    tensor-core conv must give bit-identical outputs in storage order and in sorted order; perm must be a permutation."""
    from openseg3d_b200 import spconv, synthetic
    from openseg3d_b200.core import voxelize_batch
    from openseg3d_b200.spconv.modules import sparse_conv_forward, kernel_map_tiles, _PackedWeights
    frames = 8
    pts, _ = synthetic.make_batch(list(range(frames)), 1, False)
    coors, _ = voxelize_batch(torch.from_numpy(pts).cuda(), [0.1, 0.1, 0.1], [-72, -72, -2, 72, 72, 4.4])
    x = spconv.SparseConvTensor(torch.zeros(coors.shape[0], 1, device='cuda'), coors, [64, 1440, 1440], frames)
    rb = spconv.build_strided_rulebook(x)
    torch.manual_seed(0)
    for nbr, m_in, cin, cout in ((spconv.build_subm_rulebook(x).nbr, coors.shape[0], 48, 48),
                                 (rb.fwd_nbr, coors.shape[0], 48, 96), (rb.inv_nbr, rb.out_indices.shape[0], 96, 48)):
        feats = torch.randn(m_in, cin, device='cuda').bfloat16()
        w = torch.randn(cout, 3, 3, 3, cin, device='cuda') * 0.05
        outs = []
        for order in ('0', '1'):
            monkeypatch.setenv('OS3D_MAP_ORDER', order)
            if hasattr(nbr, '_os3d_tiles'):
                del nbr._os3d_tiles
            outs.append(sparse_conv_forward(feats, nbr, w, None, _PackedWeights(), None, None, None, True))
            _, tile_mask, perm = kernel_map_tiles(nbr)
            if order == '1':
                assert bool((torch.sort(perm.long()).values == torch.arange(nbr.shape[0], device='cuda')).all())
                sorted_offsets = tile_mask
            else:
                assert perm is None
                storage_offsets = tile_mask
        assert torch.equal(outs[0], outs[1])

        def offsets_per_tile(t):
            t = t.long() & 0x7ffffff
            return sum(((t >> b) & 1) for b in range(27)).float().mean().item()
        assert offsets_per_tile(sorted_offsets) < offsets_per_tile(storage_offsets)


@pytest.mark.parametrize('cout', [48, 96, 192])
def test_spconv_bf16_pair_sum_epilogue_and_pitched_output_match_oracle(cout):
    """UpBlock's bottleneck at inference (pointtransformer.py:105-110): conv over cat([x_bottom, x_trans]) (2C channels) +
    BatchNorm + ReLU, then `x_m + channel_reduction(cat)` -- flags = 3: the residual has 2*cout channels and
    residual[:, 2c] + residual[:, 2c+1] is added AFTER the ReLU -- written through os3d_spconv_fwd_bf16_ld into the LEFT
    half of a double-width buffer (output row pitch 2*cout), against the oracle's restatement of the same lines."""
    from openseg3d_b200 import spconv
    from openseg3d_b200.spconv.modules import sparse_conv_forward, _PackedWeights
    from oracle import oracle
    rng = np.random.default_rng(cout)
    torch.manual_seed(cout)
    shape = (12, 40, 40)
    idx = _sites(rng, 2, shape, 3000)
    m, cin = idx.shape[0], 2 * cout
    cat = torch.randn(m, cin).bfloat16()
    w = (torch.randn(cout, 3, 3, 3, cin) / np.sqrt(27 * cin) * 3).bfloat16().float()
    scale, shift = torch.rand(cout) + 0.5, torch.randn(cout)
    rb = spconv.build_subm_rulebook(_tensor(idx, shape, 2, cat))
    wide = torch.full((m, 2 * cout), 7.0, dtype=torch.bfloat16, device='cuda')         # right half must stay untouched
    y = sparse_conv_forward(cat.cuda(), rb.nbr, w.cuda(), None, _PackedWeights(), scale.cuda(), shift.cuda(), cat.cuda(), 3,
                            out=wide[:, :cout])
    assert y.data_ptr() == wide.data_ptr() and y.stride(0) == 2 * cout
    assert bool((wide[:, cout:] == 7.0).all())
    nbr, _ = oracle.subm_map(idx, shape)
    x_m = (oracle.sparse_conv(cat.double(), nbr, w.double()) * scale.double() + shift.double()).clamp(min=0)
    ref = x_m + cat.double().view(m, cout, 2).sum(dim=2)                              # channel_reduction :89-103
    got = wide[:, :cout].float().cpu().double()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    mean_err = (got - ref).abs().mean().item() / ref.abs().mean().item()
    assert err < 2e-2 and mean_err < 4e-3, (err, mean_err)


def test_full_size_conv_matches_oracle():
    """BASELINE size, ONE full frame (116k voxels at level 1, 130k at level 2): the tensor-core conv on the level-1
    submanifold map, the strided map and its inverse against the oracle's sparse_conv (C restatement) on the same maps."""
    from openseg3d_b200 import spconv, synthetic
    from openseg3d_b200.core import voxelize_batch
    from openseg3d_b200.spconv.modules import sparse_conv_forward, _PackedWeights
    from oracle import oracle
    pts, _ = synthetic.make_batch([0], 1, False)
    coors, _ = voxelize_batch(torch.from_numpy(pts).cuda(), [0.1, 0.1, 0.1], [-72, -72, -2, 72, 72, 4.4])
    x = spconv.SparseConvTensor(torch.zeros(coors.shape[0], 1, device='cuda'), coors, [64, 1440, 1440], 1)
    rb = spconv.build_strided_rulebook(x)
    idx = coors.cpu().numpy()
    o_idx, o_shape, fwd, inv, _ = oracle.strided_map(idx, [64, 1440, 1440])
    sub, _ = oracle.subm_map(idx, [64, 1440, 1440])
    torch.manual_seed(0)
    for nbr, ref_nbr, m_in, cin, cout in ((spconv.build_subm_rulebook(x).nbr, sub, coors.shape[0], 48, 48),
                                          (rb.fwd_nbr, fwd, coors.shape[0], 48, 96),
                                          (rb.inv_nbr, inv, rb.out_indices.shape[0], 96, 48)):
        assert np.array_equal(nbr.cpu().numpy(), ref_nbr)
        feats = torch.randn(m_in, cin).bfloat16()
        w = (torch.randn(cout, 3, 3, 3, cin) * 0.05).bfloat16().float()
        y = sparse_conv_forward(feats.cuda(), nbr, w.cuda(), None, _PackedWeights(), None, None, None, False)
        ref = oracle.sparse_conv(feats.float(), ref_nbr, w)
        err = (y.float().cpu() - ref).abs().max().item() / ref.abs().max().item()
        assert err < 2e-2, (cin, cout, err)
