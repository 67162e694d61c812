"""Pins the oracle's stage-4 restatement (window partition, pos-embed, key masks, cosine window attention,
SWFormer block) against the reference's own modules (tests/golden/swformer_block.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle


@pytest.fixture(scope='module')
def g(golden_dir):
    return np.load(os.path.join(golden_dir, 'swformer_block.npz'))


def _binfo(g):
    return {int(r[0]): {'max_tokens': int(r[1]), 'batching_range': (int(r[2]), int(r[3]))} for r in g['levels']}


def test_partition_integer_outputs_bit_exact(g):
    binfo = _binfo(g)
    info = oracle.partition(g['coords'], g['sparse_xyz'].tolist(), g['window'].tolist(), binfo, g['feats'].shape[1])
    for s in range(2):
        assert np.array_equal(info[s]['win'], g[f'win_s{s}'])
        assert np.array_equal(info[s]['in_win'], g[f'inwin_s{s}'])
        assert np.array_equal(info[s]['lvl'], g[f'lvl_s{s}'])
        levels = [bl for bl in binfo if f'slot_s{s}_l{bl}' in g]
        assert sorted(info[s]['inds']) == sorted(levels)
        for bl in levels:
            assert np.array_equal(info[s]['inds'][bl][0], g[f'slot_s{s}_l{bl}'])
            assert np.array_equal(info[s]['inds'][bl][1], g[f'where_s{s}_l{bl}'])
            assert np.array_equal(info[s]['mask'][bl].numpy(), g[f'mask_s{s}_l{bl}'])
    for bl in info[1]['pos']:
        np.testing.assert_allclose(info[1]['pos'][bl].numpy(), g[f'pos_s1_l{bl}'], rtol=0, atol=1e-6)


def test_attention_and_block_match_reference(g):
    binfo = _binfo(g)
    x = torch.from_numpy(g['feats'])
    info = oracle.partition(g['coords'], g['sparse_xyz'].tolist(), g['window'].tolist(), binfo, x.shape[1])
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('sd.')}
    heads, depth = int(g['heads']), int(g['depth'])
    # one attention call
    p = {k[len('layers.0.win_attn.self_attn.'):]: v for k, v in sd.items() if k.startswith('layers.0.win_attn.self_attn.')}
    x3 = oracle.flat2window(x, info[0]['inds'], binfo)
    out3 = {bl: oracle.cosine_attention(x3[bl], info[0]['pos'][bl], info[0]['mask'][bl], p, heads) for bl in x3}
    attn = oracle.window2flat(out3, info[0]['inds'], x.shape[0])
    np.testing.assert_allclose(attn.numpy(), g['attn0'], rtol=1e-4, atol=2e-5)
    out = oracle.swformer_block(x, info, binfo, sd, depth, heads)
    np.testing.assert_allclose(out.numpy(), g['out'], rtol=1e-4, atol=5e-5)
