"""Generate the committed golden vectors by running the REFERENCE's own code.

Run in the build container only (needs /root/reference and numba):
    python tests/golden/make_golden.py
The GPU box has no /root/reference; tests read the .npz files written here.

What runs reference code unmodified:
  * seg3d/core/voxel/voxel_generator.py   (numba voxelizer)            -> voxelize_*.npz
  * seg3d/utils/pointops_utils.cart2polar  (restated inline: 3 numpy lines, pointops_utils.py:8-11,
    because importing that module pulls the CUDA-only seg3d.ops)
  * seg3d/ops/voxel_to_point/voxel_to_point.py                          -> stage1_gather.npz
  * seg3d/utils/swformer_utils.py, seg3d/models/layers/{point_transformer_layer,cosine_msa,drop}.py
    with seg3d.ops.get_inner_win_inds replaced by a stable within-group rank (the reference kernel is an
    atomicAdd race: ingroup_inds_cuda.cu:23; any in-window order is a valid output)   -> swformer_*.npz
"""
import hashlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))

from openseg3d_b200 import synthetic  # noqa: E402


def load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def stable_rank(group):
    g = group.cpu().numpy()
    order = np.argsort(g, kind='stable')
    gs = g[order]
    start = np.r_[0, np.nonzero(np.diff(gs))[0] + 1]
    seg_start = np.repeat(start, np.diff(np.r_[start, len(gs)]))
    rank = np.empty_like(g)
    rank[order] = np.arange(len(g)) - seg_start
    return torch.from_numpy(rank).to(group.device)


def install_stubs():
    for pkg in ['seg3d', 'seg3d.ops', 'seg3d.utils', 'seg3d.models', 'seg3d.models.layers']:
        m = types.ModuleType(pkg)
        m.__path__ = []
        sys.modules[pkg] = m
    sys.modules['seg3d.ops'].get_inner_win_inds = stable_rank


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


CART = dict(voxel_size=[0.1, 0.1, 0.1], pc_range=[-72, -72, -2, 72, 72, 4.4])
CYL = dict(voxel_size=[0.05, 0.012, 0.1], pc_range=[0, -3.1415926, -2, 75.2, 3.1415926, 5.2])


def gold_voxelize(vg_mod):
    out = {}
    cases = {
        # name: (cfg, seeds, sweeps, cylinder, beams, cols)
        'cart_small': (CART, [11, 12], 1, False, 16, 500),
        'cyl_small': (CYL, [13], 1, True, 16, 500),
        'multi_small': (CART, [14], 3, False, 16, 300),
    }
    for name, (cfg, seeds, sweeps, cyl, nb, nc) in cases.items():
        gen = vg_mod.VoxelGenerator(cfg['voxel_size'], cfg['pc_range'])
        pts, _ = synthetic.make_batch(seeds, sweeps, cyl, nb, nc)
        # add hard cases: points exactly on voxel/range boundaries and outside the range
        rng = np.random.default_rng(99)
        extra = pts[rng.choice(len(pts), 64)].copy()
        lo = np.asarray(cfg['pc_range'][:3], np.float32)
        vs = np.asarray(cfg['voxel_size'], np.float32)
        k = rng.integers(0, 60, (64, 3)).astype(np.float32)
        extra[:, 1:4] = lo + k * vs                      # lattice points
        extra[:8, 1] = np.float32(cfg['pc_range'][3])    # x == upper bound -> dropped
        extra[8:16, 3] = np.float32(-50.0)               # below range
        pts = np.concatenate([pts, extra[extra[:, 0].argsort(kind='stable')]], axis=0)
        pts = pts[pts[:, 0].argsort(kind='stable')]
        coors_all, ids_all, base = [], [], 0
        for b in range(len(seeds)):
            p = pts[pts[:, 0] == b][:, 1:]
            coors, ids = gen.generate(p)
            ids = ids.astype(np.int64)
            ids[ids != -1] += base                       # collate_batch, waymo_dataset.py:358-365
            base += coors.shape[0]
            coors_all.append(np.pad(coors, ((0, 0), (1, 0)), constant_values=b))
            ids_all.append(ids)
        np.savez_compressed(os.path.join(HERE, f'voxelize_{name}.npz'), points=pts,
                            voxel_size=np.asarray(cfg['voxel_size'], np.float32),
                            pc_range=np.asarray(cfg['pc_range'], np.float32), grid_size=gen.grid_size,
                            coors=np.concatenate(coors_all).astype(np.int32), point_voxel_ids=np.concatenate(ids_all))
        out[name] = base
    # full-size frames: checksums only
    sums = {}
    for name, cfg, sweeps, cyl in [('cart_full', CART, 1, False), ('cyl_full', CYL, 1, True), ('multi_full', CART, 3, False)]:
        gen = vg_mod.VoxelGenerator(cfg['voxel_size'], cfg['pc_range'])
        pts, _ = synthetic.make_batch([0], sweeps, cyl)
        coors, ids = gen.generate(pts[:, 1:])
        sums[name] = dict(n=int(pts.shape[0]), m=int(coors.shape[0]), points=sha(pts),
                          coors=sha(np.pad(coors, ((0, 0), (1, 0))).astype(np.int32)), ids=sha(ids.astype(np.int64)))
        print(name, sums[name]['n'], sums[name]['m'])
    import json
    with open(os.path.join(HERE, 'voxelize_full_checksums.json'), 'w') as f:
        json.dump(sums, f, indent=1)
    return out


class FakeSparse:
    def __init__(self, features, indices):
        self.features, self.indices = features, indices


def gold_swformer(ptl):
    torch.manual_seed(0)
    # a clustered voxel set so windows of several occupancies (levels) exist; two frames
    rng = np.random.default_rng(5)
    sparse_xyz = (120, 120, 16)
    window = (10, 10, 8)
    batching_info = {0: {'max_tokens': 8, 'batching_range': (0, 8)}, 1: {'max_tokens': 32, 'batching_range': (8, 32)},
                     2: {'max_tokens': 128, 'batching_range': (32, 128)},
                     3: {'max_tokens': 800, 'batching_range': (128, 100000)}}
    coords = []
    for b in range(2):
        dense = rng.integers(0, [16, 20, 20], (1400, 3))          # z, y, x  dense corner -> level 3 windows
        medium = rng.integers(0, [16, 40, 40], (500, 3)) + np.array([0, 40, 40])
        sparse = rng.integers(0, [16, 120, 120], (500, 3))
        c = np.unique(np.concatenate([dense, medium, sparse]), axis=0)
        c = c[rng.permutation(len(c))]
        coords.append(np.pad(c, ((0, 0), (1, 0)), constant_values=b))
    coords = np.concatenate(coords).astype(np.int32)
    C, heads, depth = 48, 8, 2
    feats = torch.randn(coords.shape[0], C)
    layer = ptl.SparseWindowPartitionLayer(batching_info, window, sparse_xyz)
    block = ptl.SWFormerBlock(C, heads, depth=depth, drop_path=0.0)
    with torch.no_grad():
        for n, p in block.named_parameters():                       # non-trivial LN / tau / biases
            if n.endswith('tau'):
                p.fill_(0.35)
            elif 'norm' in n or n.endswith('bias'):
                p.add_(0.1 * torch.randn_like(p))
    block.eval()
    with torch.no_grad():
        info = layer(FakeSparse(feats, torch.from_numpy(coords)))
        out = block(info, using_checkpoint=False)
        # one attention call in isolation too
        attn = block.layers[0].win_attn(feats, info['pos_dict_shift0'], info['flat2win_inds_shift0'], info['key_mask_shift0'])
    save = dict(coords=coords, feats=feats.numpy(), out=out.numpy(), attn0=attn.numpy(),
                sparse_xyz=np.asarray(sparse_xyz), window=np.asarray(window),
                levels=np.asarray([[k, v['max_tokens'], v['batching_range'][0], v['batching_range'][1]]
                                   for k, v in batching_info.items()]),
                depth=depth, heads=heads)
    for s in range(2):
        save[f'win_s{s}'] = info[f'batch_win_inds_shift{s}'].numpy()
        save[f'inwin_s{s}'] = info[f'coors_in_win_shift{s}'].numpy()
        save[f'lvl_s{s}'] = info[f'voxel_batching_level_shift{s}'].numpy()
        for bl, v in info[f'flat2win_inds_shift{s}'].items():
            if isinstance(bl, str):
                continue
            save[f'slot_s{s}_l{bl}'] = v[0].numpy()
            save[f'where_s{s}_l{bl}'] = v[1][0].numpy()
            if s == 1:                                   # keep the fixture small: padded pos-embed of one shift
                save[f'pos_s{s}_l{bl}'] = info[f'pos_dict_shift{s}'][bl].numpy()
            save[f'mask_s{s}_l{bl}'] = info[f'key_mask_shift{s}'][bl].numpy()
    for k, v in block.state_dict().items():
        save['sd.' + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, 'swformer_block.npz'), **save)
    print('swformer golden:', coords.shape[0], 'voxels; levels used:',
          [k for k in info['flat2win_inds_shift0'] if not isinstance(k, str)])


def gold_gather():
    v2p = load('ref_voxel_to_point', 'seg3d/ops/voxel_to_point/voxel_to_point.py')
    torch.manual_seed(1)
    feats = torch.randn(37, 32)
    ids = torch.randint(-1, 37, (200,))
    out = v2p.voxel_to_point(feats, ids)
    np.savez_compressed(os.path.join(HERE, 'stage1_gather.npz'), feats=feats.numpy(), ids=ids.numpy(), out=out.numpy())


if __name__ == '__main__':
    vg = load('ref_voxel_generator', 'seg3d/core/voxel/voxel_generator.py')
    print(gold_voxelize(vg))
    gold_gather()
    install_stubs()
    load('seg3d.utils.swformer_utils', 'seg3d/utils/swformer_utils.py')
    load('seg3d.models.layers.drop', 'seg3d/models/layers/drop.py')
    load('seg3d.models.layers.cosine_msa', 'seg3d/models/layers/cosine_msa.py')
    ptl = load('seg3d.models.layers.point_transformer_layer', 'seg3d/models/layers/point_transformer_layer.py')
    gold_swformer(ptl)
