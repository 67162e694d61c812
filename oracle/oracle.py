"""CPU oracle for the OpenSeg3D voxel-backbone hot path.

TEST INFRASTRUCTURE ONLY -- the checker, never the product.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / ``--impl reference`` legs may import this module; nothing under
openseg3d_b200/ does.  Integer stages call the serial C restatement in os3d_oracle.c through ctypes;
floating-point stages are plain torch-CPU / numpy restatements of the reference's algorithm (padded
windows and all), each function citing the reference file:line it follows (relative to
/root/reference).

Pinning status (DESIGN.md "Oracle"):
  * voxelize               -- PINNED against the reference's own numba code (tests/golden/voxelize_*.npz).
  * window partition, pos-embed, key masks, cosine window attention, SWFormer block
                           -- PINNED against the reference's own modules imported in the build container
                              (tests/golden/swformer_*.npz), with get_inner_win_inds replaced by the stable rank.
  * voxel_to_point         -- PINNED (reference op is pure torch; golden in tests/golden/stage1_*.npz).
  * voxel_avg_pooling      -- PINNED against the reference's own CPU function voxel_pooling_forward_cpu, compiled from
                              /root/reference into oracle/_ref/ by oracle/build_ref.py (tests/golden/voxel_avg_pooling.npz).
  * scatter mean/max (VFE) -- torch_scatter absent => restated from its documented semantics; PARITY UNPINNED
                              (the mean is cross-checked against the pinned avg pooling).
  * kernel maps, sparse conv -- spconv absent => PARITY UNPINNED against spconv; pinned against torch dense
                              conv3d / conv_transpose3d instead (tests/test_oracle_spconv.py).
"""
import ctypes
import math
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile the C oracle (gcc, a second or two)."""
    subprocess.check_call(['make', '-s', '-C', _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'libos3d_oracle.so')
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, 'os3d_oracle.c')):
            build()
        L = ctypes.CDLL(path)
        L.os3d_oracle_voxelize.restype = ctypes.c_int64
        L.os3d_oracle_subm_map.restype = ctypes.c_int64
        L.os3d_oracle_strided_map.restype = ctypes.c_int64
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ------------------------------------------------------------------------------------------------
# Stage 1
# ------------------------------------------------------------------------------------------------
def grid_size(voxel_size, pc_range):
    """voxel_generator.py:15-18."""
    r = np.asarray(pc_range, np.float32)
    v = np.asarray(voxel_size, np.float32)
    return np.round((r[3:] - r[:3]) / v).astype(np.int64)


def voxelize(points, voxel_size, pc_range, has_batch=True):
    """voxel_generator.py:55-153 per frame + collate_batch id offsets (waymo_dataset.py:347-365).

    points: float32 [N, (1+)D].  Returns (coors int32 [M,4] (b,z,y,x), point_voxel_ids int64 [N])."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    n, stride = pts.shape
    rng = np.ascontiguousarray(pc_range, dtype=np.float32)
    vs = np.ascontiguousarray(voxel_size, dtype=np.float32)
    coors = np.empty((max(n, 1), 4), np.int32)
    pvid = np.empty((n,), np.int64)
    grid = np.zeros(3, np.int32)
    m = lib().os3d_oracle_voxelize(_p(pts), ctypes.c_int64(n), ctypes.c_int(stride), ctypes.c_int(int(has_batch)),
                                   _p(rng), _p(vs), _p(coors), _p(pvid), _p(grid))
    assert m >= 0
    return coors[:m].copy(), pvid


def scatter_reduce(features, index, reduce):
    """VFE.forward (seg3d/models/voxel_encoders/vfe.py:16-27): mask ids != -1 then
    torch_scatter.scatter(src, index, dim=0, reduce): out has index.max()+1 rows; 'max' leaves 0 in rows
    nothing scattered to; 'mean' = sum / max(count, 1)."""
    mask = index != -1
    src, idx = features[mask], index[mask].long()
    m = int(idx.max()) + 1 if idx.numel() else 0
    c = src.shape[1]
    if reduce == 'max':
        out = torch.zeros(m, c, dtype=src.dtype)
        out.scatter_reduce_(0, idx[:, None].expand(-1, c), src, 'amax', include_self=False)
        return out
    if reduce == 'mean':
        out = torch.zeros(m, c, dtype=src.dtype).index_add_(0, idx, src)
        cnt = torch.zeros(m, dtype=src.dtype).index_add_(0, idx, torch.ones_like(idx, dtype=src.dtype))
        return out / cnt.clamp(min=1)[:, None]
    if reduce == 'sum':
        return torch.zeros(m, c, dtype=src.dtype).index_add_(0, idx, src)
    raise ValueError(reduce)


def voxel_avg_pooling(feats, coords, counts):
    """voxel_pooling.cpp:5-23: out[pos] += in[i] / counts[pos], ids outside [0, M) skipped."""
    m = counts.shape[0]
    ok = (coords >= 0) & (coords < m)
    idx = coords[ok].long()
    return torch.zeros(m, feats.shape[1], dtype=feats.dtype).index_add_(
        0, idx, feats[ok] / counts[idx].to(feats.dtype)[:, None])


def voxel_to_point(feats, coords):
    """seg3d/ops/voxel_to_point/voxel_to_point.py:5-17."""
    out = torch.zeros(coords.shape[0], feats.shape[-1], dtype=feats.dtype)
    valid = torch.nonzero(coords != -1).view(-1)
    out[valid] = feats[coords[valid].long()]
    return out


# ------------------------------------------------------------------------------------------------
# Stages 2-3 (spconv semantics)
# ------------------------------------------------------------------------------------------------
def subm_map(indices, spatial_shape):
    """[M,27] neighbour table for SubMConv3d(k3, p1).  Returns (nbr int32, n_pairs)."""
    idx = np.ascontiguousarray(indices, np.int32)
    shp = np.ascontiguousarray(spatial_shape, np.int32)
    nbr = np.empty((idx.shape[0], 27), np.int32)
    pairs = lib().os3d_oracle_subm_map(_p(idx), ctypes.c_int64(idx.shape[0]), _p(shp), _p(nbr))
    return nbr, int(pairs)


def strided_map(indices, spatial_shape):
    """SparseConv3d(k3, s2, p1): (out_indices [M',4] ascending, out_shape, fwd_nbr [M',27], inv_nbr [M,27], n_pairs)."""
    idx = np.ascontiguousarray(indices, np.int32)
    shp = np.ascontiguousarray(spatial_shape, np.int32)
    m = idx.shape[0]
    out_shape = np.zeros(3, np.int32)
    out_idx = np.empty((8 * m + 1, 4), np.int32)
    fwd = np.empty((8 * m + 1, 27), np.int32)
    inv = np.empty((m, 27), np.int32)
    pairs = ctypes.c_int64(0)
    mo = lib().os3d_oracle_strided_map(_p(idx), ctypes.c_int64(m), _p(shp), _p(out_shape), _p(out_idx), _p(fwd),
                                       _p(inv), ctypes.byref(pairs))
    assert mo >= 0
    return out_idx[:mo].copy(), out_shape, fwd[:mo].copy(), inv, int(pairs.value)


def pairs_from_table(nbr):
    """Canonical kernel map: sorted (k, in_row, out_row) triples of an output-stationary table."""
    out_row, k = np.nonzero(nbr >= 0)
    trip = np.stack([k, nbr[out_row, k], out_row], axis=1).astype(np.int64)
    order = np.lexsort((trip[:, 2], trip[:, 1], trip[:, 0]))
    return trip[order]


def sparse_conv(features, nbr, weight, bias=None):
    """out[r] = bias + sum_k W[:, k, :] @ in[nbr[r, k]]   (gather - GEMM - scatter, one GEMM per offset).
    weight: [Cout, 3, 3, 3, Cin] (spconv 2.x layout)."""
    nbr_t = torch.as_tensor(nbr).long()
    cout = weight.shape[0]
    w = weight.reshape(cout, 27, -1)
    out = torch.zeros(nbr_t.shape[0], cout, dtype=features.dtype)
    for k in range(27):
        rows = torch.nonzero(nbr_t[:, k] >= 0).view(-1)
        if rows.numel() == 0:
            continue
        out.index_add_(0, rows, features[nbr_t[rows, k]] @ w[:, k, :].t())
    if bias is not None:
        out = out + bias
    return out


def sparse_conv_c(features, nbr, weight, bias=None):
    """Same, through the C loop with float64 accumulation (small cases only)."""
    f = np.ascontiguousarray(features, np.float32)
    nb = np.ascontiguousarray(nbr, np.int32)
    cout, cin = weight.shape[0], weight.shape[-1]
    w = np.ascontiguousarray(np.asarray(weight, np.float32).reshape(cout, 27, cin))
    out = np.empty((nb.shape[0], cout), np.float32)
    b = np.ascontiguousarray(bias, np.float32) if bias is not None else None
    lib().os3d_oracle_spconv(_p(f), _p(nb), ctypes.c_int64(nb.shape[0]), ctypes.c_int(cin), ctypes.c_int(cout), _p(w),
                             _p(b) if b is not None else None, _p(out))
    return out


def knn_query(nsample, xyz, new_xyz, offset, new_offset):
    """knn_query_cuda_kernel (seg3d/ops/knn_query/src/knn_query_cuda.cu:67-112) restated serially: per query a max-heap
    of the nsample best squared distances (root = worst; replaced only when strictly closer), keys of the query's batch
    segment in ascending order, heap-sorted ascending at the end.  float32 arithmetic, one rounding per operation.
    Returns (idx int32 [m, nsample], dist2 float32 [m, nsample]) -- squared distances, as the kernel writes them.
    parity unpinned: the reference kernel is GPU-only and cannot run in the build container."""
    xyz = np.ascontiguousarray(xyz, np.float32)
    new_xyz = np.ascontiguousarray(new_xyz, np.float32)
    offset, new_offset = np.asarray(offset, np.int64), np.asarray(new_offset, np.int64)
    m = new_xyz.shape[0]
    idx = np.zeros((m, nsample), np.int32)
    d2o = np.zeros((m, nsample), np.float32)

    def sift(dist, ind, k):                        # reheap, knn_query_cuda.cu:23-38
        root, child = 0, 1
        while child < k:
            if child + 1 < k and dist[child + 1] > dist[child]:
                child += 1
            if dist[root] > dist[child]:
                return
            dist[root], dist[child] = dist[child], dist[root]
            ind[root], ind[child] = ind[child], ind[root]
            root, child = child, 2 * child + 1

    for q in range(m):
        b = 0
        while q >= new_offset[b]:
            b += 1
        start, end = (0 if b == 0 else int(offset[b - 1])), int(offset[b])
        d = new_xyz[q][None, :] - xyz[start:end]                                   # float32
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]           # float32, no FMA
        dist = [np.float32(1e10)] * nsample
        ind = [start] * nsample
        for i in range(end - start):
            if d2[i] < dist[0]:
                dist[0], ind[0] = d2[i], start + i
                sift(dist, ind, nsample)
        for i in range(nsample - 1, 0, -1):        # heap_sort, :41-49
            dist[0], dist[i] = dist[i], dist[0]
            ind[0], ind[i] = ind[i], ind[0]
            sift(dist, ind, i)
        idx[q], d2o[q] = ind, dist
    return idx, d2o


# ------------------------------------------------------------------------------------------------
# Stage 4 -- window partition (swformer_utils.py, point_transformer_layer.py)
# ------------------------------------------------------------------------------------------------
def voxel_majority_labels(point_voxel_ids, point_labels, num_voxels, ignore_index=255):
    """WaymoDataset.prepare_voxel_labels, seg3d/datasets/waymo_dataset.py:213-246: a 256-bin counter per voxel over its
    points (voxel_id != -1), np.argmax per voxel (first maximum = lowest label), voxels without a point -> ignore_index."""
    ids = np.asarray(point_voxel_ids, np.int64)
    lab = np.asarray(point_labels, np.int64)
    counts = np.zeros((num_voxels, 256), dtype=np.int64)
    keep = ids != -1
    np.add.at(counts, (ids[keep], lab[keep]), 1)
    out = np.full(num_voxels, ignore_index, dtype=np.uint8)
    seen = counts.sum(axis=1) > 0
    out[seen] = counts[seen].argmax(axis=1).astype(np.uint8)
    return out


def voxel_centers(coords_zyx, scale, voxel_size, pc_range):
    """get_voxel_centers, seg3d/utils/pointops_utils.py:14-22 (float32 arithmetic in the same order)."""
    c = np.asarray(coords_zyx)[:, [2, 1, 0]].astype(np.float32)
    vs = np.asarray(voxel_size, np.float32) * np.float32(scale)
    return (c + np.float32(0.5)) * vs + np.asarray(pc_range[:3], np.float32)


def aux_voxel_labels(voxel_labels, voxel_coords, aux_voxel_coords, batch_size, voxel_size, pc_range, aux_scale=8.0):
    """tools/train.py:86-104: aux voxel -> label of the nearest level-1 voxel centre of the same frame (knn_query, k = 1)."""
    c1 = voxel_centers(np.asarray(voxel_coords)[:, 1:], 1.0, voxel_size, pc_range)
    c8 = voxel_centers(np.asarray(aux_voxel_coords)[:, 1:], aux_scale, voxel_size, pc_range)
    off = np.asarray([(np.asarray(voxel_coords)[:, 0] <= b).sum() for b in range(batch_size)], np.int32)
    aoff = np.asarray([(np.asarray(aux_voxel_coords)[:, 0] <= b).sum() for b in range(batch_size)], np.int32)
    idx, _ = knn_query(1, c1, c8, off, aoff)
    return np.asarray(voxel_labels)[idx.reshape(-1)]


def ingroup_rank(group):
    g = np.ascontiguousarray(group, np.int64)
    out = np.empty_like(g)
    if g.size:
        lib().os3d_oracle_ingroup_rank(_p(g), ctypes.c_int64(g.size), _p(out))
    return out


def window_coors(coors, sparse_shape, window_shape, do_shift):
    """get_window_coors, swformer_utils.py:109-154.  coors int64 [M,4] (b,z,y,x); shapes are (x,y,z)."""
    wx, wy, wz = window_shape
    sx, sy, sz = sparse_shape
    mx = int(np.ceil(sx / wx) + 1)
    my = int(np.ceil(sy / wy) + 1)
    mz = int(np.ceil(sz / wz) + 1)
    if do_shift:
        shx, shy, shz = wx // 2, wy // 2, wz // 2
    else:
        shx, shy, shz = wx, wy, wz
    if sz == wz:
        shz = 0
    cx, cy, cz = coors[:, 3] + shx, coors[:, 2] + shy, coors[:, 1] + shz
    win = coors[:, 0] * (mx * my * mz) + (cx // wx) * (my * mz) + (cy // wy) * mz + (cz // wz)
    in_win = np.stack([cz % wz, cy % wy, cx % wx], axis=-1)
    return win, in_win


def batching_levels(batch_win_inds, batching_info):
    """batching_single_shift, point_transformer_layer.py:71-87 -> (keep_mask, level per voxel)."""
    lvl = -np.ones_like(batch_win_inds)
    inner = ingroup_rank(batch_win_inds)
    per_voxel = np.bincount(batch_win_inds)[batch_win_inds]
    target = np.zeros_like(batch_win_inds)
    for bl in batching_info:
        lo, hi = batching_info[bl]['batching_range']
        m = (per_voxel >= lo) & (per_voxel < hi)
        target[m] = batching_info[bl]['max_tokens']
        lvl[m] = bl
    return inner < target, lvl


def flat2win_inds(batch_win_inds, lvl, batching_info):
    """get_flat2win_inds, swformer_utils.py:8-31 -> {level: (slot [n_l], positions [n_l])}."""
    out = {}
    for bl in batching_info:
        m = lvl == bl
        if not m.any():
            continue
        w = batch_win_inds[m]
        conti = np.searchsorted(np.unique(w), w)           # make_continuous_inds :158-171
        out[bl] = (conti * batching_info[bl]['max_tokens'] + ingroup_rank(conti), np.nonzero(m)[0])
    return out


def flat2window(feat, inds, batching_info):
    """flat2window, swformer_utils.py:34-64 -> {level: [R, T, C]}."""
    out = {}
    for bl, (slot, pos) in inds.items():
        t = batching_info[bl]['max_tokens']
        r = int(slot.max()) // t + 1
        buf = torch.zeros(r * t, feat.shape[-1], dtype=feat.dtype)
        buf[torch.as_tensor(slot)] = feat[torch.as_tensor(pos)]
        out[bl] = buf.reshape(r, t, -1)
    return out


def window2flat(feat3d, inds, n):
    """window2flat, swformer_utils.py:67-85."""
    first = next(iter(feat3d.values()))
    out = torch.zeros(n, first.shape[-1], dtype=first.dtype)
    for bl, f in feat3d.items():
        slot, pos = inds[bl]
        out[torch.as_tensor(pos)] = f.reshape(-1, f.shape[-1])[torch.as_tensor(slot)]
    return out


def pos_embed(coors_in_win, window_shape, feat_dim, temperature=1000):
    """get_pos_embed, point_transformer_layer.py:152-205 (3-D window branch, normalize_pos=False) -> [M, C] f32."""
    wx, wy, wz = window_shape
    c = torch.as_tensor(coors_in_win)
    z, y, x = c[:, 0] - wz / 2, c[:, 1] - wy / 2, c[:, 2] - wx / 2
    pos_length = feat_dim // 3
    inv_freq = torch.arange(pos_length, dtype=torch.float32)
    inv_freq = temperature ** (2 * torch.div(inv_freq, 2, rounding_mode='floor') / pos_length)
    embs = []
    for a in (x, y, z):
        e = a[:, None] / inv_freq[None, :]
        embs.append(torch.stack([e[:, ::2].sin(), e[:, 1::2].cos()], dim=-1).flatten(1))
    return torch.cat(embs, dim=-1).float()


def partition(coors, sparse_shape, window_shape, batching_info, feat_dim):
    """SparseWindowPartitionLayer.forward, point_transformer_layer.py:36-69, for both shifts.
    Asserts no token is dropped (the reference itself cannot continue if one is: Appendix C)."""
    coors = np.asarray(coors, np.int64)
    info = {}
    for s in range(2):
        win, in_win = window_coors(coors, sparse_shape, window_shape, s == 1)
        keep, lvl = batching_levels(win, batching_info)
        assert keep.all(), 'token drop not supported (reference desyncs here too)'
        inds = flat2win_inds(win, lvl, batching_info)
        pe = pos_embed(in_win, window_shape, feat_dim)
        info[s] = dict(win=win, in_win=in_win, lvl=lvl, inds=inds, pos_flat=pe,
                       pos=flat2window(pe, inds, batching_info),
                       mask={bl: v.squeeze(2) == 0 for bl, v in
                             flat2window(torch.ones(coors.shape[0], 1), inds, batching_info).items()})
    return info


def cosine_attention(x3, pos3, key_mask, p, nhead, tau_min=0.01):
    """CosineMultiheadAttention on one batching level.  cosine_msa.py:180-408 with q = k = x + pos, v = x
    (WindowAttention.forward, point_transformer_layer.py:233-258).  x3/pos3 [R,T,C]; key_mask [R,T] bool.
    p: in_proj_weight [3C,C], in_proj_bias [3C], tau [1,1,1], out_proj.weight/bias."""
    r, t, c = x3.shape
    d = c // nhead
    wq, wk, wv = p['in_proj_weight'].chunk(3)
    bq, bk, bv = p['in_proj_bias'].chunk(3)
    qk_in = x3 + pos3
    q = F.linear(qk_in, wq, bq).view(r, t, nhead, d).transpose(1, 2)     # [R,h,T,d]
    k = F.linear(qk_in, wk, bk).view(r, t, nhead, d).transpose(1, 2)
    v = F.linear(x3, wv, bv).view(r, t, nhead, d).transpose(1, 2)
    q = F.normalize(q, dim=-1)                                           # :152-153
    k = F.normalize(k, dim=-1)
    attn = q @ k.transpose(-2, -1) / p['tau'].reshape(()).clamp(min=tau_min)   # :154,162
    attn = attn + torch.zeros(r, 1, 1, t, dtype=x3.dtype).masked_fill_(key_mask[:, None, None, :], float('-inf'))
    attn = attn.softmax(dim=-1)                                          # :172
    o = (attn @ v).transpose(1, 2).reshape(r, t, c)                      # :176
    return F.linear(o, p['out_proj.weight'], p['out_proj.bias'])


def encoder_layer(x, shift_info, batching_info, p, nhead):
    """EncoderLayer.forward, point_transformer_layer.py:288-298 (eval: DropPath/dropout are identity)."""
    inds = shift_info['inds']
    x3 = flat2window(x, inds, batching_info)
    attn_p = {k[len('win_attn.self_attn.'):]: v for k, v in p.items() if k.startswith('win_attn.self_attn.')}
    out3 = {bl: cosine_attention(x3[bl], shift_info['pos'][bl].to(x.dtype), shift_info['mask'][bl], attn_p, nhead)
            for bl in x3}
    a = window2flat(out3, inds, x.shape[0])
    c = x.shape[1]
    x = x + F.layer_norm(a, (c,), p['norm1.weight'], p['norm1.bias'])
    h = F.linear(F.gelu(F.linear(x, p['mlp.fc1.weight'], p['mlp.fc1.bias'])), p['mlp.fc2.weight'], p['mlp.fc2.bias'])
    return x + F.layer_norm(h, (c,), p['norm2.weight'], p['norm2.bias'])


def swformer_block(x, info, batching_info, p, depth, nhead=8):
    """SWFormerBlock.forward, point_transformer_layer.py:314-339: first depth//2 layers shift 0, rest shift 1."""
    for i in range(depth):
        lp = {k[len(f'layers.{i}.'):]: v for k, v in p.items() if k.startswith(f'layers.{i}.')}
        x = encoder_layer(x, info[0 if i < depth // 2 else 1], batching_info, lp, nhead)
    return x


# ------------------------------------------------------------------------------------------------
# Whole forward (eval mode): Segformer.forward segformer.py:94-146 + PointTransformer.forward
# pointtransformer.py:181-219, functional over a reference-keyed state_dict.
# ------------------------------------------------------------------------------------------------
def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def _bn(x, p, eps):
    return F.batch_norm(x, p['running_mean'], p['running_var'], p['weight'], p['bias'], False, 0.0, eps)


def _mlp_seq(x, sd, prefix, layout):
    """layout: list of ('bn'|'lin'|'relu', index) following an nn.Sequential."""
    for kind, i in layout:
        if kind == 'bn':
            x = _bn(x, _sub(sd, f'{prefix}{i}.'), 1e-5)
        elif kind == 'lin':
            x = F.linear(x, sd[f'{prefix}{i}.weight'], sd.get(f'{prefix}{i}.bias'))
        else:
            x = F.relu(x)
    return x


def _conv_bn_relu(x, nbr, sd, prefix):
    """ConvModule (spconv_utils.py:13-32): conv(bias=False) + BN(eps 1e-3) + ReLU."""
    y = sparse_conv(x, nbr, sd[prefix + '0.weight'], sd.get(prefix + '0.bias'))
    return F.relu(_bn(y, _sub(sd, prefix + '1.'), 1e-3))


def _basic_block(x, nbr, sd, prefix):
    """SparseBasicBlock.forward, pointtransformer.py:47-66."""
    y = sparse_conv(x, nbr, sd[prefix + 'conv1.weight'], sd.get(prefix + 'conv1.bias'))
    y = F.relu(_bn(y, _sub(sd, prefix + 'bn1.'), 1e-3))
    y = sparse_conv(y, nbr, sd[prefix + 'conv2.weight'], sd.get(prefix + 'conv2.bias'))
    y = _bn(y, _sub(sd, prefix + 'bn2.'), 1e-3)
    return F.relu(y + x)


def _up_block(x_bottom, x_lateral, nbr_subm, nbr_out, sd, prefix):
    """UpBlock.forward, pointtransformer.py:105-113."""
    x_trans = _basic_block(x_lateral, nbr_subm, sd, prefix + 'transform.')
    x = torch.cat([x_bottom, x_trans], dim=1)
    x_m = _conv_bn_relu(x, nbr_subm, sd, prefix + 'bottleneck.')
    x = x.view(x.shape[0], x_m.shape[1], -1).sum(dim=2) + x_m            # channel_reduction :89-103
    return _conv_bn_relu(x, nbr_out, sd, prefix + 'out.')


def backbone_forward(sd, voxel_features, voxel_coords, sparse_shape_zyx, batching_info, window_shape, depths,
                     stats=None):
    """PointTransformer.forward.  sd keys are relative to 'point_transformer.'."""
    shapes = [np.asarray(sparse_shape_zyx, np.int32)]
    idx = [np.ascontiguousarray(voxel_coords, np.int32)]
    subm, fwd, inv = [], [], []
    for lvl in range(3):
        o_idx, o_shape, f, i, _ = strided_map(idx[lvl], shapes[lvl])
        idx.append(o_idx); shapes.append(o_shape); fwd.append(f); inv.append(i)
    pair_counts = {}
    for lvl in range(4):
        nb, npairs = subm_map(idx[lvl], shapes[lvl])
        subm.append(nb)
        pair_counts[f'subm{lvl + 1}'] = npairs
    if stats is not None:
        stats['voxels'] = [int(i.shape[0]) for i in idx]
        stats['pairs'] = pair_counts

    def stage(x, lvl):
        c = x.shape[1]
        sparse_xyz = (np.asarray(sparse_shape_zyx, np.float64)[::-1] / (2 ** lvl)).tolist()
        info = partition(np.asarray(idx[lvl], np.int64), sparse_xyz, window_shape, batching_info[lvl], c)
        return swformer_block(x, info, batching_info[lvl], _sub(sd, f'swformer_block{lvl + 1}.1.'), depths[lvl])

    x1 = stage(_conv_bn_relu(voxel_features, subm[0], sd, 'conv_input.'), 0)
    x2 = stage(_conv_bn_relu(x1, fwd[0], sd, 'conv_down1.'), 1)
    x3 = stage(_conv_bn_relu(x2, fwd[1], sd, 'conv_down2.'), 2)
    x4 = stage(_conv_bn_relu(x3, fwd[2], sd, 'conv_down3.'), 3)
    aux = F.linear(x4, sd['aux_voxel_classifier.0.weight'])
    y4 = _up_block(x4, x4, subm[3], inv[2], sd, 'up4.')
    y3 = _up_block(y4, x3, subm[2], inv[1], sd, 'up3.')
    y2 = _up_block(y3, x2, subm[1], inv[0], sd, 'up2.')
    y1 = _up_block(y2, x1, subm[0], subm[0], sd, 'up1.')
    if stats is not None:            # per-stage features for the parity error tables (tests/test_gpu_full_frame.py)
        stats['stages'] = dict(enc1=x1, enc2=x2, enc3=x3, enc4=x4, up4=y4, up3=y3, up2=y2, up1=y1)
    return dict(voxel_features=y1, voxel_out=F.linear(y1, sd['voxel_classifier.0.weight']), aux_voxel_out=aux,
                aux_voxel_coords=idx[3], voxel_coords=idx[0])


POINT_ENCODER = [('bn', 0), ('lin', 1), ('bn', 2), ('relu', 3), ('lin', 4), ('bn', 5), ('relu', 6), ('lin', 7), ('bn', 8),
                 ('relu', 9), ('lin', 10)]
FUSION_ENCODER = [('lin', 0), ('bn', 1), ('relu', 2), ('lin', 3), ('bn', 4), ('relu', 5), ('lin', 6), ('bn', 7), ('relu', 8)]
CLASSIFIER = [('lin', 0), ('bn', 1), ('relu', 2), ('lin', 4)]


def segformer_forward(sd, points, voxel_size, pc_range, batching_info, window_shape, depths, multi_sweeps=False,
                      stats=None, differentiable=False):
    """Raw collated points [sum N, 1+D] float32 (numpy) -> dict of torch tensors.  Eval mode.
    differentiable=True: ``sd`` holds CPU leaf tensors (any float dtype) that are used as they are, so torch autograd
    through this restatement is the gradient oracle of the backward tests."""
    if not differentiable:
        sd = {k: v.detach().float().cpu() for k, v in sd.items()}
    dt = next(v.dtype for v in sd.values() if v.dtype.is_floating_point)
    coors, pvid = voxelize(points, voxel_size, pc_range, has_batch=True)
    pv = torch.from_numpy(pvid)
    pts = torch.from_numpy(np.ascontiguousarray(points[:, 1:], np.float32)).to(dt)
    bidx = torch.from_numpy(np.ascontiguousarray(points[:, 0])).long()
    if multi_sweeps:
        cur = pts[:, 3] == 0
        cur_pts = pts[cur]
    else:
        cur_pts = pts
    ppf = _mlp_seq(cur_pts, sd, 'point_encoder.', POINT_ENCODER)
    vf = scatter_reduce(pts, pv, 'mean') if multi_sweeps else scatter_reduce(ppf, pv, 'max')
    grid = grid_size(voxel_size, pc_range)
    bb = backbone_forward(_sub(sd, 'point_transformer.'), vf, coors, grid[::-1], batching_info, window_shape, depths,
                          stats=stats)
    pvf = voxel_to_point(bb['voxel_features'], pv[cur] if multi_sweeps else pv)
    fused = _mlp_seq(torch.cat([ppf, pvf], dim=1), sd, 'fusion_encoder.', FUSION_ENCODER)
    pb = bidx[cur] if multi_sweeps else bidx
    se_in = scatter_reduce(fused, pb, 'mean')                             # se_layer.py:24-29
    se = torch.sigmoid(F.linear(F.relu(F.linear(se_in, sd['se.fc.0.weight'])), sd['se.fc.2.weight']))
    fused = fused + fused * se[pb]
    out = dict(bb)
    out['point_out'] = _mlp_seq(fused, sd, 'classifier.', CLASSIFIER)
    out['point_voxel_ids'] = pv
    if stats is not None:
        stats['points'] = int(points.shape[0])
    return out
