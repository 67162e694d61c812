"""Stage 4 modules: sparse window partition, cosine window attention, SWFormer block.

Class names, constructor arguments, forward signatures and state_dict keys mirror
seg3d/models/layers/point_transformer_layer.py:11-339 and seg3d/models/layers/cosine_msa.py:413-501, so a reference
checkpoint loads unchanged and PointTransformer can be written exactly as in the reference.  What differs is the
execution: windows are never padded to [R, T, C]; the partition layer emits window SEGMENTS over the flat voxel list
(no host sync) and attention runs variable-length over them (libos3d: os3d_window_partition, os3d_window_attention).
The padded per-level view the reference exposes (flat2win inds, padded pos-embed, key masks) is still available
through ``materialize()`` for callers and tests that want it.
"""
import ctypes
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.utils.checkpoint

from .. import _lib
from .._lib import WindowCfg
from ..ops.pooling import scatter_mean
from ..ops.linear import PackedLinearCache, WideLinear, linear_bf16
from ..ops.mlp_chain import GELU as MLP_GELU, NONE as MLP_NONE, MlpChain, SwformerMlp

_MLP_MODE = os.environ.get('OS3D_MLP_CHAIN', '1')
_QKV_MODE = os.environ.get('OS3D_QKV', '1')             # '0': in-projections as library GEMMs + separate table add (A-B runs)
# 'v1' (default): attention_tc.cu, one CTA per (128-query tile, head), 4-8 CTAs per SM.  'v2': attention_v2.cu, the
# warp-specialised all-heads-per-CTA design -- parity-tested, but measured slower at levels 1-2
_TRAIN_CKPT = os.environ.get('OS3D_TRAIN_CHECKPOINT', '1') != '0'   # 0: keep activations (180 GB of HBM) instead of recomputing each layer
_TRAIN_TC = os.environ.get('OS3D_TRAIN_ATTN_TC', '1') != '0'      # training forward on the tensor-core kernel (bf16, no dropout)
_ATTN_IMPL = os.environ.get('OS3D_ATTN', 'v1')


class WindowSegments(object):
    """Device-side result of one shift's partition (see include/os3d.h: os3d_window_partition)."""

    def __init__(self, cfg, batching_info, m):
        self.cfg, self.batching_info, self.m = cfg, batching_info, m
        self.lvl_tokens = (ctypes.c_int * 4)(*[cfg.lvl_tokens[i] for i in range(4)])

    def sum_sq_tokens(self):
        """sum over windows of n^2 (bench accounting only: useful attention FLOPs = 4 * sum n^2 * C); one host read, cached."""
        if not hasattr(self, '_sum_sq'):
            n = torch.unique(self.win_id, return_counts=True)[1].double()
            self._sum_sq = float((n * n).sum().item())
        return self._sum_sq

    def check_no_drop(self):
        """One host read: raises if any token would be dropped (the reference cannot continue either, App. C)."""
        info = self.level_info.cpu()
        if int(info[12]) or int(info[15]):
            raise RuntimeError(f'window batching would drop {int(info[12]) + int(info[15])} tokens; unsupported '
                               '(the reference desyncs features and indices in that case)')
        return info


class Flat2WinInds(dict):
    """``flat2win_inds_shift{i}``: carries the segment representation under string keys ('segments',
    'voxel_batching_level', 'batching_info'); ``materialize()`` adds the reference's per-level entries
    {level: (flat2window_inds, torch.where(mask))} (swformer_utils.py:8-31) -- that costs a host sync per level, which is
    why the hot path does not use them."""

    def materialize(self):
        seg = self['segments']
        lvl, rank, inner = seg.level, seg.win_rank.long(), seg.inner.long()
        for bl in seg.batching_info:
            mask = lvl == bl
            if not bool(mask.any()):
                continue
            t = seg.batching_info[bl]['max_tokens']
            self[bl] = (rank[mask] * t + inner[mask], torch.where(mask))
        return self

    def levels(self):
        return {k: v for k, v in self.items() if not isinstance(k, str)}


def flat2window(feat, inds):
    """Reference-layout scatter [M, C] -> {level: [R, T, C]} (swformer_utils.py:34-64); test / API helper."""
    if not inds.levels():
        inds.materialize()
    binfo = inds['batching_info']
    out = {}
    for bl, (slot, where) in inds.levels().items():
        t = binfo[bl]['max_tokens']
        r = int(torch.div(slot, t, rounding_mode='floor').max().item()) + 1
        buf = torch.zeros((r * t, feat.shape[-1]), dtype=feat.dtype, device=feat.device)
        buf[slot] = feat[where[0]]
        out[bl] = buf.reshape(r, t, -1)
    return out


def window2flat(feat_3d_dict, inds):
    """Inverse of flat2window (swformer_utils.py:67-85)."""
    first = next(iter(feat_3d_dict.values()))
    n = sum(v[0].shape[0] for v in inds.levels().values())
    out = torch.zeros((n, first.shape[-1]), dtype=first.dtype, device=first.device)
    for bl, f in feat_3d_dict.items():
        slot, where = inds[bl]
        out[where[0]] = f.reshape(-1, f.shape[-1])[slot]
    return out


class KeyMaskDict(dict):
    """``key_mask_shift{i}``: {level: [R, T] bool, True on padded slots} as get_key_padding_mask builds it
    (point_transformer_layer.py:210-220).  The hot path never pads, so the masks are materialised on first access
    (one host sync per level, like the reference)."""

    def __init__(self, inds):
        super().__init__()
        self._inds, self._done = inds, False

    def _ensure(self):
        if not self._done:
            self._done = True
            inds = self._inds
            if not inds.levels():
                inds.materialize()
            binfo = inds['batching_info']
            dev = inds['segments'].level.device
            for bl, (slot, where) in inds.levels().items():
                t = binfo[bl]['max_tokens']
                r = int(torch.div(slot, t, rounding_mode='floor').max().item()) + 1
                mask = torch.ones(r * t, dtype=torch.bool, device=dev)
                mask[slot] = False
                super().__setitem__(bl, mask.reshape(r, t))
        return self

    def __getitem__(self, k):
        return dict.__getitem__(self._ensure(), k)

    def __contains__(self, k):
        return dict.__contains__(self._ensure(), k)

    def __iter__(self):
        return dict.__iter__(self._ensure())

    def __len__(self):
        return dict.__len__(self._ensure())

    def keys(self):
        return dict.keys(self._ensure())

    def items(self):
        return dict.items(self._ensure())

    def values(self):
        return dict.values(self._ensure())

    def get(self, k, default=None):
        return dict.get(self._ensure(), k, default)


class _PartitionInfo(dict):
    """The dict SparseWindowPartitionLayer.forward returns.  'voxel_coords' (indices as int64, point_transformer_layer.py:45)
    and 'voxel_keep_inds' (arange: nothing is ever dropped) are materialised on first access -- the hot path reads neither."""

    def __init__(self, indices):
        super().__init__()
        self._indices = indices

    def __missing__(self, key):
        if key == 'voxel_coords':
            val = self._indices.long()
        elif key == 'voxel_keep_inds':
            val = torch.arange(self._indices.shape[0], device=self._indices.device, dtype=torch.long)
        else:
            raise KeyError(key)
        self[key] = val
        return val

    def __contains__(self, key):
        return key in ('voxel_coords', 'voxel_keep_inds') or dict.__contains__(self, key)

    def keys(self):
        for k in ('voxel_coords', 'voxel_keep_inds'):
            self[k]
        return dict.keys(self)


class PosDict(dict):
    """``pos_dict_shift{i}``.  'flat' ([M, C] sinusoidal embedding, point_transformer_layer.py:152-207) is computed on
    first access: the bf16 hot path never needs it, because (x + pos) W^T = x W^T + pos W^T and pos takes only
    window-volume many values -- it uses ``table`` ([win_z * win_y * win_x, C] fp32) and ``pos_idx`` (int32 [M], the
    table row of each voxel) instead."""

    def __init__(self, layer, seg, feat_dim, dtype):
        super().__init__()
        self._layer, self._seg, self._c, self._dtype = layer, seg, feat_dim, dtype
        self._idx = None

    def __missing__(self, key):
        if key != 'flat':
            raise KeyError(key)
        self['flat'] = self._layer.get_pos_embed(self._seg, self._c, self._dtype)
        return self['flat']

    @property
    def table(self):
        return self._layer.pos_table(self._c, self._seg.in_win.device)

    def table_as(self, dtype):
        return self._layer.pos_table(self._c, self._seg.in_win.device, dtype)

    def add_to(self, x):
        """x + pos without materialising pos: one gather-add from the table (os3d_add_table_rows)."""
        x = x.contiguous()
        if x.shape[1] % 8:
            return x + self['flat'].to(x.dtype)
        out = torch.empty_like(x)
        _lib.call('os3d_add_table_rows', x, self.table_as(x.dtype), self.pos_idx, x.shape[0], x.shape[1], x.element_size(),
                  out, work=lambda: 2 * x.numel() * x.element_size() + x.shape[0] * 4)
        return out

    @property
    def pos_idx(self):
        if self._idx is None:
            self._idx = getattr(self._seg, 'pos_idx', None)          # written by os3d_window_partition
            if self._idx is None:
                wx, wy, wz = [int(w) for w in self._layer.window_shape]
                iw = self._seg.in_win                      # (z, y, x)
                self._idx = ((iw[:, 0] * wy + iw[:, 1]) * wx + iw[:, 2]).contiguous()
        return self._idx


class SparseWindowPartitionLayer(nn.Module):
    """Same constructor and output keys as point_transformer_layer.py:11-69.  forward(x) -> dict with
    batch_win_inds_shift{0,1}, coors_in_win_shift{0,1}, voxel_features, voxel_coords, voxel_keep_inds,
    voxel_batching_level_shift{0,1}, flat2win_inds_shift{0,1}, pos_dict_shift{0,1}, key_mask_shift{0,1}."""

    def __init__(self, batching_info, window_shape, sparse_shape, normalize_pos=False, pos_temperature=1000):
        super().__init__()
        if normalize_pos:
            raise NotImplementedError('normalize_pos=True is never used by the reference model')
        if len(window_shape) != 3:
            raise NotImplementedError('3-D windows only (the reference model uses [10, 10, 8])')
        self.batching_info = batching_info
        self.sparse_shape = sparse_shape
        self.window_shape = window_shape
        self.normalize_pos = normalize_pos
        self.pos_temperature = pos_temperature
        self._tables = {}
        self._may_drop = self._can_drop(batching_info, window_shape)

    @staticmethod
    def _can_drop(batching_info, window_shape):
        """True when some window occupancy 1..volume falls outside every batching range or above its level's max_tokens
        (batching_single_shift's keep_mask, point_transformer_layer.py:71-87).  False for the reference's configs."""
        volume = 1
        for w in window_shape:
            volume *= int(w)
        ok = [False] * (volume + 1)
        for info in batching_info.values():
            lo, hi = [int(v) for v in info['batching_range']]
            for n in range(max(lo, 1), min(hi, volume + 1)):
                if n <= int(info['max_tokens']):
                    ok[n] = True
        return not all(ok[1:])

    @torch.no_grad()
    def pos_table(self, feat_dim, device, dtype=torch.float32):
        """fp32 embedding of every in-window position, row (z * win_y + y) * win_x + x (cached, also per cast dtype)."""
        if dtype != torch.float32:
            ckey = (feat_dim, str(device), dtype)
            if ckey not in self._tables:
                self._tables[ckey] = self.pos_table(feat_dim, device).to(dtype).contiguous()
            return self._tables[ckey]
        key = (feat_dim, str(device))
        if key not in self._tables:
            wx, wy, wz = [int(w) for w in self.window_shape]
            z, y, x = torch.meshgrid(torch.arange(wz), torch.arange(wy), torch.arange(wx), indexing='ij')
            grid = torch.stack([z, y, x], dim=-1).reshape(-1, 3).int().to(device).contiguous()
            out = torch.empty((grid.shape[0], feat_dim), dtype=torch.float32, device=device)
            _lib.call('os3d_pos_embed', grid, grid.shape[0], feat_dim, wx, wy, wz, float(self.pos_temperature), 4, out)
            self._tables[key] = out
        return self._tables[key]

    def _cfg(self, do_shift):
        """get_window_coors' scalar setup, swformer_utils.py:109-131."""
        wx, wy, wz = [int(w) for w in self.window_shape]
        sx, sy, sz = [float(s) for s in self.sparse_shape]
        assert sz < sx, 'Usually holds... in case of wrong order'
        cfg = WindowCfg()
        cfg.sparse_x, cfg.sparse_y, cfg.sparse_z = int(sx), int(sy), int(sz)
        cfg.win_x, cfg.win_y, cfg.win_z = wx, wy, wz
        cfg.nwin_x, cfg.nwin_y, cfg.nwin_z = [int(np.ceil(s / w) + 1) for s, w in ((sx, wx), (sy, wy), (sz, wz))]
        shift = (wx // 2, wy // 2, wz // 2) if do_shift else (wx, wy, wz)
        cfg.shift_x, cfg.shift_y = shift[0], shift[1]
        cfg.shift_z = 0 if sz == wz else shift[2]
        levels = sorted(self.batching_info)
        if len(levels) > 4 or levels != list(range(len(levels))):
            raise NotImplementedError('batching levels must be 0..n-1 with n <= 4')
        cfg.n_levels = len(levels)
        for i, bl in enumerate(levels):
            cfg.lvl_lo[i], cfg.lvl_hi[i] = [int(v) for v in self.batching_info[bl]['batching_range']]
            cfg.lvl_tokens[i] = int(self.batching_info[bl]['max_tokens'])
        return cfg

    @torch.no_grad()
    def partition(self, indices, batch_size, do_shift):
        indices = indices.int().contiguous() if indices.dtype != torch.int32 else indices.contiguous()
        m, dev = indices.shape[0], indices.device
        cfg = self._cfg(do_shift)
        n_win = batch_size * cfg.nwin_x * cfg.nwin_y * cfg.nwin_z
        nb = (n_win + 1023) // 1024
        i32 = dict(dtype=torch.int32, device=dev)
        seg = WindowSegments(cfg, self.batching_info, m)
        seg.win_id = torch.empty(m, dtype=torch.int64, device=dev)
        seg.in_win = torch.empty((m, 3), **i32)
        seg.pos_idx = torch.empty(m, **i32)              # row in the window's position-embedding table
        seg.level, seg.win_rank, seg.inner = torch.empty(m, **i32), torch.empty(m, **i32), torch.empty(m, **i32)
        seg.order = torch.empty(m, **i32)
        seg.seg_start, seg.seg_len = torch.empty(m + 1, **i32), torch.empty(m + 1, **i32)
        seg.pos_seg = torch.empty((max(m, 1), 2), **i32)
        seg.level_info = torch.empty(16, **i32)
        win_count, win_meta = torch.empty(n_win, **i32), torch.empty(n_win * 3, **i32)
        block_sums = torch.empty((nb + 1) * 5, **i32)
        _lib.call('os3d_window_partition', indices, m, batch_size, ctypes.byref(cfg), win_count, win_meta, block_sums, nb,
                  seg.win_id, seg.in_win, seg.level, seg.win_rank, seg.inner, seg.order, seg.seg_start, seg.seg_len,
                  seg.pos_seg, seg.level_info, seg.pos_idx,
                  work=lambda: m * (16 + 4 * 13) + 8 * n_win)       # coords in; ids / ranks / order / segments out; histogram
        return seg

    @torch.no_grad()
    def get_pos_embed(self, seg, feat_dim, dtype):
        """Flat [M, C] sinusoidal embedding of the in-window coordinates (point_transformer_layer.py:152-205)."""
        out = torch.empty((seg.m, feat_dim), dtype=dtype, device=seg.in_win.device)
        wx, wy, wz = [int(w) for w in self.window_shape]
        _lib.call('os3d_pos_embed', seg.in_win, seg.m, feat_dim, wx, wy, wz, float(self.pos_temperature),
                  out.element_size(), out)
        return out

    def forward(self, x):
        feats, indices = x.features, x.indices
        batch_size = getattr(x, 'batch_size', None)
        if batch_size is None:
            batch_size = int(indices[:, 0].max().item()) + 1
        info = _PartitionInfo(indices)            # 'voxel_coords' / 'voxel_keep_inds' are built on first access
        info['voxel_features'] = feats
        for i in range(2):
            seg = self.partition(indices, batch_size, i == 1)
            info[f'batch_win_inds_shift{i}'] = seg.win_id
            info[f'coors_in_win_shift{i}'] = seg.in_win
            info[f'voxel_batching_level_shift{i}'] = seg.level
            info[f'flat2win_inds_shift{i}'] = Flat2WinInds(segments=seg, voxel_batching_level=seg.level,
                                                           batching_info=self.batching_info)
            info[f'pos_dict_shift{i}'] = PosDict(self, seg, feats.shape[1], feats.dtype)
            info[f'key_mask_shift{i}'] = KeyMaskDict(info[f'flat2win_inds_shift{i}'])    # materialised on first access
            if self._may_drop:
                # a batching_info that can leave tokens without a level or over a level's capacity: the reference drops
                # them and then desynchronises features and indices -- fail loudly instead of
                # returning rows no kernel writes.  One host read per partition, only for such configurations.
                seg.check_no_drop()
        return info


class _WindowAttentionFunction(torch.autograd.Function):
    """Differentiable variable-length cosine window attention (training path): forward = os3d_window_attention on
    already-normalised q / k (the normalisation itself stays in torch so autograd covers it), backward =
    os3d_window_attention_bwd.  Attention dropout uses a seed-hashed keep mask regenerated in the backward."""

    @staticmethod
    def forward(ctx, qn, kn, v, tau, tau_min, heads, seg, drop_p, seed):
        m, c = qn.shape
        qn, kn, v = qn.contiguous(), kn.contiguous(), v.contiguous()
        tau32 = tau.detach().float().reshape(1).contiguous()
        d = c // heads
        dp = (d + 15) // 16 * 16
        if qn.dtype == torch.bfloat16 and dp <= 48 and m > 0 and _TRAIN_TC:
            # forward on the tensor cores (os3d_window_attention_bf16_tc_drop): heads zero-padded to the MMA's K granule
            # in one buffer (the padding changes neither dot products nor norms), the padded output cut back to [M, C].
            # The backward kernel recomputes the probabilities from q / k / v and this output.
            if dp == d:
                qkv = torch.cat([qn, kn, v], dim=1)
            else:
                qkv = torch.zeros((m, 3 * heads, dp), dtype=qn.dtype, device=qn.device)
                qkv[:, :, :d] = torch.cat([qn, kn, v], dim=1).view(m, 3 * heads, d)
                qkv = qkv.view(m, 3 * heads * dp)
            hd = heads * dp
            out_p = torch.empty((m, hd), dtype=qn.dtype, device=qn.device)
            _lib.call('os3d_window_attention_bf16_tc_drop', qkv, _lib._Raw(qkv.data_ptr() + hd * 2),
                      _lib._Raw(qkv.data_ptr() + 2 * hd * 2), 3 * hd, 3 * hd, m, heads, dp, seg.order, seg.pos_seg,
                      seg.level_info, tau32, float(tau_min), float(drop_p), int(seed), out_p, hd)
            out = out_p if dp == d else out_p.view(m, heads, dp)[:, :, :d].reshape(m, c)
        else:
            out = torch.empty_like(qn)
            _lib.call('os3d_window_attention', qn, kn, v, c, c, m, c, heads, seg.order, seg.seg_start, seg.seg_len,
                      seg.level_info, ctypes.byref(seg.lvl_tokens), tau32, float(tau_min), float(drop_p), int(seed),
                      qn.element_size(), out)
        ctx.save_for_backward(qn, kn, v, out, tau32)
        ctx.args = (tau_min, heads, seg, drop_p, seed, tau.shape, tau.dtype)
        return out

    @staticmethod
    def backward(ctx, gout):
        qn, kn, v, out, tau32 = ctx.saved_tensors
        tau_min, heads, seg, drop_p, seed, tau_shape, tau_dtype = ctx.args
        m, c = qn.shape
        gout = gout.contiguous()
        gq, gk, gv = torch.empty_like(qn), torch.empty_like(kn), torch.empty_like(v)
        stats = torch.empty(m * heads * 3, dtype=torch.float32, device=qn.device)
        g_inv = torch.zeros(1, dtype=torch.float32, device=qn.device)
        _lib.call('os3d_window_attention_bwd', qn, kn, v, out, gout, m, c, heads, seg.order, seg.seg_start, seg.seg_len,
                  seg.level_info, ctypes.byref(seg.lvl_tokens), tau32, float(tau_min), float(drop_p), int(seed),
                  qn.element_size(), stats, gq, gk, gv, g_inv)
        # d(1 / max(tau, tau_min)) / d tau = -1 / tau^2 above the clamp, 0 below
        gtau = torch.where(tau32 > tau_min, -g_inv / (tau32 * tau32), torch.zeros_like(g_inv))
        return gq, gk, gv, gtau.reshape(tau_shape).to(tau_dtype), None, None, None, None, None


class CosineMultiheadAttention(nn.MultiheadAttention):
    """Parameters identical to the reference subclass (cosine_msa.py:413-431): in_proj_weight / in_proj_bias /
    out_proj.{weight,bias} from nn.MultiheadAttention plus the shared temperature ``tau`` [1, 1, 1]."""

    def __init__(self, embed_dim, num_heads, dropout=0., bias=True, batch_first=False, cosine=True, tau_min=0.01,
                 non_shared_tau=False):
        super().__init__(embed_dim, num_heads, dropout, bias)
        if not cosine or non_shared_tau:
            raise NotImplementedError('only the shared-tau cosine attention the reference model builds')
        self.tau_min = tau_min
        self.tau = nn.Parameter(torch.ones(1, 1, 1))
        self._cast = {}

    def params(self, dtype):
        """(in_proj_weight, in_proj_bias, out_proj.weight, out_proj.bias) in the compute dtype, cached."""
        if dtype == self.in_proj_weight.dtype:
            return self.in_proj_weight, self.in_proj_bias, self.out_proj.weight, self.out_proj.bias
        tag = (dtype, self.in_proj_weight._version, self.in_proj_bias._version, self.out_proj.weight._version,
               self.out_proj.bias._version, self.in_proj_weight.data_ptr())
        if self._cast.get('tag') != tag:
            self._cast = {'tag': tag, 'p': tuple(p.detach().to(dtype) for p in (
                self.in_proj_weight, self.in_proj_bias, self.out_proj.weight, self.out_proj.bias))}
        return self._cast['p']

    def params_head_padded(self, dtype):
        """Projection weights with every head's d channels padded to dp = ceil(d / 16) * 16 (zero rows / columns), the
        layout os3d_window_attention_bf16_tc consumes: (w_qk [2*H*dp, C], b_qk, w_v [H*dp, C], b_v, w_out [C, H*dp],
        b_out, dp).  Zero padding changes neither dot products nor norms."""
        tag = (dtype, self.in_proj_weight._version, self.in_proj_bias._version, self.out_proj.weight._version,
               self.out_proj.bias._version, self.in_proj_weight.data_ptr())
        if self._cast.get('pad_tag') != tag:
            c, h = self.embed_dim, self.num_heads
            d = c // h
            dp = (d + 15) // 16 * 16
            w_in, b_in = self.in_proj_weight.detach().float(), self.in_proj_bias.detach().float()

            def pad_rows(w, b):        # [C, C], [C] -> [H*dp, C], [H*dp]
                wp = w.new_zeros(h, dp, c)
                wp[:, :d] = w.view(h, d, c)
                bp = b.new_zeros(h, dp)
                bp[:, :d] = b.view(h, d)
                return wp.reshape(h * dp, c), bp.reshape(h * dp)

            wq, bq = pad_rows(w_in[:c], b_in[:c])
            wk, bk = pad_rows(w_in[c:2 * c], b_in[c:2 * c])
            wv, bv = pad_rows(w_in[2 * c:], b_in[2 * c:])
            wo = self.out_proj.weight.detach().float().new_zeros(c, h, dp)
            wo[:, :, :d] = self.out_proj.weight.detach().float().view(c, h, d)
            vals = (torch.cat([wq, wk]), torch.cat([bq, bk]), wv, bv, wo.reshape(c, h * dp), self.out_proj.bias.detach().float())
            self._cast['pad_tag'] = tag
            self._cast['pad'] = tuple(t.to(dtype).contiguous() for t in vals) + (dp,)
        return self._cast['pad']

    def _out_proj_chunks(self):
        """Head-padded output projection [C, H*dp] as a tensor-core image (PackedLinearCache chunks)."""
        tag = (self.out_proj.weight._version, self.out_proj.bias._version, self.out_proj.weight.data_ptr())
        hit = self.__dict__.get('_o_chunks')
        if hit is None or hit[0] != tag:
            _, _, _, _, w_o, b_o, _ = self.params_head_padded(torch.float32)
            hit = (tag, PackedLinearCache().get('o', w_o, b_o, max_width=512))
            self.__dict__['_o_chunks'] = hit
        return hit[1]

    def attention_heads(self, feat, pos_dict, seg):
        """bf16 tensor-core path up to (not including) the output projection: returns ([M, H*dp] heads, out-proj chunks).
        q, k are normalised inside the attention kernel's gather."""
        m = feat.shape[0]
        w_qk, b_qk, w_v, b_v, _, _, dp = self.params_head_padded(feat.dtype)
        o_c = self._out_proj_chunks()
        hd = self.num_heads * dp
        use_v2 = _ATTN_IMPL == 'v2' and self._fixed_max_ok() and (self.num_heads * dp) % (96 if dp == 48 else 128) == 0
        # q, k are L2-normalised in the projection's epilogue (free there) for head widths up to 32; a 48-wide head would
        # leave three of the four epilogue warps of a TMEM lane quarter idle (measured 0.53 vs 0.26 ms at level 4), so
        # level 4 keeps the normalisation in the attention kernel's gather unless the v2 kernel needs it
        prenorm = dp <= 32 or use_v2
        proj = self._qkv_proj(pos_dict, prenorm) if isinstance(pos_dict, PosDict) else None
        if proj is not None:
            # ONE kernel: q | k | v = x [W_q | W_k | W_v]^T with the position term as a table row ((x + pos) W^T =
            # x W^T + (pos W^T)[pos_idx], biases folded in), head-padded layout; q, k L2-normalised in its epilogue when
            # the attention kernel expects that (v2) -- os3d_wide_linear_bf16, no library GEMM, no gather-add pass
            lin, table = proj
            qkv = lin(feat, table=table, tab_idx=pos_dict.pos_idx)                           # [M, 3*H*dp]: q | k | v
            q_ptr, ld, ldv = qkv, 3 * hd, 3 * hd
            k_ptr, v_ptr = _lib._Raw(qkv.data_ptr() + hd * 2), _lib._Raw(qkv.data_ptr() + 2 * hd * 2)
            normalized = prenorm
        else:
            qk = F.linear(pos_dict.add_to(feat) if pos_dict is not None else feat, w_qk, b_qk)   # [M, 2*H*dp]: q | k
            v_ptr = F.linear(feat, w_v, b_v)                                                      # [M, H*dp]
            q_ptr, ld, ldv = qk, 2 * hd, hd
            k_ptr = _lib._Raw(qk.data_ptr() + hd * 2)
            normalized = False
        out = torch.empty((m, hd), dtype=feat.dtype, device=feat.device)
        work = lambda: _lib.Work(4.0 * seg.sum_sq_tokens() * self.embed_dim, 4 * m * hd * 2)     # noqa: E731  (q, k, v in; heads out)
        tau = self.tau.detach().float().reshape(1)
        if use_v2:
            # warp-specialised kernel, all heads of a group per CTA; q / k normalised beforehand, fixed-maximum softmax
            if not normalized:
                _lib.call('os3d_qk_normalize', q_ptr, k_ptr, ld, m, hd, self.num_heads, 2, work=lambda: 4 * m * hd * 2)
            _lib.call('os3d_window_attention_bf16_v2', q_ptr, k_ptr, v_ptr, ld, ldv, m, self.num_heads, dp, seg.order,
                      seg.pos_seg, seg.level_info, tau, float(self.tau_min), out, hd, work=work)
        else:
            _lib.call('os3d_window_attention_bf16_tc_prenorm' if normalized else 'os3d_window_attention_bf16_tc', q_ptr, k_ptr,
                      v_ptr, ld, ldv, m, self.num_heads, dp, seg.order, seg.pos_seg, seg.level_info, tau,
                      float(self.tau_min), out, hd, work=work)
        return out, o_c

    def _qkv_proj(self, pos_dict, normalize=False):
        """(WideLinear over [W_q | W_k | W_v] head-padded, position table [window volume, 2*H*dp] bf16 = pos W_qk^T + b_qk)
        for this layer and this partition layer's embedding table, or None when the shape is outside the kernel.  Rebuilt
        when a parameter changes."""
        if _QKV_MODE == '0':
            return None
        pos_tab = pos_dict.table                                     # fp32 [volume, C], cached on the partition layer
        tag = (self.in_proj_weight._version, self.in_proj_bias._version, self.in_proj_weight.data_ptr(), pos_tab.data_ptr(),
               bool(normalize))
        hit = self.__dict__.get('_qkv_hit')
        if hit is None or hit[0] != tag:
            w_qk, b_qk, w_v, b_v, _, _, dp = self.params_head_padded(torch.float32)
            hd = self.num_heads * dp
            val = None
            gran = dp if normalize else 16        # epilogue granule: a whole head only when it has to be normalised
            if WideLinear.fits(self.embed_dim, 3 * hd, 2 * hd, gran):
                with torch.no_grad():
                    w_all = torch.cat([w_qk, w_v], dim=0)                                        # [3*H*dp, C]
                    bias = torch.cat([torch.zeros_like(b_qk), b_v])                              # q | k bias lives in the table
                    table = (pos_tab @ w_qk.t() + b_qk).to(torch.bfloat16).contiguous()          # [volume, 2*H*dp]
                    val = (WideLinear(w_all, bias, gran, n_norm=2 * hd, normalize=normalize), table)
            hit = self.__dict__['_qkv_hit'] = (tag, val)
        return hit[1]

    def _fixed_max_ok(self):
        """True when log2(e) / max(tau, tau_min) <= 60: unit-vector scores are then within 120 binades of the fixed softmax
        maximum 1, so no row can underflow (os3d_window_attention_bf16_v2's precondition).  One host read of tau per
        parameter version (constant at inference)."""
        tag = (self.tau._version, self.tau.data_ptr())
        hit = self.__dict__.get('_tau_ok')
        if hit is None or hit[0] != tag:
            t = max(float(self.tau.detach().float().reshape(-1)[0].item()), float(self.tau_min))
            hit = self.__dict__['_tau_ok'] = (tag, 1.4426950408889634 / t <= 60.0)
        return hit[1]

    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, attn_mask=None):
        """The reference signature (cosine_msa.py:433-501) on padded windows: query / key / value [T, R, C] (sequence
        first), key_padding_mask [R, T] bool (True = padding).  Runs the COSINE attention of this class -- never the
        dot-product forward nn.MultiheadAttention would inherit -- by viewing every window's valid slots as one segment of
        the flat row list [T * R, C] and calling the same variable-length kernels as forward_segments().  Rows of padded
        slots come back as out_proj.bias (their attention output is zero).  Returns (out [T, R, C], weights): weights
        are the head-averaged probabilities [R, T, T] when need_weights (computed with plain torch ops; the
        reference's callers discard them, point_transformer_layer.py:253) else None."""
        if attn_mask is not None:
            raise NotImplementedError('attn_mask is never used by the reference model')
        if query.dim() != 3 or key.shape != query.shape or value.shape != query.shape:
            raise RuntimeError('CosineMultiheadAttention.forward expects query / key / value of shape [T, R, C]')
        if torch.is_grad_enabled() and (query.requires_grad or self.in_proj_weight.requires_grad):
            raise NotImplementedError('padded-window forward is an inference entry point; training goes through '
                                      'WindowAttention / forward_segments')
        _lib.require_cuda(query, key, value)
        t, r, c = query.shape
        h = self.num_heads
        dev = query.device
        valid = torch.ones((r, t), dtype=torch.bool, device=dev) if key_padding_mask is None else ~key_padding_mask.bool()
        w_in, b_in, w_out, b_out = self.params(query.dtype)
        q = F.linear(query.reshape(t * r, c), w_in[:c], b_in[:c])
        k = F.linear(key.reshape(t * r, c), w_in[c:2 * c], b_in[c:2 * c])
        v = F.linear(value.reshape(t * r, c), w_in[2 * c:], b_in[2 * c:]).contiguous()
        qk = torch.cat([q, k], dim=1).contiguous()                   # [T*R, 2C]: q | k, as forward_segments lays them out
        # segments: window r' = the valid slots of row r' of the mask, in slot order; flat row of slot (r', t') = t' * R + r'
        seg_len_all = valid.sum(dim=1).int()
        nz = seg_len_all > 0
        rt = torch.nonzero(valid)                                    # sorted by (window, slot)
        order = (rt[:, 1] * r + rt[:, 0]).int().contiguous()
        seg_len = seg_len_all[nz].contiguous()
        seg_start = (torch.cumsum(seg_len, 0) - seg_len).int().contiguous()
        n_win, n_tok = int(seg_len.shape[0]), int(order.shape[0])
        level_info = torch.zeros(16, dtype=torch.int32, device=dev)
        level_info[0], level_info[13], level_info[14] = n_win, n_win, n_tok
        es = query.element_size()
        k_ptr = _lib._Raw(qk.data_ptr() + c * es)
        _lib.call('os3d_qk_normalize', qk, k_ptr, 2 * c, t * r, c, h, es)
        out = torch.zeros((t * r, c), dtype=query.dtype, device=dev)
        if n_tok:
            lvl_tokens = (ctypes.c_int * 4)(t, 0, 0, 0)
            _lib.call('os3d_window_attention', qk, k_ptr, v, 2 * c, c, t * r, c, h, order, seg_start, seg_len, level_info,
                      ctypes.byref(lvl_tokens), self.tau.detach().float().reshape(1), float(self.tau_min), 0.0, 0, es, out)
        res = F.linear(out, w_out, b_out).reshape(t, r, c)
        weights = None
        if need_weights:
            d = c // h
            qn = qk[:, :c].reshape(t, r, h, d).permute(1, 2, 0, 3).float()           # [R, h, T, d] (already normalised)
            kn = qk[:, c:].reshape(t, r, h, d).permute(1, 2, 0, 3).float()
            s_ = qn @ kn.transpose(-2, -1) / self.tau.detach().float().reshape(()).clamp(min=self.tau_min)
            s_ = s_.masked_fill(~valid[:, None, None, :], float('-inf'))
            weights = s_.softmax(dim=-1).mean(dim=1).to(query.dtype)
        return res, weights

    def tensor_core_ok(self, feat):
        return feat.dtype == torch.bfloat16 and self.embed_dim // self.num_heads <= 48 and self.embed_dim % 8 == 0

    def forward_segments(self, feat, pos_dict, seg):
        """feat: flat [M, C]; pos_dict: PosDict or None; seg: WindowSegments.  Returns [M, C]."""
        m, c = feat.shape
        if torch.is_grad_enabled() and (feat.requires_grad or self.in_proj_weight.requires_grad):
            return self._forward_train(feat, pos_dict, seg)
        if self.tensor_core_ok(feat):
            heads, o_c = self.attention_heads(feat, pos_dict, seg)
            return linear_bf16(heads, o_c)
        w_in, b_in, w_out, b_out = self.params(feat.dtype)
        qk_in = (pos_dict.add_to(feat) if isinstance(pos_dict, PosDict) else feat + pos_dict['flat'].to(feat.dtype)) \
            if pos_dict is not None else feat
        qk = F.linear(qk_in, w_in[:2 * c], b_in[:2 * c])           # [M, 2C]: q | k   (q = k = x + pos)
        v = F.linear(feat, w_in[2 * c:], b_in[2 * c:])             # [M, C]           (v = x)
        es = feat.element_size()
        k_ptr = qk.data_ptr() + c * es                            # k = columns [C, 2C) of the same rows
        _lib.call('os3d_qk_normalize', qk, k_ptr, 2 * c, m, c, self.num_heads, es)
        out = torch.empty((m, c), dtype=feat.dtype, device=feat.device)
        _lib.call('os3d_window_attention', qk, k_ptr, v, 2 * c, c, m, c, self.num_heads, seg.order, seg.seg_start,
                  seg.seg_len, seg.level_info, ctypes.byref(seg.lvl_tokens), self.tau.detach().float().reshape(1),
                  float(self.tau_min), 0.0, 0, es, out)
        return F.linear(out, w_out, b_out)

    def _forward_train(self, feat, pos_dict, seg):
        """Differentiable path (cosine_multi_head_attention_forward, cosine_msa.py:180-408): projections and the q / k
        normalisation are torch ops (autograd), the attention core is _WindowAttentionFunction."""
        m, c = feat.shape
        h = self.num_heads
        dt = feat.dtype
        w_in, b_in = self.in_proj_weight.to(dt), self.in_proj_bias.to(dt)
        qk_in = feat if pos_dict is None else feat + pos_dict['flat'].to(dt)
        qk = F.linear(qk_in, w_in[:2 * c], b_in[:2 * c])
        v = F.linear(feat, w_in[2 * c:], b_in[2 * c:])
        qn = F.normalize(qk[:, :c].reshape(m, h, c // h).float(), dim=-1).reshape(m, c).to(dt)
        kn = F.normalize(qk[:, c:].reshape(m, h, c // h).float(), dim=-1).reshape(m, c).to(dt)
        drop_p = self.dropout if self.training else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if drop_p > 0 else 0
        out = _WindowAttentionFunction.apply(qn, kn, v, self.tau, self.tau_min, h, seg, drop_p, seed)
        return F.linear(out, self.out_proj.weight.to(dt), self.out_proj.bias.to(dt))


class WindowAttention(nn.Module):
    def __init__(self, d_model, nhead, attn_drop, cosine=True, tau_min=0.01):
        super().__init__()
        self.self_attn = CosineMultiheadAttention(d_model, nhead, dropout=attn_drop, batch_first=False, tau_min=tau_min,
                                                  cosine=cosine, non_shared_tau=False)

    def forward(self, feat_2d, pos_dict, ind_dict, key_padding_dict=None):
        """Same call as the reference (point_transformer_layer.py:233): feat_2d [M, C]; pos_dict / ind_dict are the
        partition layer's pos_dict_shift{i} / flat2win_inds_shift{i}; key_padding_dict is unused (no padding exists)."""
        if feat_2d.dtype not in (torch.float32, torch.bfloat16):
            raise RuntimeError('window attention runs in float32 or bfloat16')
        return self.self_attn.forward_segments(feat_2d.contiguous(), pos_dict, ind_dict['segments'])


def _cast_like(mod, name, dtype):
    p = getattr(mod, name)
    if p is None or p.dtype == dtype:
        return p
    if torch.is_grad_enabled() and p.requires_grad:
        return p.to(dtype)                    # differentiable cast: the gradient reaches the fp32 parameter
    cache = mod.__dict__.setdefault('_os3d_cast', {})
    tag = (dtype, p._version, p.data_ptr())
    hit = cache.get(name)
    if hit is None or hit[0] != tag:
        hit = (tag, p.detach().to(dtype))
        cache[name] = hit
    return hit[1]


def linear_in(mod, x):
    """nn.Linear applied in x's dtype (cached weight copies) -- explicit bf16 instead of autocast.  bf16 inference runs on
    the persistent tcgen05 Linear kernel (os3d_linear_tc_bf16; out_features padded to a multiple of 16 with zero rows:
    the 22-class voxel heads), everything else (fp32, training) on the library GEMM."""
    n, k = mod.weight.shape
    if (x.dtype == torch.bfloat16 and not torch.is_grad_enabled() and x.dim() == 2 and k % 8 == 0 and n <= 256
            and x.is_cuda and x.shape[0] > 0):
        hit = mod.__dict__.get('_os3d_tc')
        tag = (mod.weight.data_ptr(), mod.weight._version, None if mod.bias is None else mod.bias._version)
        if hit is None or hit[0] != tag:
            n_pad = (n + 15) // 16 * 16
            with torch.no_grad():
                w = mod.weight.new_zeros((n_pad, k), dtype=torch.float32)
                w[:n] = mod.weight.detach().float()
                b = None
                if mod.bias is not None:
                    b = mod.weight.new_zeros(n_pad, dtype=torch.float32)
                    b[:n] = mod.bias.detach().float()
            chunks = PackedLinearCache().get('w', w, b, max_width=256)
            hit = mod.__dict__['_os3d_tc'] = (tag, chunks if _lib.lib().os3d_linear_tc_fits(k, n_pad) else None)
        if hit[1] is not None:
            out = linear_bf16(x, hit[1])
            return out if out.shape[1] == n else out[:, :n]
    return F.linear(x, _cast_like(mod, 'weight', x.dtype), _cast_like(mod, 'bias', x.dtype))


def layer_norm_in(mod, x):
    return F.layer_norm(x, mod.normalized_shape, _cast_like(mod, 'weight', x.dtype), _cast_like(mod, 'bias', x.dtype),
                        mod.eps)


def _ln_params(mod):
    return (_cast_like(mod, 'weight', torch.float32), _cast_like(mod, 'bias', torch.float32), mod.eps)


def residual_layer_norm(mod, x, resid):
    """resid + LayerNorm(x) in one pass (os3d_layernorm_residual); falls back to torch ops for odd widths / training."""
    c = x.shape[-1]
    if (c % 8 or c > 1024 or x.dtype not in (torch.float32, torch.bfloat16) or not mod.elementwise_affine
            or (torch.is_grad_enabled() and (x.requires_grad or mod.weight.requires_grad))):
        return resid + layer_norm_in(mod, x)
    x, resid = x.contiguous(), resid.contiguous()
    out = torch.empty_like(x)
    _lib.call('os3d_layernorm_residual', x, resid, _cast_like(mod, 'weight', torch.float32),
              _cast_like(mod, 'bias', torch.float32), x.shape[0], c, float(mod.eps), x.element_size(), out)
    return out


class MLP(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return self.drop(linear_in(self.fc2, self.drop(self.act(linear_in(self.fc1, x)))))


class DropPath(nn.Module):
    """Stochastic depth per row (seg3d/models/layers/drop.py:4-34); identity at inference."""

    def __init__(self, drop_prob=0., scale_by_keep=True):
        super().__init__()
        self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0. or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


class EncoderLayer(nn.Module):
    """Post-norm residual layer, point_transformer_layer.py:278-298."""

    def __init__(self, d_model, nhead, mlp_hidden_dim=256, drop=0., attn_drop=0.1, drop_path=0.):
        super().__init__()
        self.win_attn = WindowAttention(d_model, nhead, attn_drop)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.mlp = MLP(in_features=d_model, hidden_features=mlp_hidden_dim, drop=drop)

    def forward(self, x, pos_dict, ind_dict, key_padding_mask_dict=None):
        mha = self.win_attn.self_attn
        if not self.training and not torch.is_grad_enabled() and mha.tensor_core_ok(x) and isinstance(pos_dict, PosDict):
            # bf16 inference: the two projections that are followed by residual + LayerNorm run on the tcgen05 kernel
            # with that epilogue fused (os3d_linear_bf16): the [M, C] projection output never goes to HBM
            x = x.contiguous()
            heads, o_c = mha.attention_heads(x, pos_dict, ind_dict['segments'])
            cache = self.__dict__.setdefault('_lin', PackedLinearCache())
            x1 = linear_bf16(heads, o_c, residual=x, ln=_ln_params(self.norm1))
            chain = self._mlp_chain()
            if chain is not None:
                # fc1 -> GELU -> fc2 -> LayerNorm + residual in one kernel: the [M, 2C] hidden tensor stays on chip
                return chain(x1, residual=x1, ln=_ln_params(self.norm2))
            # wider layers (weights beyond shared memory): fc1 is a library GEMM + the in-place GELU kernel --
            # os3d_linear_bf16's GELU epilogue measured the same (L3: 0.29 ms vs 0.12 + 0.16 ms)
            fc1 = self._fc1_wide()
            if fc1 is not None:
                h = fc1(x1)                                                                      # fc1 + GELU, one kernel
            else:
                h = linear_in(self.mlp.fc1, x1)
                _lib.call('os3d_gelu_bf16', h, h.numel(), h, work=lambda: 2 * h.numel() * 2)      # in place
            return linear_bf16(h, cache.get('fc2', self.mlp.fc2.weight, self.mlp.fc2.bias, max_width=512), residual=x1,
                               ln=_ln_params(self.norm2))
        attn = self.win_attn(x, pos_dict, ind_dict, key_padding_mask_dict)
        if self.training or torch.is_grad_enabled():
            x = x + self.drop_path(layer_norm_in(self.norm1, attn))
            return x + self.drop_path(layer_norm_in(self.norm2, self.mlp(x)))
        x = residual_layer_norm(self.norm1, attn, x)
        return residual_layer_norm(self.norm2, self.mlp(x), x)


    def _fc1_wide(self):
        """fc1 + GELU on os3d_wide_linear_bf16 for the layers whose MLP is beyond the fused kernel (level 4), or None."""
        if _QKV_MODE == '0':
            return None
        fc1 = self.mlp.fc1
        tag = (fc1.weight.data_ptr(), fc1.weight._version, fc1.bias._version)
        hit = self.__dict__.get('_fc1_hit')
        if hit is None or hit[0] != tag:
            n, k = fc1.weight.shape
            dp = next((d for d in (16, 32) if n % d == 0 and WideLinear.fits(k, n, 0, d)), None)
            val = WideLinear(fc1.weight, fc1.bias, dp, gelu=True) if dp is not None else None
            hit = self.__dict__['_fc1_hit'] = (tag, val)
        return hit[1]

    def _mlp_chain(self):
        """The fused MLP + norm2 + residual kernel for this layer, or None when its width is beyond the kernels (C > 192).
        Rebuilt when a parameter changes.  OS3D_MLP_CHAIN: 1 (default) = weights streamed (os3d_swformer_mlp_bf16, C <= 192),
        chain = weights resident (os3d_mlp_chain_bf16, C <= 96), 0 = unfused."""
        if _MLP_MODE == '0':
            return None
        fc1, fc2 = self.mlp.fc1, self.mlp.fc2
        tag = tuple((t.data_ptr(), t._version) for t in (fc1.weight, fc1.bias, fc2.weight, fc2.bias))
        hit = self.__dict__.get('_chain')
        if hit is None or hit[0] != tag:
            fn = None
            h, c = fc1.weight.shape
            if _MLP_MODE != 'chain' and SwformerMlp.fits(c, h):
                mlp = SwformerMlp(fc1.weight, fc1.bias, fc2.weight, fc2.bias)
                fn = lambda x, residual, ln: mlp(x, ln)                      # noqa: E731  (the residual is x itself)
            elif MlpChain.fits([(h, c), (c, h)]):
                fn = MlpChain([(fc1.weight, fc1.bias, MLP_GELU), (fc2.weight, fc2.bias, MLP_NONE)])
            hit = self.__dict__['_chain'] = (tag, fn)
        return hit[1]


class SWFormerBlock(nn.Module):
    """depth encoder layers: the first depth//2 on the unshifted windows, the rest on the half-window shift
    (point_transformer_layer.py:300-339)."""

    def __init__(self, d_model, nhead, depth=4, mlp_ratio=2., attn_drop=0.1, drop=0., drop_path=0.):
        super().__init__()
        self.depth = depth
        self.layers = nn.ModuleList([
            EncoderLayer(d_model, nhead, int(d_model * mlp_ratio), attn_drop=attn_drop, drop=drop,
                         drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path) for i in range(depth)])

    def forward(self, batch_dict, using_checkpoint=True):
        """Training with ``using_checkpoint`` (the reference's default, point_transformer_layer.py:321-323): every encoder
        layer runs under torch.utils.checkpoint -- only its input is kept, the layer is recomputed in the backward (the
        attention dropout mask is regenerated from its saved seed, torch's RNG state is restored for DropPath)."""
        x = batch_dict['voxel_features']
        ckpt = using_checkpoint and self.training and torch.is_grad_enabled() and _TRAIN_CKPT
        for i, layer in enumerate(self.layers):
            s = 0 if i < int(self.depth / 2) else 1
            args = (batch_dict[f'pos_dict_shift{s}'], batch_dict[f'flat2win_inds_shift{s}'], batch_dict[f'key_mask_shift{s}'])
            if ckpt and x.requires_grad:
                x = torch.utils.checkpoint.checkpoint(layer, x, *args, use_reentrant=False)
            else:
                x = layer(x, *args)
        return x


class FlattenSELayer(nn.Module):
    """Per-sample squeeze-excite over flat rows (seg3d/models/layers/se_layer.py:6-30); the scatter(mean) over the
    batch index goes through libos3d instead of torch_scatter."""

    def __init__(self, channel, reduction=4):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(channel, channel // reduction, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(channel // reduction, channel, bias=False), nn.Sigmoid())

    def gate(self, x, indices, batch_size=None):
        """Per-frame channel gate [B, C] (fp32): sigmoid(fc(mean over the frame's points))."""
        pooled = scatter_mean(x, indices.long(), batch_size).float()
        if torch.is_grad_enabled():
            return self.fc(pooled)
        # B rows x (C -> C / r -> C): two products of a handful of rows, done as broadcast multiply + reduce (no GEMM launch)
        h = torch.relu((pooled[:, None, :] * self.fc[0].weight.float()[None]).sum(-1))
        return torch.sigmoid((h[:, None, :] * self.fc[2].weight.float()[None]).sum(-1))

    def residual_forward(self, x, indices, batch_size=None):
        """x + forward(x) in one pass over the points (inference): x * (1 + gate[batch index])."""
        if torch.is_grad_enabled() or x.shape[1] % 8 or x.dtype not in (torch.float32, torch.bfloat16):
            return x + self.forward(x, indices, batch_size)
        indices = indices.long().contiguous()
        x = x.contiguous()
        out = torch.empty_like(x)
        _lib.call('os3d_scale_rows_by_table', x, self.gate(x, indices, batch_size).contiguous(), indices, x.shape[0],
                  x.shape[1], 1.0, x.element_size(), out, work=lambda: 2 * x.numel() * x.element_size() + x.shape[0] * 8)
        return out

    def forward(self, x, indices, batch_size=None):
        indices = indices.long()
        pooled = scatter_mean(x, indices, batch_size)
        gate = self.fc(pooled.float()).to(x.dtype)          # B rows: fp32 whatever the activation dtype
        # index_select, not gate[indices]: its backward is an index_add (atomics), advanced indexing's is a sort-based
        # index_put that takes 134 ms for 360k duplicates of 2 rows
        return x * gate.index_select(0, indices)
