"""SubMConv3d / SparseConv3d / SparseInverseConv3d / SparseSequential over libos3d (stage 3).

Parameter names and layouts follow spconv 2.x so reference checkpoints load unchanged: ``weight``
[Cout, kz, ky, kx, Cin], optional ``bias`` [Cout].
"""
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib
from .rulebook import build_strided_rulebook, build_subm_rulebook
from .tensor import SparseConvTensor


class SparseModule(nn.Module):
    """Marker base class: SparseSequential hands these the SparseConvTensor, everything else the features."""
    pass


def _triple(v):
    return tuple(v) if isinstance(v, (list, tuple)) else (v, v, v)


def sparse_conv_forward(features, nbr, weight, bias, packed_cache, scale=None, shift=None, residual=None, relu=False,
                        out=None):
    """out[r] = epilogue( sum_k features[nbr[r, k]] @ W[:, k, :].T ).

    fp32 features -> exact FP32-pipe kernel (+ torch epilogue); bf16 features -> tcgen05 kernel with the epilogue
    (scale/shift = folded bias + BatchNorm, residual, ReLU) fused.  ``relu`` is the C ABI's flag word: bit 0 = ReLU,
    bit 1 = ``residual`` has 2*cout channels and residual[:, 2c] + residual[:, 2c+1] is added after the ReLU.
    ``out``: optional destination [m_out, cout] with unit column stride -- e.g. one half of a wider row-major buffer."""
    _lib.require_cuda(features, nbr)
    features = features.contiguous()
    m_out, cout, cin = nbr.shape[0], weight.shape[0], weight.shape[-1]
    if features.shape[1] != cin:
        raise RuntimeError(f'sparse conv: features have {features.shape[1]} channels, weight expects {cin}')
    dst = out
    if dst is not None and (tuple(dst.shape) != (m_out, cout) or dst.stride(1) != 1 or dst.dtype != features.dtype):
        raise RuntimeError('sparse conv: bad output buffer')
    if features.dtype == torch.float32:
        w = packed_cache.get('f32', weight)
        out = torch.empty((m_out, cout), dtype=torch.float32, device=features.device)
        _lib.call('os3d_spconv_fwd_f32', features, nbr, m_out, cin, cout, w, bias if scale is None else None, out,
                  work=lambda: _conv_work(nbr, features.shape[0], cin, cout, 4, residual))
        if scale is not None:
            out = out * scale + shift
        flags = int(relu)
        if residual is not None and not flags & 2:
            out = out + residual
        if flags & 1:
            out = F.relu_(out)
        if residual is not None and flags & 2:       # UpBlock: + channel_reduction(residual) after the ReLU
            out = out + residual.view(m_out, cout, 2).sum(dim=2)
        if dst is not None:
            dst.copy_(out)
            return dst
        return out
    if features.dtype == torch.bfloat16:
        w, cin_pad = packed_cache.get('bf16', weight)
        if cin_pad != cin:
            features = F.pad(features, (0, cin_pad - cin))
        if scale is None and bias is not None:
            scale, shift = torch.ones_like(bias, dtype=torch.float32), bias.float()
        out = dst if dst is not None else torch.empty((m_out, cout), dtype=torch.bfloat16, device=features.device)
        if m_out == 0 or features.shape[0] == 0:
            return out.zero_()
        nbr_t, tile_mask, perm = kernel_map_tiles(nbr)
        _lib.call('os3d_spconv_fwd_bf16_ld', features, features.shape[0], nbr_t, tile_mask, perm, m_out, cin_pad, cout, w,
                  scale, shift, residual.contiguous() if residual is not None else None, int(relu),
                  _lib._Raw(out.data_ptr()), out.stride(0),
                  work=lambda: _conv_work(nbr, features.shape[0], cin, cout, 2, residual))
        return out
    raise RuntimeError(f'sparse conv: unsupported feature dtype {features.dtype}')


def _conv_work(nbr, m_in, cin, cout, es, residual):
    """Algorithmic work of one sparse conv (profiling only; one host read, cached on the map): FLOPs = 2 * pairs * Cin *
    Cout; bytes = input rows + output rows (+ residual rows) + the 27 map entries of every output row."""
    pairs = getattr(nbr, '_os3d_pairs', None)
    if pairs is None:
        pairs = nbr._os3d_pairs = int((nbr >= 0).sum().item())
    m_out = nbr.shape[0]
    nbytes = es * (m_in * cin + m_out * cout) + 27 * 4 * m_out
    if residual is not None:
        nbytes += residual.numel() * residual.element_size()
    return _lib.Work(2.0 * cin * cout * pairs, nbytes, dict(cin=cin, cout=cout, m_out=m_out, pairs=pairs))


def kernel_map_tiles(nbr):
    """(nbr_t [27, m_pad], tile_mask [m_pad / 128], perm [m]) of a kernel map -- its tile form in mask-grouped row order
    (os3d_kernel_map_order) -- built on first use and kept on the map tensor so every conv sharing the map (same
    ``indice_key``) reuses it."""
    hit = getattr(nbr, '_os3d_tiles', None)
    if hit is None:
        m = nbr.shape[0]
        n_tiles = (m + 127) // 128
        nbr_t = torch.empty((27, n_tiles * 128), dtype=torch.int32, device=nbr.device)
        tile_mask = torch.empty(n_tiles, dtype=torch.int32, device=nbr.device)
        perm = None
        if os.environ.get('OS3D_MAP_ORDER', '1') != '0':
            import ctypes
            scratch = torch.empty((3, m), dtype=torch.int32, device=nbr.device)      # keys, sorted keys, row ids
            perm = torch.empty(m, dtype=torch.int32, device=nbr.device)
            nbytes = ctypes.c_int64(0)
            _lib.lib().os3d_kernel_map_order_scratch(m, ctypes.byref(nbytes))
            temp = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device=nbr.device)
            _lib.call('os3d_kernel_map_order', nbr, m, scratch[0], scratch[1], scratch[2], perm, temp, nbytes.value,
                      work=lambda: m * 27 * 4 + m * 4 * 4)     # map read once; keys / row ids / perm written
        _lib.call('os3d_kernel_map_tiles', nbr, m, perm, nbr_t, tile_mask, work=lambda: 2 * m * 27 * 4 + m * 4)
        hit = nbr._os3d_tiles = (nbr_t, tile_mask, perm)
    return hit


class _SparseConvFunction(torch.autograd.Function):
    """Differentiable sparse convolution (conv + bias, no fused epilogue) for training.

    forward : out[r] = b + sum_k in[nbr[r, k]] W[k]^T                          -- the forward kernels
    dgrad   : din[j] = sum_k dout[nbr_bwd[j, k]] Wb[k]^T                       -- the SAME kernels on the transposed
              problem: the output-stationary map of the transposed conv is the strided pair's other table
              (SparseConv3d <-> SparseInverseConv3d) or, for a submanifold conv, the map itself with the offsets
              mirrored (nbr[r, k] = j  <=>  nbr[j, 26 - k] = r); Wb[k] = W[k]^T (W[26 - k]^T when mirrored).
    wgrad   : dW[:, k, :] = dout^T . in[nbr[:, k]]   (rows without a neighbour contribute zero) -- 27 library GEMMs
              over masked gathers; a native gathered-wgrad kernel is listed under "next" in DESIGN.md.
    """

    @staticmethod
    def forward(ctx, features, weight, bias, nbr, nbr_bwd, mirror, m_in, packed_cache):
        out = sparse_conv_forward(features.detach(), nbr, weight.detach(), None if bias is None else bias.detach(),
                                  packed_cache)
        ctx.save_for_backward(features, weight)
        ctx.nbr, ctx.nbr_bwd, ctx.mirror, ctx.has_bias, ctx.m_in = nbr, nbr_bwd, mirror, bias is not None, m_in
        return out

    @staticmethod
    def backward(ctx, gout):
        features, weight = ctx.saved_tensors
        gout = gout.contiguous()
        cout, cin = weight.shape[0], weight.shape[-1]
        gin = gw = gb = None
        if ctx.needs_input_grad[0]:
            w3 = weight.detach().reshape(cout, 27, cin)
            if ctx.mirror:
                w3 = w3.flip(1)
            wb = w3.permute(2, 1, 0).contiguous()                       # [cin, 27, cout]: a conv weight with the roles swapped
            pad = (-cin) % 16 if gout.dtype == torch.bfloat16 else 0    # tensor-core kernel: output channels % 16
            if pad:
                wb = F.pad(wb, (0, 0, 0, 0, 0, pad))
            parts = []
            n_out = cin + pad
            step = n_out if n_out <= 512 else n_out // ((n_out + 511) // 512)
            for off in range(0, n_out, step):
                parts.append(sparse_conv_forward(gout, ctx.nbr_bwd, wb[off:off + step].reshape(-1, 3, 3, 3, cout), None,
                                                 _PackedWeights()))
            gin = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
            if pad:
                gin = gin[:, :cin].contiguous()
        if ctx.needs_input_grad[1]:
            # dW[:, k, :] = sum over the PAIRS of offset k of dout[out row]^T in[in row]: only rows that have a neighbour at
            # offset k enter the product (6-17 of 27 offsets per row on lidar frames), from a pair list built once per
            # kernel map (one host read of the 27 pair counts) and shared by every conv that uses the map
            out_rows, in_rows, bounds = _pair_lists(ctx.nbr)
            gw = torch.zeros((cout, 27, cin), dtype=torch.float32, device=weight.device)
            x = features.detach()
            for k in range(27):
                lo, hi = bounds[k], bounds[k + 1]
                if hi > lo:
                    gw[:, k, :] = torch.mm(gout.index_select(0, out_rows[lo:hi]).t(), x.index_select(0, in_rows[lo:hi])).float()
            gw = gw.reshape(weight.shape).to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gout.float().sum(dim=0)
        return gin, gw, gb, None, None, None, None, None


def _pair_lists(nbr):
    """(output rows, input rows, bounds[28]) of a kernel map's pairs grouped by offset (offset-major, ascending output
    row inside an offset) -- the per-offset pair lists spconv calls indice pairs; cached on the map tensor."""
    hit = getattr(nbr, '_os3d_pairs_by_offset', None)
    if hit is None:
        nt = nbr.t()                                              # [27, m_out]
        kk, rr = torch.nonzero(nt >= 0, as_tuple=True)            # sorted by offset, then output row
        counts = torch.bincount(kk, minlength=27).tolist()        # one host read per kernel map
        bounds = [0]
        for c in counts:
            bounds.append(bounds[-1] + int(c))
        hit = nbr._os3d_pairs_by_offset = (rr.contiguous(), nt[kk, rr].long().contiguous(), bounds)
    return hit


class _PackedWeights(object):
    """Kernel-layout copies of a conv weight, rebuilt when the parameter changes (version counter / storage)."""

    def __init__(self):
        self._store = {}

    def get(self, kind, weight):
        tag = (weight.data_ptr(), weight._version, weight.device)
        hit = self._store.get(kind)
        if hit is not None and hit[0] == tag:
            return hit[1]
        cout, cin = weight.shape[0], weight.shape[-1]
        w32 = weight.detach().float().contiguous()
        if kind == 'f32':
            packed = torch.empty((27, cin, cout), dtype=torch.float32, device=weight.device)
            _lib.call('os3d_pack_weight_f32', w32, cin, cout, packed)
            val = packed
        else:
            cin_pad = (cin + 7) // 8 * 8
            import ctypes
            elems = ctypes.c_int64(0)
            _lib.lib().os3d_spconv_bf16_packed_elems(cin_pad, cout, ctypes.byref(elems))
            packed = torch.empty(elems.value, dtype=torch.bfloat16, device=weight.device)
            _lib.call('os3d_pack_weight_bf16', w32, cin, cout, packed)
            val = (packed, cin_pad)
        self._store[kind] = (tag, val)
        return val


class _SparseConvBase(SparseModule):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation, bias, indice_key):
        super().__init__()
        if _triple(kernel_size) != (3, 3, 3) or _triple(dilation) != (1, 1, 1):
            raise NotImplementedError('openseg3d_b200.spconv implements 3x3x3 kernels with dilation 1 (all the reference uses)')
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding, self.dilation = _triple(kernel_size), _triple(stride), _triple(padding), (1, 1, 1)
        self.indice_key = indice_key
        self.weight = nn.Parameter(torch.empty(out_channels, 3, 3, 3, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self._packed = _PackedWeights()
        self.reset_parameters()

    def reset_parameters(self):
        # same scheme torch / spconv use for conv layers: kaiming_uniform(a=sqrt(5)) over fan_in = 27 * Cin
        fan_in = 27 * self.in_channels
        bound = math.sqrt(6.0 / ((1 + 5) * fan_in))
        nn.init.uniform_(self.weight, -bound, bound)
        if self.bias is not None:
            nn.init.uniform_(self.bias, -1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in))

    def extra_repr(self):
        return (f'{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, '
                f'padding={self.padding}, bias={self.bias is not None}, indice_key={self.indice_key}')

    # subclasses: _table(x) -> (nbr, out SparseConvTensor prototype); _table_bwd(x) -> (map of the transposed conv, mirror)
    def forward(self, x, scale=None, shift=None, residual=None, relu=False, out=None, wide_out=False):
        """``out``: destination view for the features (inference).  ``wide_out``: allocate a [m_out, 2 * cout] buffer, write
        the features into its left half and hang the buffer on the result as ``_os3d_wide`` -- the decoder block that
        consumes this tensor as x_bottom writes its lateral branch into the right half instead of torch.cat."""
        nbr, out_proto = self._table(x)
        wide = None
        if wide_out and out is None and not torch.is_grad_enabled():
            wide = torch.empty((nbr.shape[0], 2 * self.out_channels), dtype=x.features.dtype, device=x.features.device)
            out = wide[:, :self.out_channels]
        if torch.is_grad_enabled() and (x.features.requires_grad or self.weight.requires_grad):
            if scale is not None or residual is not None or relu or out is not None:
                raise RuntimeError('the fused conv epilogue is an inference path; training runs conv, BatchNorm, ReLU separately')
            nbr_bwd, mirror = self._table_bwd(x)
            feats = _SparseConvFunction.apply(x.features, self.weight, self.bias, nbr, nbr_bwd, mirror, x.features.shape[0],
                                              self._packed)
        else:
            feats = sparse_conv_forward(x.features, nbr, self.weight, self.bias, self._packed, scale, shift, residual, relu,
                                        out=out)
        res = out_proto(feats)
        if wide is not None:
            res._os3d_wide = wide
        return res


class SubMConv3d(_SparseConvBase):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None):
        super().__init__(in_channels, out_channels, kernel_size, 1, 1, dilation, bias, indice_key)

    def _table(self, x):
        rb = x.find_indice_pair(self.indice_key)
        if rb is None or rb.kind != 'subm' or rb.indices is not x.indices:
            rb = build_subm_rulebook(x)
            if self.indice_key is not None:
                x.indice_dict[self.indice_key] = rb
        self._last_rb = rb
        return rb.nbr, x.replace_feature

    def _table_bwd(self, x):
        return self._last_rb.nbr, True          # the submanifold map is its own transpose with the offsets mirrored


class SparseConv3d(_SparseConvBase):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None):
        if _triple(stride) != (2, 2, 2) or _triple(padding) != (1, 1, 1):
            raise NotImplementedError('SparseConv3d: kernel 3 / stride 2 / padding 1 only (all the reference uses)')
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, bias, indice_key)

    def _table(self, x):
        rb = x.find_indice_pair(self.indice_key)
        if rb is None or rb.kind != 'strided' or rb.in_indices is not x.indices:
            rb = build_strided_rulebook(x)
            if self.indice_key is not None:
                x.indice_dict[self.indice_key] = rb
        self._last_rb = rb
        return rb.fwd_nbr, lambda f: SparseConvTensor(f, rb.out_indices, rb.out_shape, x.batch_size, x.indice_dict,
                                                      x._site_table)

    def _table_bwd(self, x):
        return self._last_rb.inv_nbr, False


class SparseInverseConv3d(_SparseConvBase):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key=None, bias=True, algo=None):
        super().__init__(in_channels, out_channels, kernel_size, 1, 0, 1, bias, indice_key)

    def _table(self, x):
        rb = x.find_indice_pair(self.indice_key)
        if rb is None or rb.kind != 'strided':
            raise RuntimeError(f'SparseInverseConv3d needs the rulebook of an earlier SparseConv3d with indice_key={self.indice_key}')
        if rb.out_indices.shape[0] != x.features.shape[0]:
            raise RuntimeError('SparseInverseConv3d input does not live on the sites that SparseConv3d produced')
        return rb.inv_nbr, lambda f: SparseConvTensor(f, rb.in_indices, rb.in_shape, x.batch_size, x.indice_dict,
                                                      x._site_table)

    def _table_bwd(self, x):
        return x.find_indice_pair(self.indice_key).fwd_nbr, False


def bn_scale_shift(bn, conv_bias=None):
    """Eval-mode BatchNorm1d (and the conv bias before it) as y = acc * scale + shift, in fp32.  Cached on the module and
    recomputed only when a parameter or running statistic changes (seven tiny kernels per conv per forward otherwise)."""
    srcs = [bn.running_var, bn.running_mean] + ([bn.weight, bn.bias] if bn.affine else []) + \
           ([conv_bias] if conv_bias is not None else [])
    tag = tuple((t.data_ptr(), t._version) for t in srcs)
    hit = bn.__dict__.get('_os3d_fold')
    if hit is not None and hit[0] == tag:
        return hit[1]
    with torch.no_grad():
        inv = torch.rsqrt(bn.running_var.float() + bn.eps)
        scale = inv * bn.weight.float() if bn.affine else inv
        shift = -bn.running_mean.float() * scale
        if bn.affine:
            shift = shift + bn.bias.float()
        if conv_bias is not None:
            shift = shift + conv_bias.float() * scale
        out = (scale.contiguous(), shift.contiguous())
    bn.__dict__['_os3d_fold'] = (tag, out)
    return out


class SparseSequential(SparseModule):
    """Applies SparseModules to the SparseConvTensor and plain nn.Modules to its ``.features``.

    At inference a (conv, BatchNorm1d, ReLU) run is executed as ONE kernel: BatchNorm folds into a per-channel
    scale/shift applied in the conv epilogue together with the ReLU."""

    def __init__(self, *args):
        super().__init__()
        for i, m in enumerate(args):
            self.add_module(str(i), m)

    def __getitem__(self, i):
        return list(self._modules.values())[i]

    def __len__(self):
        return len(self._modules)

    def forward(self, x, wide_out=False):
        """``wide_out`` (inference): the last fused (conv, BatchNorm[, ReLU]) group writes into the left half of a double-width
        buffer (see _SparseConvBase.forward)."""
        mods = list(self._modules.values())
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, _SparseConvBase) and not self.training and not torch.is_grad_enabled() and i + 1 < len(mods) \
                    and isinstance(mods[i + 1], nn.BatchNorm1d) and isinstance(x, SparseConvTensor):
                relu = i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
                scale, shift = bn_scale_shift(mods[i + 1], m.bias)
                step = 3 if relu else 2
                x = m(x, scale=scale, shift=shift, relu=relu, wide_out=wide_out and i + step == len(mods))
                i += step
                continue
            if isinstance(m, SparseModule):
                x = m(x)
            elif isinstance(x, SparseConvTensor):
                if x.features.shape[0] > 0:
                    x = x.replace_feature(m(x.features))
            else:
                x = m(x)
            i += 1
        return x
