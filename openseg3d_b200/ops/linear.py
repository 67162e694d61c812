"""nn.Linear on the tcgen05 kernel with the SWFormer epilogues fused (os3d_linear_bf16): bias, ReLU / exact GELU,
residual + LayerNorm, and the position-embedding table term of the q / k projection.  bf16 activations only; there is
no fallback -- other dtypes raise."""
import ctypes
import os

import torch

from .. import _lib

RELU, GELU, LAYERNORM, TABLE = 1, 4, 8, 16
_USE_TC = os.environ.get('OS3D_LINEAR_TC', '1') != '0'


class PackedLinearCache(object):
    """Kernel-layout copies of Linear weights (the UMMA B-operand image) and fp32 biases, rebuilt when the source
    parameter changes (version counter / storage)."""

    def __init__(self):
        self._store = {}

    def get(self, key, weight, bias=None, max_width=256):
        """``max_width``: widest output chunk per launch.  256 keeps two CTAs per SM (tensor memory: 512 columns);
        the LayerNorm epilogue needs the whole row in one chunk (pass 512)."""
        tag = (weight.data_ptr(), weight._version, weight.device,
               None if bias is None else (bias.data_ptr(), bias._version))
        hit = self._store.get(key)
        if hit is not None and hit[0] == tag:
            return hit[1]
        n, k = weight.shape
        if k % 8:
            raise RuntimeError(f'linear_bf16: in_features {k} must be a multiple of 8')
        w32 = weight.detach().float().contiguous()
        chunks = []
        n_chunks = (n + max_width - 1) // max_width
        while n % n_chunks or (n // n_chunks) % 16:
            n_chunks += 1
        step = n // n_chunks
        if n % n_chunks or step % (32 if step > 256 else 16):
            raise RuntimeError(f'linear_bf16: out_features {n} cannot be cut into multiples of 16 (32 above 256)')
        for c in range(n_chunks):
            elems = ctypes.c_int64(0)
            _lib.lib().os3d_linear_bf16_packed_elems(k, step, ctypes.byref(elems))
            packed = torch.empty(elems.value, dtype=torch.bfloat16, device=weight.device)
            _lib.call('os3d_pack_linear_bf16', w32[c * step:(c + 1) * step].contiguous(), k, step, packed)
            b = None if bias is None else bias.detach().float()[c * step:(c + 1) * step].contiguous()
            chunks.append((packed, b, c * step, step))
        self._store[key] = (tag, chunks)
        return chunks


def linear_bf16(x, chunks, flags=0, residual=None, ln=None, table=None, tab_idx=None):
    """out = epilogue(x @ W.T + b).  ``chunks``: PackedLinearCache.get(...).  ln = (gamma f32, beta f32, eps);
    table: list of per-chunk [rows, width] bf16 tables (or one tensor when there is one chunk), tab_idx int32 [M]."""
    _lib.require_cuda(x)
    if x.dtype != torch.bfloat16:
        raise RuntimeError('linear_bf16 takes bfloat16 activations')
    x = x.contiguous()
    m, k = x.shape
    n = sum(c[3] for c in chunks)
    out = torch.empty((m, n), dtype=torch.bfloat16, device=x.device)
    if m == 0:
        return out
    if ln is not None:
        flags |= LAYERNORM
        if len(chunks) != 1:
            raise RuntimeError('linear_bf16: the LayerNorm epilogue needs the whole row in one tile (out_features <= 512)')
    if table is not None:
        flags |= TABLE
        if isinstance(table, torch.Tensor):
            table = [table]
    gamma, beta, eps = ln if ln is not None else (None, None, 0.0)
    for i, (packed, bias, off, width) in enumerate(chunks):
        tab = table[i] if table is not None else None
        dst = out if off == 0 and width == n else out[:, off:]
        # persistent kernel (weights resident, overlapped epilogue) when the weights fit; else the conv kernel's dense mode
        entry = 'os3d_linear_tc_bf16' if _lib.lib().os3d_linear_tc_fits(k, width) and _USE_TC else 'os3d_linear_bf16'
        _lib.call(entry, x, m, k, width, packed, bias, flags,
                  residual.contiguous() if residual is not None else None, gamma, beta, float(eps), tab, tab_idx,
                  tab.shape[1] if tab is not None else 0, _Ptr(dst), n,
                  work=lambda w=width: _lib.Work(2.0 * m * k * w,
                                                 2.0 * (m * k + m * w * (2 if residual is not None else 1) + k * w)
                                                 + (4.0 * m if tab is not None else 0.0)))
    return out


class _Ptr(object):
    """Raw device pointer argument (a column slice of a row-major tensor is not contiguous, its base pointer is all the
    kernel needs: the row pitch is passed separately)."""

    def __init__(self, t):
        self.ptr = t.data_ptr()


class WideLinear(object):
    """A wide Linear layer on os3d_wide_linear_bf16 (qkv_tc.cu): weights streamed in chunks of nc output columns, outputs
    staged and stored coalesced.  ``weight`` [n, k] (fp32 / any float), ``dp``: head width -- the granule of the per-head
    L2 normalisation of the first ``n_norm`` columns (the attention in-projection's q | k) or just the processing granule.
    ``fits(k, n, n_norm, dp)`` says whether the kernel takes the shape."""

    @staticmethod
    def fits(k, n, n_norm, dp):
        return bool(_lib.lib().os3d_wide_linear_plan(k, n, n_norm, dp, None))

    def __init__(self, weight, bias, dp, n_norm=0, normalize=False, gelu=False):
        n, k = weight.shape
        nc = ctypes.c_int(0)
        if not _lib.lib().os3d_wide_linear_plan(k, n, n_norm, dp, ctypes.byref(nc)):
            raise RuntimeError(f'wide linear: shape k={k} n={n} dp={dp} does not fit os3d_wide_linear_bf16')
        self.k, self.n, self.dp, self.nc, self.n_norm = k, n, dp, nc.value, n_norm
        self.normalize, self.mode_rest = int(normalize), 2 if gelu else 0
        w32 = weight.detach().float().contiguous()
        elems = ctypes.c_int64(0)
        _lib.lib().os3d_linear_bf16_packed_elems(k, self.nc, ctypes.byref(elems))
        self.w_img = torch.empty((n // self.nc, elems.value), dtype=torch.bfloat16, device=weight.device)
        for c in range(n // self.nc):
            _lib.call('os3d_pack_linear_bf16', w32[c * self.nc:(c + 1) * self.nc].contiguous(), k, self.nc, self.w_img[c])
        self.bias = None if bias is None else bias.detach().float().contiguous()

    def __call__(self, x, table=None, tab_idx=None, out=None):
        """x [m, k] bf16 -> [m, n] bf16.  table [rows, >= n_norm] bf16 + tab_idx int32 [m]: additive row term of the first
        n_norm columns (replaces the bias there)."""
        _lib.require_cuda(x)
        if x.dtype != torch.bfloat16 or x.shape[1] != self.k:
            raise RuntimeError('wide linear takes bfloat16 activations [m, k]')
        x = x.contiguous()
        m = x.shape[0]
        if out is None:
            out = torch.empty((m, self.n), dtype=torch.bfloat16, device=x.device)
        if m:
            _lib.call('os3d_wide_linear_bf16', x, m, self.k, self.n, self.dp, self.w_img, self.bias, table, tab_idx,
                      table.shape[1] if table is not None else 0, self.n_norm, self.normalize, self.mode_rest, out,
                      out.stride(0),
                      work=lambda: _lib.Work(2.0 * m * self.k * self.n, 2.0 * (m * self.k + m * self.n + self.k * self.n)
                                             + (4.0 * m + 2.0 * m * self.n_norm if table is not None else 0.0)))
        return out
