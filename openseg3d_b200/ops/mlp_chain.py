"""A chain of Linear layers (bias, ReLU / exact GELU between them, optional residual + LayerNorm at the end) as ONE
persistent tcgen05 kernel with the activations kept on chip (os3d_mlp_chain_bf16, csrc/mlp_tc.cu).  Used at inference
for Segformer's point-wise MLPs with BatchNorm folded (seg3d/models/segmentors/segformer.py:21-32,58-76) and for the
SWFormer MLP (seg3d/models/layers/point_transformer_layer.py:260-298).  bf16 only; there is no fallback."""
import ctypes

import torch

from .. import _lib

NONE, RELU, GELU = 0, 1, 2


def _pack(w32):
    n, k = w32.shape
    elems = ctypes.c_int64(0)
    _lib.lib().os3d_linear_bf16_packed_elems(k, n, ctypes.byref(elems))
    packed = torch.empty(elems.value, dtype=torch.bfloat16, device=w32.device)
    _lib.call('os3d_pack_linear_bf16', w32.contiguous(), k, n, packed)
    return packed


class MlpChain(object):
    """``layers``: list of (weight fp32 [n, k], bias fp32 [n] or None, act) with act in {NONE, RELU, GELU}.
    ``front``: optional (weight fp32 [64, k32 <= 16], bias fp32 [64] or None, act) evaluated in fp32 inside the kernel
    from an fp32 input.  The last layer's width is zero-padded to a multiple of 16 (outputs beyond n_out are dropped)."""

    def __init__(self, layers, front=None):
        dev = layers[0][0].device
        self.n_out = layers[-1][0].shape[0]
        self.k_in = layers[0][0].shape[1]
        self._keep = []
        arr = (_lib.MlpLayer * len(layers))()
        self.flops_per_row = 0.0
        for i, (w, b, act) in enumerate(layers):
            w = w.detach().float()
            n, k = w.shape
            if i == len(layers) - 1 and n % 16:
                pad = 16 - n % 16
                w = torch.cat([w, w.new_zeros(pad, k)])
                b = None if b is None else torch.cat([b.detach().float(), w.new_zeros(pad)])
                n += pad
            packed = _pack(w)
            bias = None if b is None else b.detach().float().contiguous()
            self._keep += [packed, bias]
            arr[i].w, arr[i].bias = packed.data_ptr(), (bias.data_ptr() if bias is not None else None)
            arr[i].k, arr[i].n, arr[i].act = k, n, act
            self.flops_per_row += 2.0 * k * n
        self.layers, self.n_layers = arr, len(layers)
        self.front = None
        if front is not None:
            w, b, act = front
            if w.shape[0] != 64 or w.shape[1] > 16:
                raise RuntimeError('MlpChain: the fp32 front layer is [64, k <= 16]')
            w32 = w.detach().float().t().contiguous()                      # [k32][64]
            b32 = None if b is None else b.detach().float().contiguous()
            self.front = (w32, b32, int(act), w.shape[1])
            self.k_in = w.shape[1]
            self.flops_per_row += 2.0 * 64 * w.shape[1]
        if not _lib.lib().os3d_mlp_chain_fits(self.layers, self.n_layers, 1 if front is not None else 0):
            raise RuntimeError('MlpChain: this chain does not fit the kernel (2..4 layers, widths multiples of 16 up to '
                               '256, all weights resident in shared memory)')
        self.device = dev

    @staticmethod
    def fits(shapes, front=False):
        """shapes: list of (n, k)."""
        arr = (_lib.MlpLayer * len(shapes))()
        for i, (n, k) in enumerate(shapes):
            arr[i].w, arr[i].k, arr[i].n, arr[i].act = 1, k, (n + 15) // 16 * 16, 0
        return bool(_lib.lib().os3d_mlp_chain_fits(arr, len(shapes), 1 if front else 0))

    def __call__(self, x, residual=None, ln=None, out=None, out_dtype=torch.bfloat16):
        """x: bf16 [m, k] (row pitch = x.stride(0), unit column stride), or fp32 [m, k32] when the chain has a front layer.
        residual: bf16 [m, n]; ln = (gamma f32, beta f32, eps).  out: optional preallocated [m, >= n_out] view with unit
        column stride (e.g. a column slice of a wider buffer)."""
        _lib.require_cuda(x)
        m = x.shape[0]
        if x.dim() != 2 or x.shape[1] != self.k_in or x.stride(1) != 1:
            raise RuntimeError(f'MlpChain: input must be [m, {self.k_in}] with unit column stride')
        if out is None:
            out = torch.empty((m, self.n_out), dtype=out_dtype, device=x.device)
        elif out.shape[0] != m or out.shape[1] != self.n_out or out.stride(1) != 1 or out.dtype not in (torch.bfloat16, torch.float32):
            raise RuntimeError('MlpChain: bad output buffer')
        if m == 0:
            return out
        P = _lib._Raw
        if self.front is not None:
            if x.dtype != torch.float32:
                raise RuntimeError('MlpChain: the fp32 front layer takes float32 input')
            w32, b32, act32, k32 = self.front
            xa, ldx, x32a, ld32 = None, 0, P(x.data_ptr()), x.stride(0)
        else:
            if x.dtype != torch.bfloat16:
                raise RuntimeError('MlpChain takes bfloat16 activations')
            w32, b32, act32, k32 = None, None, 0, 0
            xa, ldx, x32a, ld32 = P(x.data_ptr()), x.stride(0), None, 0
        gamma, beta, eps = ln if ln is not None else (None, None, 0.0)
        if residual is not None and (residual.dtype != torch.bfloat16 or residual.stride(1) != 1):
            raise RuntimeError('MlpChain: residual must be bfloat16 with unit column stride')
        _lib.call('os3d_mlp_chain_bf16', xa, m, ldx, x32a, ld32, k32, w32, b32, act32, self.layers, self.n_layers,
                  P(residual.data_ptr()) if residual is not None else None,
                  residual.stride(0) if residual is not None else 0, gamma, beta, float(eps),
                  P(out.data_ptr()), out.stride(0), self.n_out, 1 if out.dtype == torch.float32 else 0,
                  work=lambda: _lib.Work(self.flops_per_row * m,
                                         m * (x.shape[1] * x.element_size() + self.n_out * out.element_size()
                                              + (self.n_out * 2 if residual is not None else 0))))
        return out


class SwformerMlp(object):
    """out = x + LayerNorm(fc2(GELU(fc1(x)))) as one kernel with the hidden tensor on chip and the weights streamed from L2
    (os3d_swformer_mlp_bf16, csrc/mlp2_tc.cu); C <= 192."""

    def __init__(self, w1, b1, w2, b2):
        self.h, self.c = w1.shape
        if tuple(w2.shape) != (self.c, self.h) or not self.fits(self.c, self.h):
            raise RuntimeError(f'SwformerMlp: fc1 {tuple(w1.shape)} / fc2 {tuple(w2.shape)} do not fit the kernel')
        self.w1, self.w2 = _pack(w1.detach().float()), _pack(w2.detach().float())
        self.b1 = None if b1 is None else b1.detach().float().contiguous()
        self.b2 = None if b2 is None else b2.detach().float().contiguous()

    @staticmethod
    def fits(c, h):
        return bool(_lib.lib().os3d_swformer_mlp_fits(c, h))

    def __call__(self, x, ln):
        _lib.require_cuda(x)
        if x.dtype != torch.bfloat16 or x.dim() != 2 or x.shape[1] != self.c:
            raise RuntimeError(f'SwformerMlp takes bfloat16 [m, {self.c}]')
        x = x.contiguous()
        out = torch.empty_like(x)
        gamma, beta, eps = ln
        m = x.shape[0]
        _lib.call('os3d_swformer_mlp_bf16', x, m, self.c, self.h, self.w1, self.b1, self.w2, self.b2, gamma, beta, float(eps),
                  out, work=lambda: _lib.Work(4.0 * m * self.c * self.h, 3.0 * m * self.c * 2))
        return out
