"""os3d_linear_bf16 (the Linear layers of the SWFormer encoder layer on the tcgen05 kernel, fused epilogues) against a
float64 torch restatement of nn.Linear / nn.GELU / nn.LayerNorm / residual on the same bf16-rounded operands
(seg3d/models/layers/point_transformer_layer.py:260-298).  Tolerance: bf16 rel 2e-2 of the output scale (north star)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _run(m, k, n, mode, seed=0):
    from openseg3d_b200.ops.linear import GELU, RELU, PackedLinearCache, linear_bf16
    torch.manual_seed(seed)
    x = torch.randn(m, k).bfloat16()
    w = (torch.randn(n, k) / np.sqrt(k)).bfloat16().float()
    b = torch.randn(n)
    cache = PackedLinearCache()
    chunks = cache.get('w', w.cuda(), b.cuda(), max_width=512 if mode == 'ln' else 256)
    ref = x.double() @ w.double().t() + b.double()
    kw = {}
    if mode == 'gelu':
        kw['flags'] = GELU
        ref = F.gelu(ref)
    elif mode == 'relu':
        kw['flags'] = RELU
        ref = ref.clamp(min=0)
    elif mode == 'ln':
        gamma, beta, res = torch.rand(n) + 0.5, torch.randn(n), torch.randn(m, n).bfloat16()
        kw.update(ln=(gamma.cuda(), beta.cuda(), 1e-5), residual=res.cuda())
        ref = res.double() + F.layer_norm(ref, (n,), gamma.double(), beta.double(), 1e-5)
    elif mode == 'table':
        tab = torch.randn(800, n).bfloat16()
        idx = torch.randint(0, 800, (m,), dtype=torch.int32)
        step = chunks[0][3]
        kw.update(table=[tab[:, o:o + step].contiguous().cuda() for _, _, o, _ in chunks], tab_idx=idx.cuda())
        ref = ref + tab.double()[idx.long()]
    y = linear_bf16(x.cuda(), chunks, **kw)
    assert y.dtype == torch.bfloat16 and y.shape == (m, n)
    got = y.float().cpu().double()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    mean_err = (got - ref).abs().mean().item() / ref.abs().mean().item()
    assert err < 2e-2 and mean_err < 4e-3, (err, mean_err)


@pytest.mark.parametrize('m,k,n,mode', [
    (1000, 48, 256, 'table'), (5000, 96, 256, 'plain'), (3000, 192, 512, 'table'), (777, 384, 768, 'table'),
    (4000, 128, 48, 'ln'), (4000, 128, 96, 'ln'), (2500, 256, 192, 'ln'), (1300, 384, 384, 'ln'),
    (4100, 48, 96, 'gelu'), (2000, 192, 384, 'gelu'), (900, 384, 768, 'gelu'), (2000, 768, 384, 'ln'),
    (129, 96, 48, 'relu'), (1, 64, 64, 'plain'), (40000, 96, 192, 'gelu')])
def test_linear_bf16_epilogues(m, k, n, mode):
    _run(m, k, n, mode)


def test_linear_bf16_rejects_bad_input():
    from openseg3d_b200.ops.linear import PackedLinearCache, linear_bf16
    cache = PackedLinearCache()
    with pytest.raises(RuntimeError):
        cache.get('w', torch.randn(48, 50).cuda())                       # in_features not a multiple of 8
    chunks = cache.get('w2', torch.randn(48, 64).cuda())
    with pytest.raises(RuntimeError):
        linear_bf16(torch.randn(10, 64).cuda(), chunks)                  # fp32 activations
    with pytest.raises(RuntimeError):
        linear_bf16(torch.randn(10, 64).bfloat16(), chunks)              # CPU tensor
    assert linear_bf16(torch.zeros(0, 64, dtype=torch.bfloat16).cuda(), chunks).shape == (0, 48)


def test_gelu_bf16_matches_erf_gelu():
    """os3d_gelu_bf16 (fast erf, |error| <= 1.5e-7) against torch's exact erf GELU in float64: never more than one bf16
    ulp apart, bit-identical after rounding for > 98 % of inputs."""
    from openseg3d_b200 import _lib
    x = torch.cat([torch.linspace(-12, 12, 80000), torch.randn(19992) * 3, torch.tensor([0.0, -0.0, 1e-8, -1e-8, 30.0, -30.0, 5.5, -5.5])])
    x = x.bfloat16().cuda()
    out = torch.empty_like(x)
    _lib.call('os3d_gelu_bf16', x, x.numel(), out)
    ref = F.gelu(x.double()).cpu()
    got = out.double().cpu()
    ulp = ref.abs() * 2.0 ** -7 + 1e-12          # one bf16 ulp; results below 1e-12 in magnitude are all 'zero'
    assert bool(((got - ref).abs() <= ulp).all()), float(((got - ref).abs() / ulp).max())
    big = ref.abs() > 1e-6                         # below that float64 erf saturates to exactly -1 (ref = -0.0)
    exact = (out.cpu()[big] == ref.bfloat16()[big]).float().mean().item()
    assert exact > 0.98, exact                       # measured 98.7 %: the rest differ by one bf16 ulp
