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


@pytest.mark.parametrize('k,heads,d,m', [(48, 8, 6, 1), (48, 8, 6, 70001), (96, 8, 12, 300), (192, 8, 24, 40000), (384, 8, 48, 20011)])
@pytest.mark.parametrize('normalize', [False, True])
def test_wide_linear_qkv_projection_matches_float64(k, heads, d, m, normalize):
    """os3d_wide_linear_bf16 as the attention in-projection: q | k | v = x [Wq | Wk | Wv]^T with the position term as a
    table row and (optionally) per-head L2 normalisation of q and k, head-padded layout -- against a float64 restatement
    of cosine_msa.py:48-63 + :152-153 on the same bf16-rounded operands.  Ragged row counts (1 row, not a multiple of the
    128-row tile), every level's width."""
    import torch.nn.functional as F
    from openseg3d_b200.ops.linear import WideLinear
    torch.manual_seed(k + m)
    dp = (d + 15) // 16 * 16
    hd = heads * dp
    x = torch.randn(m, k, device='cuda').bfloat16()
    w = torch.zeros(3, heads, dp, k, device='cuda')
    w[:, :, :d] = torch.randn(3, heads, d, k, device='cuda') / k ** 0.5
    w = w.bfloat16().float().reshape(3 * hd, k)                       # head-padded rows: pad rows are zero
    b_v = torch.zeros(heads, dp, device='cuda')
    b_v[:, :d] = torch.randn(heads, d, device='cuda') * 0.1
    bias = torch.cat([torch.zeros(2 * hd, device='cuda'), b_v.reshape(-1)])
    table = torch.zeros(800, 2, heads, dp, device='cuda')
    table[:, :, :, :d] = torch.randn(800, 2, heads, d, device='cuda') * 0.5
    table = table.reshape(800, 2 * hd).bfloat16()
    idx = torch.randint(0, 800, (m,), device='cuda', dtype=torch.int32)
    lin = WideLinear(w, bias, dp if normalize else 16, n_norm=2 * hd, normalize=normalize)
    out = lin(x, table=table, tab_idx=idx)
    assert out.shape == (m, 3 * hd) and out.dtype == torch.bfloat16
    ref = x.double() @ w.double().t()
    ref[:, :2 * hd] += table.double()[idx.long()]
    ref[:, 2 * hd:] += bias.double()[2 * hd:]
    if normalize:
        qk = F.normalize(ref[:, :2 * hd].reshape(m, 2 * heads, dp), dim=-1).reshape(m, 2 * hd)
        ref = torch.cat([qk, ref[:, 2 * hd:]], dim=1)
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-2, err                                             # bf16 output rounding: 2^-9 of the value


@pytest.mark.parametrize('m', [1, 128, 20011])
def test_wide_linear_fc1_gelu_matches_float64(m):
    """os3d_wide_linear_bf16 as MLP.fc1 + GELU of the level-4 layers (384 -> 768), point_transformer_layer.py:260-276."""
    import torch.nn.functional as F
    from openseg3d_b200.ops.linear import WideLinear
    torch.manual_seed(m)
    x = torch.randn(m, 384, device='cuda').bfloat16()
    w = (torch.randn(768, 384, device='cuda') / 384 ** 0.5).bfloat16().float()
    b = torch.randn(768, device='cuda') * 0.1
    dp = next(d for d in (16, 32) if WideLinear.fits(384, 768, 0, d))
    out = WideLinear(w, b, dp, gelu=True)(x)
    ref = F.gelu(x.double() @ w.double().t() + b.double())
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-2, err
