"""os3d_mlp_chain_bf16 (a chain of Linear layers in one tcgen05 kernel, activations on chip) against a float64 torch
restatement of the same chain -- nn.Linear / ReLU / nn.GELU / residual + nn.LayerNorm -- with the hidden activations
rounded to bf16 where the kernel rounds them (segformer.py:21-32,58-76; point_transformer_layer.py:260-298).
Tolerance: bf16 rel 2e-2 of the output scale (north star)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ACT = {0: lambda t: t, 1: lambda t: t.clamp(min=0), 2: F.gelu}


def _chain(widths, acts, seed, bias=True):
    torch.manual_seed(seed)
    layers = []
    for (k, n), a in zip(zip(widths[:-1], widths[1:]), acts):
        w = (torch.randn(n, k) * np.sqrt(2.0 / k)).bfloat16().float()
        b = torch.randn(n) * 0.5 if bias else None
        layers.append((w, b, a))
    return layers


def _ref(x, layers, front=None, residual=None, ln=None):
    a = x.double()
    if front is not None:
        w, b, act = front
        a = ACT[act](a @ w.double().t() + (b.double() if b is not None else 0.0)).float().bfloat16().double()
    for i, (w, b, act) in enumerate(layers):
        a = a @ w.double().t() + (b.double() if b is not None else 0.0)
        if i + 1 < len(layers):
            a = ACT[act](a).float().bfloat16().double()
    if ln is not None:
        a = F.layer_norm(a, (a.shape[1],), ln[0].double(), ln[1].double(), ln[2])
    if residual is not None:
        a = a + residual.double()
    return ACT[layers[-1][2]](a)


def _check(got, ref, tol=2e-2):
    got = got.float().cpu().double()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    assert err <= tol * max(ref.abs().max().item(), 1e-6), (err, ref.abs().max().item())


def _cuda(layers):
    return [(w.cuda(), None if b is None else b.cuda(), a) for w, b, a in layers]


@pytest.mark.parametrize('m', [1, 127, 1000, 300001])
def test_point_encoder_chain(m):
    """fp32 front layer (raw coordinates, row pitch 7 like batch_dict['points'][:, 1:]) + 64 -> 128 -> 256 -> 64."""
    from openseg3d_b200.ops.mlp_chain import MlpChain
    layers = _chain([64, 128, 256, 64], [1, 1, 0], 1)
    torch.manual_seed(2)
    front = (torch.randn(64, 6) * 0.3, torch.randn(64) * 0.5, 1)
    raw = torch.randn(m, 7) * torch.tensor([1.0, 30.0, 30.0, 2.0, 1.0, 1.0, 1.0])
    chain = MlpChain(_cuda(layers), front=tuple(t.cuda() if torch.is_tensor(t) else t for t in front))
    y = chain(raw.cuda()[:, 1:])
    assert y.shape == (m, 64) and y.dtype == torch.bfloat16
    _check(y, _ref(raw[:, 1:], layers, front=front))


@pytest.mark.parametrize('m', [5, 4096, 200000])
def test_fusion_chain_direct_mode(m):
    """96 -> 256 -> 128 -> 64, ReLU after every layer: the widest chain (x tiles land in the activation buffer)."""
    from openseg3d_b200.ops.mlp_chain import MlpChain
    layers = _chain([96, 256, 128, 64], [1, 1, 1], 3)
    x = torch.randn(m, 96).bfloat16()
    y = MlpChain(_cuda(layers))(x.cuda())
    _check(y, _ref(x, layers))


@pytest.mark.parametrize('m', [3, 777, 150000])
def test_classifier_chain_ragged_output(m):
    """64 -> 64 (ReLU) -> 22 classes: output rows of 22 bf16 (44 bytes), weight rows zero-padded to 32."""
    from openseg3d_b200.ops.mlp_chain import MlpChain
    layers = _chain([64, 64, 22], [1, 0], 4, bias=False)
    x = torch.randn(m, 64).bfloat16()
    y = MlpChain(_cuda(layers))(x.cuda())
    assert y.shape == (m, 22)
    _check(y, _ref(x, layers))
    y32 = MlpChain(_cuda(layers))(x.cuda(), out_dtype=torch.float32)
    assert y32.dtype == torch.float32
    _check(y32, _ref(x, layers))


@pytest.mark.parametrize('c,m', [(48, 1000), (96, 129), (96, 250000), (48, 250000)])
def test_swformer_mlp_chain(c, m):
    """x + LayerNorm(fc2(GELU(fc1(x)))) (EncoderLayer, point_transformer_layer.py:278-298)."""
    from openseg3d_b200.ops.mlp_chain import MlpChain
    layers = _chain([c, 2 * c, c], [2, 0], 5)
    torch.manual_seed(6)
    x = torch.randn(m, c).bfloat16()
    gamma, beta = torch.rand(c) + 0.5, torch.randn(c)
    y = MlpChain(_cuda(layers))(x.cuda(), residual=x.cuda(), ln=(gamma.cuda(), beta.cuda(), 1e-5))
    _check(y, _ref(x, layers, residual=x, ln=(gamma, beta, 1e-5)))


def test_strided_input_and_output_views():
    """Input and output as column slices of wider row-major buffers (the concat feeding fusion_encoder is written in place)."""
    from openseg3d_b200.ops.mlp_chain import MlpChain
    layers = _chain([64, 128, 64], [1, 1], 7)
    m = 5000
    wide_in = torch.randn(m, 96).bfloat16().cuda()
    wide_out = torch.zeros(m, 96, dtype=torch.bfloat16, device='cuda')
    MlpChain(_cuda(layers))(wide_in[:, :64], out=wide_out[:, 32:])
    _check(wide_out[:, 32:], _ref(wide_in[:, :64].cpu(), layers))
    assert (wide_out[:, :32] == 0).all()


def test_rejects_what_does_not_fit():
    from openseg3d_b200.ops.mlp_chain import MlpChain
    assert not MlpChain.fits([(384, 192), (192, 384)])          # weights beyond shared memory
    assert not MlpChain.fits([(64, 64)])                         # a single layer is os3d_linear_tc_bf16's job
    with pytest.raises(RuntimeError):
        MlpChain(_cuda(_chain([192, 384, 192], [2, 0], 8)))
    with pytest.raises(RuntimeError):
        MlpChain(_cuda(_chain([64, 64, 64], [1, 0], 9)))(torch.randn(4, 64).cuda())      # fp32 without a front layer


@pytest.mark.parametrize('c,m', [(48, 1), (48, 1000), (96, 129), (192, 4000), (192, 300000), (96, 250000), (48, 250000),
                                 (64, 33000), (128, 70000)])
def test_swformer_mlp_streamed_weights(c, m):
    """os3d_swformer_mlp_bf16: the same operation with the weights streamed in hidden-dimension chunks (C up to 192)."""
    from openseg3d_b200.ops.mlp_chain import SwformerMlp
    layers = _chain([c, 2 * c, c], [2, 0], 11)
    torch.manual_seed(12)
    x = (torch.randn(m, c) * 1.5).bfloat16()
    gamma, beta = torch.rand(c) + 0.5, torch.randn(c)
    (w1, b1, _), (w2, b2, _) = layers
    mlp = SwformerMlp(w1.cuda(), b1.cuda(), w2.cuda(), b2.cuda())
    y = mlp(x.cuda(), (gamma.cuda(), beta.cuda(), 1e-5))
    assert y.shape == (m, c) and y.dtype == torch.bfloat16
    _check(y, _ref(x, layers, residual=x, ln=(gamma, beta, 1e-5)))


def test_swformer_mlp_rejects_wide_layers():
    from openseg3d_b200.ops.mlp_chain import SwformerMlp
    assert not SwformerMlp.fits(384, 768)
    assert SwformerMlp.fits(192, 384)
    with pytest.raises(RuntimeError):
        SwformerMlp(torch.zeros(768, 384).cuda(), None, torch.zeros(384, 768).cuda(), None)
