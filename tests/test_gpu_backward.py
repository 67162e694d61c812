"""Backward of the hot path (BASELINE config 5): gradients of the CUDA ops against torch autograd through the CPU
oracle's restatement of the same op, on the same operands.  fp32: rel 1e-4; bf16: rel 2e-2 (north star)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sites(rng, batch, shape, n):
    z, y, x = shape
    lin = rng.choice(batch * z * y * x, size=n, replace=False)
    b, r = np.divmod(lin, z * y * x)
    zz, r = np.divmod(r, y * x)
    yy, xx = np.divmod(r, x)
    return np.stack([b, zz, yy, xx], axis=1).astype(np.int32)


def _rel(a, b):
    return float((a.double().cpu() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


@pytest.mark.parametrize('kind,cin,cout,dtype', [('subm', 16, 32, torch.float32), ('subm', 48, 48, torch.bfloat16),
                                                ('strided', 32, 48, torch.float32), ('strided', 48, 96, torch.bfloat16),
                                                ('inverse', 48, 32, torch.float32), ('inverse', 96, 48, torch.bfloat16),
                                                ('subm', 40, 48, torch.bfloat16)])
def test_sparse_conv_backward_matches_oracle_autograd(kind, cin, cout, dtype):
    from openseg3d_b200 import spconv
    from oracle import oracle
    rng = np.random.default_rng(cin + cout)
    torch.manual_seed(cin * 7 + cout)
    shape = (10, 30, 30)
    idx = _sites(rng, 2, shape, 2500)
    bf16 = dtype == torch.bfloat16
    rnd = (lambda t: t.bfloat16().float()) if bf16 else (lambda t: t)
    x = spconv.SparseConvTensor(torch.zeros(idx.shape[0], 1).cuda(), torch.from_numpy(idx).cuda(), shape, 2)
    if kind == 'subm':
        conv = spconv.SubMConv3d(cin, cout, 3, padding=1, bias=True, indice_key='k').cuda()
        nbr, _ = oracle.subm_map(idx, shape)
        m_in = idx.shape[0]
    else:
        down = spconv.SparseConv3d(8, 8, 3, stride=2, padding=1, bias=False, indice_key='s').cuda()
        with torch.no_grad():
            y0 = down(x.replace_feature(torch.zeros(idx.shape[0], 8).cuda()))          # builds the strided rulebook
        _, _, fwd, inv, _ = oracle.strided_map(idx, shape)
        if kind == 'strided':
            conv = spconv.SparseConv3d(cin, cout, 3, stride=2, padding=1, bias=False, indice_key='s').cuda()
            nbr, m_in = fwd, idx.shape[0]
        else:
            conv = spconv.SparseInverseConv3d(cin, cout, 3, bias=False, indice_key='s').cuda()
            nbr, m_in = inv, y0.features.shape[0]
            x = y0
    with torch.no_grad():
        conv.weight.copy_(rnd(conv.weight))
    feats = rnd(torch.randn(m_in, cin))
    f_gpu = feats.to(dtype).cuda().requires_grad_(True)
    out = conv(x.replace_feature(f_gpu)).features
    gout = rnd(torch.randn(out.shape))
    out.backward(gout.to(dtype).cuda())

    f_ref = feats.double().requires_grad_(True)
    w_ref = conv.weight.detach().cpu().double().requires_grad_(True)
    b_ref = conv.bias.detach().cpu().double().requires_grad_(True) if conv.bias is not None else None
    ref = oracle.sparse_conv(f_ref, nbr, w_ref, b_ref)
    ref.backward(gout.double())
    tol = 2e-2 if bf16 else 1e-4
    assert _rel(out.detach().float(), ref.detach()) < tol
    assert _rel(f_gpu.grad.float(), f_ref.grad) < tol, _rel(f_gpu.grad.float(), f_ref.grad)
    assert _rel(conv.weight.grad, w_ref.grad) < tol, _rel(conv.weight.grad, w_ref.grad)
    if b_ref is not None:
        assert _rel(conv.bias.grad, b_ref.grad) < tol


@pytest.mark.parametrize('c,heads,dtype', [(24, 2, torch.float32), (48, 8, torch.float32), (96, 8, torch.bfloat16),
                                           (192, 8, torch.float32)])
def test_window_attention_backward_matches_oracle_autograd(c, heads, dtype):
    """Gradients of WindowAttention (projections, normalisation, attention core, temperature) against torch autograd
    through the oracle's padded-window restatement of cosine_msa.py on the same parameters."""
    from openseg3d_b200 import spconv
    from openseg3d_b200.models import SparseWindowPartitionLayer, WindowAttention
    from openseg3d_b200.models.segmentors import default_batching_info
    from oracle import oracle
    rng = np.random.default_rng(c)
    torch.manual_seed(c)
    shape = (8, 40, 40)                                       # z, y, x
    idx = _sites(rng, 2, shape, 1500)
    binfo = default_batching_info()[0]
    layer = SparseWindowPartitionLayer(binfo, (10, 10, 8), (shape[2], shape[1], shape[0]))
    attn = WindowAttention(c, heads, 0.0).cuda()
    bf16 = dtype == torch.bfloat16
    rnd = (lambda t: t.bfloat16().float()) if bf16 else (lambda t: t)
    with torch.no_grad():
        attn.self_attn.tau.fill_(0.4)
        for p in attn.parameters():
            p.copy_(rnd(p))
    feats = rnd(torch.randn(idx.shape[0], c))
    x = spconv.SparseConvTensor(feats.to(dtype).cuda(), torch.from_numpy(idx).cuda(), shape, 2)
    info = layer(x)
    f_gpu = feats.to(dtype).cuda().requires_grad_(True)
    out = attn(f_gpu, info['pos_dict_shift0'], info['flat2win_inds_shift0'], info['key_mask_shift0'])
    gout = rnd(torch.randn(out.shape))
    out.backward(gout.to(dtype).cuda())

    # oracle: padded windows per batching level, torch ops in float64 on the CPU
    sd = {k: v.detach().cpu().double().requires_grad_(True) for k, v in attn.self_attn.state_dict(keep_vars=True).items()}
    part = oracle.partition(idx, (shape[2], shape[1], shape[0]), (10, 10, 8), binfo, c)
    f_ref = feats.double().requires_grad_(True)
    shift = part[0]
    x3 = oracle.flat2window(f_ref, shift['inds'], binfo)
    p3 = oracle.flat2window(torch.as_tensor(shift['pos_flat']).double(), shift['inds'], binfo)
    o3 = {lvl: oracle.cosine_attention(x3[lvl], p3[lvl], shift['mask'][lvl], sd, heads) for lvl in x3}
    ref = oracle.window2flat(o3, shift['inds'], idx.shape[0])
    ref.backward(gout.double())
    tol = 3e-2 if bf16 else 2e-4
    assert _rel(out.detach().float(), ref.detach()) < tol
    assert _rel(f_gpu.grad.float(), f_ref.grad) < tol, _rel(f_gpu.grad.float(), f_ref.grad)
    for name in ('in_proj_weight', 'in_proj_bias', 'out_proj.weight', 'out_proj.bias', 'tau'):
        g = dict(attn.self_attn.named_parameters())[name].grad
        # tau: one scalar = a cancelling sum over every (query, key) pair, accumulated with float atomics
        assert _rel(g, sd[name].grad) < (10 * tol if name == 'tau' else tol), (name, _rel(g, sd[name].grad))


def test_segformer_gradients_match_oracle_autograd():
    """Whole hot path, fp32, BatchNorm in eval mode / no dropout: d loss / d parameters of the
    CUDA model against torch autograd through the oracle's forward (float64) over the same state_dict."""
    from openseg3d_b200 import synthetic
    from openseg3d_b200.models import build_segformer
    from openseg3d_b200.models.segmentors import default_batching_info, DATASET_CONFIGS
    from oracle import oracle
    depths = (1, 2, 1, 1)
    model = build_segformer('waymo_one_sweep', depths=depths).cuda().eval()
    with torch.no_grad():
        g = torch.Generator().manual_seed(1)
        for n, b in model.named_buffers():
            if n.endswith('running_mean'):
                b.copy_(0.1 * torch.randn(b.shape, generator=g))
            elif n.endswith('running_var'):
                b.copy_(torch.empty(b.shape).uniform_(0.5, 1.5, generator=g))
        for n, p in model.named_parameters():
            if n.endswith('tau'):
                p.fill_(0.3)
    pts, _ = synthetic.make_batch([0], 1, False, 16, 200)
    res = model({'points': torch.from_numpy(pts).cuda(), 'batch_size': 1})
    torch.manual_seed(0)
    w_p, w_v, w_a = (torch.randn(res[k].shape) for k in ('point_out', 'voxel_out', 'aux_voxel_out'))
    loss = (res['point_out'] * w_p.cuda()).sum() + (res['voxel_out'] * w_v.cuda()).sum() + (res['aux_voxel_out'] * w_a.cuda()).sum()
    loss.backward()

    params = {n for n, _ in model.named_parameters()}
    sd = {k: (v.detach().cpu().double().requires_grad_(k in params) if v.dtype.is_floating_point else v.cpu())
          for k, v in model.state_dict().items()}
    c = DATASET_CONFIGS['waymo_one_sweep']
    ref = oracle.segformer_forward(sd, pts, c['voxel_size'], c['point_cloud_range'], default_batching_info(), [10, 10, 8],
                                   list(depths), differentiable=True)
    ref_loss = (ref['point_out'] * w_p.double()).sum() + (ref['voxel_out'] * w_v.double()).sum() + \
        (ref['aux_voxel_out'] * w_a.double()).sum()
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 1e-3 * abs(ref_loss.item())
    errs = []
    for name, p in model.named_parameters():
        rg = sd[name].grad
        if rg is None:
            continue
        assert p.grad is not None, name
        errs.append((_rel(p.grad, rg), name))
    errs.sort(reverse=True)
    print('parameters checked', len(errs), 'worst:', errs[:8])
    assert len(errs) > 100, len(errs)
    # fp32 end to end through ~60 layers: 1e-4 per op (north star) accumulates; the bound here is on the whole chain
    assert errs[0][0] < 1e-2, errs[:5]
    assert sum(e for e, _ in errs) / len(errs) < 1e-3


def test_bf16_training_step_runs_and_learns():
    """BASELINE config 5 in miniature: train() mode (batch-stat BatchNorm, attention dropout, DropPath), bf16 backbone,
    cross-entropy on synthetic labels, SGD: the loss of a fixed batch goes down over a few steps."""
    import torch.nn.functional as F
    from openseg3d_b200 import synthetic
    from openseg3d_b200.models import build_segformer
    torch.manual_seed(0)
    model = build_segformer('waymo_one_sweep', compute_dtype=torch.bfloat16, depths=(1, 1, 1, 1)).cuda().train()
    opt = torch.optim.SGD(model.parameters(), lr=0.02, momentum=0.9)
    pts, _ = synthetic.make_batch([0, 1], 1, False, 16, 300)
    dev = torch.from_numpy(pts).cuda()
    labels = torch.randint(0, 22, (pts.shape[0],), device='cuda')
    losses = []
    for _ in range(6):
        out = model({'points': dev, 'batch_size': 2})
        loss = F.cross_entropy(out['point_out'].float(), labels)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0], losses
