"""Whole-model parity at BASELINE size: ONE full synthetic frame (seed 0, 64 beams x 2650 columns, ~181k points; 3 sweeps
~544k) through Segformer with the default depths [3, 4, 8, 3], CUDA path vs ``oracle.segformer_forward`` over the same
state_dict, for the three dataset configs.

Tolerances are the north star's (BASELINE.json): fp32 rel 1e-4 and identical argmax labels > 99.9 %; bf16 rel 2e-2.
rel = max |a - b| / max |b| over a tensor (max-norm), the same definition for every stage and output.

bf16 argmax agreement: on RANDOM-INIT weights the fp32 logit margins are continuous down to zero (0.1 % of the points of
this frame have a top-2 margin below 3e-5 of the largest logit), so NO bf16 arithmetic can give > 99.9 % identical labels
here -- rounding only the WEIGHTS to bf16 and keeping every activation and accumulation in fp32 (the oracle itself, run
twice) already flips more labels than that.  The test therefore pins the bf16 path three ways: (1) rel < 2e-2 on
every output, (2) labels identical wherever the oracle's own top-2 margin exceeds the stated tolerance band
(2 * 2e-2 * max |logit|) -- the labels a rel-2e-2 perturbation can not legitimately change, (3) agreement not worse than
the weights-only-bf16 floor minus a small slack.  Every number is written to gpurun_out/parity_full_frame.json (copied
to profiles/ by the builder) together with the per-stage error table.
"""
import functools
import json
import os

import numpy as np
import pytest
import torch

from openseg3d_b200 import synthetic

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEPTHS = (3, 4, 8, 3)
CASES = [('waymo_one_sweep', 1, False), ('waymo_one_sweep_cylinder', 1, True), ('waymo_multi_sweeps', 3, False)]
STAGES = ['enc1', 'enc2', 'enc3', 'enc4', 'up4', 'up3', 'up2', 'up1']


def _model(config, dtype):
    from openseg3d_b200.models import build_segformer
    model = build_segformer(config, compute_dtype=dtype, depths=DEPTHS).eval()
    with torch.no_grad():                      # non-trivial BatchNorm statistics and temperatures
        g = torch.Generator().manual_seed(1)
        for n, b in model.named_buffers():
            if n.endswith('running_mean'):
                b.copy_(0.1 * torch.randn(b.shape, generator=g))
            elif n.endswith('running_var'):
                b.copy_(torch.empty(b.shape).uniform_(0.5, 1.5, generator=g))
        for n, p in model.named_parameters():
            if n.endswith('tau'):
                p.fill_(0.3)
    return model


@functools.lru_cache(maxsize=None)
def _frame(config):
    sweeps, cyl = next((s, c) for n, s, c in CASES if n == config)
    return synthetic.make_batch([0], sweeps, cyl)[0]


@functools.lru_cache(maxsize=None)
def _oracle(config, bf16_weights=False):
    """The oracle's forward on the frame (cached: fp32 and bf16 tests share it).  bf16_weights: every floating-point
    parameter rounded to bf16 first, arithmetic still fp32 -- the perturbation floor of ANY bf16 implementation."""
    from openseg3d_b200.models.segmentors import default_batching_info, DATASET_CONFIGS
    from oracle import oracle
    sd = _model(config, torch.float32).state_dict()
    if bf16_weights:
        sd = {k: (v.bfloat16().float() if v.dtype.is_floating_point and not k.endswith(('running_mean', 'running_var', 'tau'))
                  else v) for k, v in sd.items()}
    c = DATASET_CONFIGS[config]
    stats = {}
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        ref = oracle.segformer_forward(sd, _frame(config), c['voxel_size'], c['point_cloud_range'], default_batching_info(),
                                       [10, 10, 8], list(DEPTHS), multi_sweeps=c['use_multi_sweeps'], stats=stats)
    ref['stages'] = stats['stages']
    return ref


def _gpu(config, dtype):
    model = _model(config, dtype).cuda()
    stages = {}
    pt = model.point_transformer
    hooks = []
    for i in range(4):
        hooks.append(getattr(pt, f'swformer_block{i + 1}')[1].register_forward_hook(
            lambda m, a, out, k=f'enc{i + 1}': stages.__setitem__(k, out.detach().float().cpu())))
        hooks.append(getattr(pt, f'up{i + 1}').register_forward_hook(
            lambda m, a, out, k=f'up{i + 1}': stages.__setitem__(k, out.features.detach().float().cpu())))
    batch = {'points': torch.from_numpy(_frame(config)).cuda(), 'batch_size': 1}
    with torch.no_grad():
        res = model(batch)
    torch.cuda.synchronize()
    for h in hooks:
        h.remove()
    return res, stages, batch


def _rel(a, b):
    a, b = a.float().cpu().double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def _record(config, dtype, table):
    out_dir = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, 'parity_full_frame.json')
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[f'{config}/{dtype}'] = table
    json.dump(data, open(path, 'w'), indent=1, sort_keys=True)


def _compare(config, dtype):
    res, stages, batch = _gpu(config, dtype)
    ref = _oracle(config)
    assert np.array_equal(batch['voxel_coords'].cpu().numpy(), ref['voxel_coords'])
    assert np.array_equal(batch['point_voxel_ids'].cpu().numpy(), ref['point_voxel_ids'].numpy())
    assert np.array_equal(res['aux_voxel_coords'].cpu().numpy(), ref['aux_voxel_coords'])
    table = {'points': int(batch['points'].shape[0]), 'voxels': int(ref['voxel_coords'].shape[0]), 'depths': list(DEPTHS)}
    table['stage_rel'] = {k: _rel(stages[k], ref['stages'][k]) for k in STAGES}
    table['output_rel'] = {k: _rel(res[k], ref[k]) for k in ('voxel_out', 'aux_voxel_out', 'point_out')}
    lab, lab_ref = res['point_out'].float().argmax(1).cpu(), ref['point_out'].argmax(1)
    table['argmax_agreement'] = float((lab == lab_ref).float().mean())
    top2 = ref['point_out'].topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    table['max_abs_logit'] = float(ref['point_out'].abs().max())
    table['margin_quantiles'] = {q: float(np.quantile(margin.numpy(), float(q))) for q in ('0.001', '0.01', '0.05', '0.5')}
    return res, ref, table, lab, lab_ref, margin


@pytest.mark.parametrize('config,sweeps,cyl', CASES)
def test_full_frame_fp32(config, sweeps, cyl):
    res, ref, table, lab, lab_ref, margin = _compare(config, torch.float32)
    _record(config, 'fp32', table)
    for k, v in {**table['stage_rel'], **table['output_rel']}.items():
        assert v < 1e-4, (k, v)
    assert table['argmax_agreement'] > 0.999, table['argmax_agreement']


@pytest.mark.parametrize('config,sweeps,cyl', CASES)
def test_full_frame_bf16(config, sweeps, cyl):
    res, ref, table, lab, lab_ref, margin = _compare(config, torch.bfloat16)
    tol = 2e-2
    band = 2 * tol * table['max_abs_logit']                 # two logits may each move by tol * max |logit|
    safe = margin > band
    table['safe_fraction'] = float(safe.float().mean())
    table['argmax_agreement_outside_band'] = float((lab[safe] == lab_ref[safe]).float().mean()) if bool(safe.any()) else 1.0
    if config == 'waymo_one_sweep':                         # the floor any bf16 path has: weights rounded, arithmetic fp32
        floor = _oracle(config, True)
        table['weights_only_bf16_floor'] = {
            'argmax_agreement': float((floor['point_out'].argmax(1) == lab_ref).float().mean()),
            'point_out_rel': _rel(floor['point_out'], ref['point_out'])}
    _record(config, 'bf16', table)
    for k, v in table['output_rel'].items():
        assert v < tol, (k, v)
    for k, v in table['stage_rel'].items():
        assert v < 2 * tol, (k, v)                          # internal stages: reported; gate loosely
    assert table['argmax_agreement_outside_band'] == 1.0, table
    if 'weights_only_bf16_floor' in table:
        assert table['argmax_agreement'] > table['weights_only_bf16_floor']['argmax_agreement'] - 0.03, table
