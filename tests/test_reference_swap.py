"""The drop-in boundary, exercised with the REFERENCE's own files: seg3d/models/segmentors/segformer.py,
seg3d/models/backbones/pointtransformer.py, seg3d/utils/spconv_utils.py, seg3d/models/voxel_encoders/vfe.py,
seg3d/models/layers/se_layer.py (and, with --layers 0, its point_transformer_layer.py / cosine_msa.py /
swformer_utils.py) imported unmodified from a copy of the reference package and run on openseg3d_b200 through
openseg3d_b200.compat.install() -- the alias INTEGRATION.md §A describes.  Each case runs in its own interpreter
(tools/ref_swap_check.py) because the swap edits sys.modules."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(layers, device):
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ref_swap_check.py'), '--layers', str(layers),
                        '--device', device], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    res = json.loads(p.stdout.strip().splitlines()[-1])
    if 'skipped' in res:
        pytest.skip(res['skipped'])
    return res


@pytest.mark.parametrize('layers', [0, 1])
def test_reference_segformer_state_dict_is_identical(layers):
    """CPU: the reference's Segformer built on the swapped modules has exactly our state_dict keys / shapes, and a
    strict load succeeds."""
    res = _run(layers, 'cpu')
    assert res['keys_equal'] and res['shapes_equal'] and res["n_keys"] > 250, res


@pytest.mark.gpu
@pytest.mark.parametrize('layers', [0, 1])
def test_reference_segformer_forward_matches(layers):
    """GPU: reference Segformer.forward (its own Python, our ops) == openseg3d_b200.models.Segformer on the same weights."""
    res = _run(layers, 'cuda')
    assert res['coords_equal'], res
    for k in ('rel_point_out', 'rel_voxel_out', 'rel_aux_voxel_out'):
        assert res[k] < 1e-4, res
    assert res['argmax_agreement'] > 0.999, res
