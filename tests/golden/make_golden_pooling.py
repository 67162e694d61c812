"""Golden vectors for voxel_avg_pooling from the REFERENCE's own CPU function (voxel_pooling.cpp:5-23), compiled from
/root/reference by oracle/build_ref.py.  Run in the build container (the reference does not exist on the GPU box):
    python tests/golden/make_golden_pooling.py      ->  tests/golden/voxel_avg_pooling.npz
Cases: ragged voxel occupancy, ids outside [0, M) (skipped by the reference: `if (pos < 0 || pos >= N1) continue`),
an empty voxel (count 0, never addressed), one voxel holding most points (long serial accumulation)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402

ext = build_ref.load()
assert ext is not None, 'needs /root/reference'
rng = np.random.default_rng(7)
out = {}
for name, (n, m, c) in {'small': (64, 9, 8), 'ragged': (1500, 200, 16), 'heavy': (3000, 20, 16)}.items():
    feats = (rng.standard_normal((n, c)) * 3).astype(np.float32)
    if name == 'heavy':
        ids = np.where(rng.random(n) < 0.8, 3, rng.integers(0, m, n)).astype(np.int32)
    else:
        ids = rng.integers(-2, m + 2, n).astype(np.int32)        # a few ids outside [0, m): skipped
    ids[ids == m - 1] = 0                                        # voxel m - 1 stays empty
    inside = (ids >= 0) & (ids < m)
    counts = np.bincount(ids[inside], minlength=m).astype(np.int32)
    ref = ext.voxel_pooling_forward_cpu(torch.from_numpy(feats), torch.from_numpy(ids), torch.from_numpy(counts)).numpy()
    out[f'{name}_feats'], out[f'{name}_ids'], out[f'{name}_counts'], out[f'{name}_out'] = feats, ids, counts, ref
np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'voxel_avg_pooling.npz'), **out)
print({k: v.shape for k, v in out.items()})
