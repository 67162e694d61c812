"""One bench-shaped forward (8 synthetic frames, bf16) after N warm-up forwards -- the command ncu profiles.
    python tools/one_step.py [warmup=1]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from openseg3d_b200 import synthetic  # noqa: E402
from openseg3d_b200.models import build_segformer  # noqa: E402

warm = int(sys.argv[1]) if len(sys.argv) > 1 else 1
frames = 8
model = build_segformer('waymo_one_sweep', compute_dtype=torch.bfloat16).cuda().eval()
pts, _ = synthetic.make_batch(list(range(frames)), 1, False)
dev = torch.from_numpy(pts).cuda()
with torch.no_grad():
    for _ in range(warm + 1):
        out = model({'points': dev, 'batch_size': frames})
torch.cuda.synchronize()
print('ok', tuple(out['point_out'].shape))
