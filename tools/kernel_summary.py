"""Kernel-time summary of one bench step with torch.profiler (CUPTI): cheaper than an ncu launch list, used while
iterating.  The judged evidence stays the ncu files under profiles/."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from openseg3d_b200 import synthetic  # noqa: E402
from openseg3d_b200.models import build_segformer  # noqa: E402

frames = 8
model = build_segformer('waymo_one_sweep', compute_dtype=torch.bfloat16).cuda().eval()
pts, _ = synthetic.make_batch(list(range(frames)), 1, False)
dev = torch.from_numpy(pts).cuda()
for _ in range(3):
    with torch.no_grad():
        model({'points': dev, 'batch_size': frames})
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    with torch.no_grad():
        model({'points': dev, 'batch_size': frames})
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == 'CUDA']
tot = sum(r[1] for r in rows)
print(f'total kernel time {tot:.2f} ms')
for k, ms, n in sorted(rows, key=lambda r: -r[1])[:45]:
    print(f'{ms:9.3f} ms {100 * ms / tot:5.1f}% n={n:4d}  {k[:110]}')
