"""Kernel-time summary of one bf16 training step (torch.profiler): python tools/train_kernel_summary.py [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from openseg3d_b200 import synthetic  # noqa: E402
from openseg3d_b200.models import build_segformer  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2
model = build_segformer('waymo_one_sweep', compute_dtype=torch.bfloat16).cuda().train()
opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
pts, _ = synthetic.make_batch(list(range(frames)), 1, False)
dev = torch.from_numpy(pts).cuda()
labels = torch.randint(0, 22, (pts.shape[0],), device='cuda')


def step():
    out = model({'points': dev, 'batch_size': frames})
    loss = F.cross_entropy(out['point_out'].float(), labels)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == 'CUDA']
tot = sum(r[1] for r in rows)
print(f'total kernel time {tot:.2f} ms')
for k, ms, n in sorted(rows, key=lambda r: -r[1])[:25]:
    print(f'{ms:9.3f} ms {100 * ms / tot:5.1f}% n={n:4d}  {k[:110]}')
