"""Which python lines launch the small torch kernels of a forward:  python tools/trace_ops.py
Counts aten ops that launch a CUDA kernel, grouped by the innermost openseg3d_b200 source line on the python stack."""
import collections
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
with torch.no_grad():
    for _ in range(2):
        model({'points': dev, 'batch_size': frames})
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
        model({'points': dev, 'batch_size': frames})
        torch.cuda.synchronize()
SMALL = {'aten::fill_', 'aten::copy_', 'aten::add', 'aten::add_', 'aten::mul', 'aten::mul_', 'aten::rsqrt', 'aten::neg', 'aten::cat',
         'aten::zero_', 'aten::sub', 'aten::div', 'aten::where', 'aten::sum', 'aten::index_select', 'aten::index', 'aten::gt',
         'aten::eq', 'aten::ne', 'aten::bitwise_and', 'aten::cumsum', 'aten::nonzero', 'aten::unique_consecutive',
         'aten::masked_select', 'aten::clamp', 'aten::sqrt', 'aten::addmm', 'aten::mm', 'aten::_addmm_activation'}
agg = collections.defaultdict(int)
for ev in prof.key_averages(group_by_stack_n=12):
    if ev.key not in SMALL:
        continue
    site = next((s for s in (ev.stack or []) if 'openseg3d_b200' in s), (ev.stack or ['?'])[0])
    agg[(ev.key, site.split('openseg3d_b200/')[-1][:80])] += ev.count
print('ops counted:', sum(agg.values()))
for (name, site), n in sorted(agg.items(), key=lambda kv: -kv[1])[:50]:
    print(f'{n:5d}  {name:24s} {site}')
