"""Time os3d_swformer_mlp_bf16 at the three level sizes of the 8-frame bench batch:  python tools/mlp_bench.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from openseg3d_b200.ops.mlp_chain import SwformerMlp  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for c, m in ((48, 932119), (96, 1019225), (192, 467294)):
    torch.manual_seed(c)
    mlp = SwformerMlp(torch.randn(2 * c, c, device='cuda') / c ** 0.5, torch.randn(2 * c, device='cuda'),
                      torch.randn(c, 2 * c, device='cuda') / (2 * c) ** 0.5, torch.randn(c, device='cuda'))
    x = torch.randn(m, c, device='cuda').bfloat16()
    ln = (torch.ones(c, device='cuda'), torch.zeros(c, device='cuda'), 1e-5)
    for _ in range(3):
        y = mlp(x, ln)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        y = mlp(x, ln)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = (3 * m * c * 2) / 1e9
    print(f'C={c:4d} M={m:8d}  {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s algorithmic  {8.0 * m * c * c / ms / 1e9:.0f} TFLOP/s')
