"""One os3d_swformer_mlp_bf16 launch (for ncu):  python tools/run_mlp2.py C M [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from openseg3d_b200.ops.mlp_chain import SwformerMlp  # noqa: E402

c, m = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mlp = SwformerMlp(torch.randn(2 * c, c, device='cuda') / c ** 0.5, torch.randn(2 * c, device='cuda'),
                  torch.randn(c, 2 * c, device='cuda') / (2 * c) ** 0.5, torch.randn(c, device='cuda'))
x = torch.randn(m, c, device='cuda').bfloat16()
ln = (torch.ones(c, device='cuda'), torch.zeros(c, device='cuda'), 1e-5)
for _ in range(reps):
    y = mlp(x, ln)
torch.cuda.synchronize()
print('ok', float(y.float().abs().mean()))
