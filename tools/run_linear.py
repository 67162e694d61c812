"""Time os3d_linear_bf16 on bench-shaped problems:  python tools/run_linear.py M K N MODE [max_width]
MODE: plain | gelu | ln | table"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from openseg3d_b200.ops.linear import GELU, PackedLinearCache, linear_bf16  # noqa: E402

m, k, n, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
mw = int(sys.argv[5]) if len(sys.argv) > 5 else (512 if mode == 'ln' else 256)
x = torch.randn(m, k, device='cuda').bfloat16()
w, b = torch.randn(n, k, device='cuda') * 0.05, torch.randn(n, device='cuda')
chunks = PackedLinearCache().get('w', w, b, max_width=mw)
kw = {}
if mode == 'gelu':
    kw['flags'] = GELU
elif mode == 'ln':
    kw.update(ln=(torch.ones(n, device='cuda'), torch.zeros(n, device='cuda'), 1e-5),
              residual=torch.randn(m, n, device='cuda').bfloat16())
elif mode == 'table':
    kw.update(table=[torch.randn(800, c[3], device='cuda').bfloat16() for c in chunks],
              tab_idx=torch.randint(0, 800, (m,), device='cuda', dtype=torch.int32))
for _ in range(3):
    linear_bf16(x, chunks, **kw)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    linear_bf16(x, chunks, **kw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
byt = 2.0 * m * (k + n * (2 if mode == 'ln' else 1))
print(f'M={m} K={k} N={n} {mode} width<={mw}: {ms:.3f} ms, {byt / ms / 1e6:.0f} GB/s algorithmic, {2.0 * m * k * n / ms / 1e9:.0f} TFLOP/s')
