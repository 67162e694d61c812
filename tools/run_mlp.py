"""Time os3d_mlp_chain_bf16 on the bench's shapes:  python tools/run_mlp.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from openseg3d_b200.ops.mlp_chain import MlpChain  # noqa: E402


def chain(widths, acts):
    return [(torch.randn(n, k, device='cuda') / k ** 0.5, torch.randn(n, device='cuda'), a)
            for (k, n), a in zip(zip(widths[:-1], widths[1:]), acts)]


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


N = 1451776
cases = [
    ('point_encoder 6->64->128->256->64', [64, 128, 256, 64], [1, 1, 0], N, True),
    ('fusion 96->256->128->64', [96, 256, 128, 64], [1, 1, 1], N, False),
    ('classifier 64->64->22', [64, 64, 22], [1, 0], N, False),
    ('swformer mlp C=48', [48, 96, 48], [2, 0], 932000, False),
    ('swformer mlp C=96', [96, 192, 96], [2, 0], 1020000, False),
]
for name, widths, acts, m, front in cases:
    layers = chain(widths, acts)
    fr = (torch.randn(64, 6, device='cuda'), torch.randn(64, device='cuda'), 1) if front else None
    c = MlpChain(layers, front=fr)
    x = torch.randn(m, 7, device='cuda')[:, 1:] if front else torch.randn(m, widths[0], device='cuda').bfloat16()
    kw = {}
    if name.startswith('swformer'):
        kw = dict(residual=x, ln=(torch.ones(widths[-1], device='cuda'), torch.zeros(widths[-1], device='cuda'), 1e-5))
    ms = timeit(lambda: c(x, **kw))
    byts = m * (x.shape[1] * x.element_size() + widths[-1] * 2 * (2 if kw else 1))
    print(f'{name:40s} m={m:8d} {ms:7.3f} ms  {c.flops_per_row * m / ms / 1e9:8.1f} TFLOP/s  {byts / ms / 1e6:7.1f} GB/s', flush=True)

from openseg3d_b200.ops.mlp_chain import SwformerMlp  # noqa: E402
for c, m in ((48, 932000), (96, 1020000), (192, 467000)):
    mlp = SwformerMlp(torch.randn(2 * c, c, device='cuda') / c ** 0.5, torch.randn(2 * c, device='cuda'),
                      torch.randn(c, 2 * c, device='cuda') / (2 * c) ** 0.5, torch.randn(c, device='cuda'))
    x = torch.randn(m, c, device='cuda').bfloat16()
    ln = (torch.ones(c, device='cuda'), torch.zeros(c, device='cuda'), 1e-5)
    ms = timeit(lambda: mlp(x, ln))
    print(f'swformer mlp (streamed weights) C={c:3d}      m={m:8d} {ms:7.3f} ms  {8.0 * m * c * c / ms / 1e9:8.1f} TFLOP/s  '
          f'{3 * m * c * 2 / ms / 1e6:7.1f} GB/s', flush=True)
