"""Time the attention in-projection of every level (8-frame sizes) on os3d_wide_linear_bf16 against the library-GEMM path
(two F.linear + os3d_add_table_rows):  python tools/run_qkv.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from openseg3d_b200 import _lib  # noqa: E402
from openseg3d_b200.ops.linear import WideLinear  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    for m, c, d in ((932119, 48, 6), (1019225, 96, 12), (467294, 192, 24), (168889, 384, 48)):
        heads, dp = 8, (d + 15) // 16 * 16
        hd = heads * dp
        x = torch.randn(m, c, device='cuda').bfloat16()
        w = (torch.randn(3 * hd, c, device='cuda') / c ** 0.5)
        bias = torch.zeros(3 * hd, device='cuda')
        table = torch.randn(800, 2 * hd, device='cuda').bfloat16()
        ptab = torch.randn(800, c, device='cuda').bfloat16()
        idx = torch.randint(0, 800, (m,), device='cuda', dtype=torch.int32)
        lin = WideLinear(w, bias, dp, n_norm=2 * hd, normalize=True)
        wq, wv = w[:2 * hd].bfloat16().contiguous(), w[2 * hd:].bfloat16().contiguous()
        bq, bv = bias[:2 * hd].bfloat16(), bias[2 * hd:].bfloat16()
        xp = torch.empty_like(x)

        def lib_path():
            _lib.call('os3d_add_table_rows', x, ptab, idx, m, c, 2, xp)
            F.linear(xp, wq, bq)
            F.linear(x, wv, bv)

        t_own = timed(lambda: lin(x, table=table, tab_idx=idx), reps)
        t_lib = timed(lib_path, reps)
        nbytes = 2.0 * m * (c + 3 * hd)
        print(f'C={c} m={m} n={3 * hd} nc={lin.nc}: wide_linear {t_own:.3f} ms ({nbytes / t_own / 1e6:.0f} GB/s)  '
              f'library path {t_lib:.3f} ms ({nbytes / t_lib / 1e6:.0f} GB/s)', flush=True)
    m = 168889
    x = torch.randn(m, 384, device='cuda').bfloat16()
    w = torch.randn(768, 384, device='cuda') / 20
    b = torch.randn(768, device='cuda')
    lin = WideLinear(w, b, next(d for d in (16, 32) if WideLinear.fits(384, 768, 0, d)), gelu=True)
    wb, bb = w.bfloat16(), b.bfloat16()

    def lib_fc1():
        h = F.linear(x, wb, bb)
        _lib.call('os3d_gelu_bf16', h, h.numel(), h)

    print(f'fc1+GELU 384->768 m={m}: wide_linear {timed(lambda: lin(x), reps):.3f} ms  library + gelu kernel {timed(lib_fc1, reps):.3f} ms')


if __name__ == '__main__':
    main()
