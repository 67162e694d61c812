"""Timing-only experiments on swformer_mlp_tc_kernel variants (tools/_bin/libmlp2_*.so, zero weights)."""
import ctypes
import sys
import torch

for name in sys.argv[1:]:
    L = ctypes.CDLL(f'tools/_bin/libmlp2_{name}.so')
    f = L.os3d_swformer_mlp_bf16
    f.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 6 + [ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
    for c, m in ((48, 932000), (96, 1020000), (192, 467000)):
        h = 2 * c
        x = torch.randn(m, c, device='cuda').bfloat16()
        w1 = torch.zeros(((c + 63) // 64) * h * 64, dtype=torch.bfloat16, device='cuda')
        w2 = torch.zeros(((h + 63) // 64) * c * 64, dtype=torch.bfloat16, device='cuda')
        g, b = torch.ones(c, device='cuda'), torch.zeros(c, device='cuda')
        out = torch.empty_like(x)
        st = torch.cuda.current_stream().cuda_stream
        call = lambda: f(x.data_ptr(), m, c, h, w1.data_ptr(), None, w2.data_ptr(), None, g.data_ptr(), b.data_ptr(), 1e-5, out.data_ptr(), st)
        for _ in range(3):
            assert call() == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            call()
        e1.record()
        torch.cuda.synchronize()
        print(f'{name:8s} C={c:3d} {e0.elapsed_time(e1) / 10:.3f} ms', flush=True)
