"""Per-chunk timeline (clock64, CTA 0) of swformer_mlp_tc_kernel: python tools/exp_mlp2_trace.py C M"""
import ctypes
import sys
import torch

c, m = int(sys.argv[1]), int(sys.argv[2])
L = ctypes.CDLL('tools/_bin/libmlp2_TRACE.so')
f = L.os3d_swformer_mlp_bf16
f.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 6 + [ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
L.os3d_exp_set_trace.argtypes = [ctypes.c_void_p]
trace = torch.zeros(3 * 64 * 8, dtype=torch.int64, device='cuda')
assert L.os3d_exp_set_trace(trace.data_ptr()) == 0
h = 2 * c
x = torch.randn(m, c, device='cuda').bfloat16()
w1 = torch.zeros(((c + 63) // 64) * h * 64, dtype=torch.bfloat16, device='cuda')
w2 = torch.zeros(((h + 63) // 64) * c * 64, dtype=torch.bfloat16, device='cuda')
g, b = torch.ones(c, device='cuda'), torch.zeros(c, device='cuda')
out = torch.empty_like(x)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    assert f(x.data_ptr(), m, c, h, w1.data_ptr(), None, w2.data_ptr(), None, g.data_ptr(), b.data_ptr(), 1e-5, out.data_ptr(), st) == 0
torch.cuda.synchronize()
t = trace.cpu().reshape(3, 64, 8)
t0 = int(t[t > 0].min())
names = [['loop', 'acc1_full', 'h_free', 'ld_done', 'math_done', 'arrived', 'final_done', ''],
         ['m1_begin', 'acc1_free', 'm1_issued', 'pre_w2', 'w2_issued', 'm2_issued', 'w1_full', 'h_ready']]
for cc in range(6, 26):
    print(f'chunk {cc:2d}  EPI ' + ' '.join(f'{names[0][k]}={int(t[0, cc, k]) - t0:6d}' for k in range(7)))
    print(f'          MMA ' + ' '.join(f'{names[1][k]}={int(t[1, cc, k]) - t0:6d}' for k in (0, 1, 6, 2, 3, 4, 7, 5)))
for tt in range(1, 6):
    print(f'final {tt}: ' + ' '.join(f'{n}={int(t[2, tt, k]) - t0:6d}' for k, n in enumerate(['start', 'acc2_full', 'ld_done', 'sums', 'bar1', 'bar2', 'chunk0', 'end'])))
