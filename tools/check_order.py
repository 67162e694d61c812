"""Offsets per 128-row tile of the level-L submanifold map in storage order vs mask-grouped order, and the cost of
building the order:  python tools/check_order.py LEVEL"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from openseg3d_b200 import _lib, spconv, synthetic  # noqa: E402
from openseg3d_b200.core import voxelize_batch  # noqa: E402

level = int(sys.argv[1])
frames = 8
pts, _ = synthetic.make_batch(list(range(frames)), 1, False)
coors, _ = voxelize_batch(torch.from_numpy(pts).cuda(), [0.1, 0.1, 0.1], [-72, -72, -2, 72, 72, 4.4])
x = spconv.SparseConvTensor(torch.zeros(coors.shape[0], 1, device='cuda'), coors, [64, 1440, 1440], frames)
for _ in range(level - 1):
    rb = spconv.build_strided_rulebook(x)
    x = spconv.SparseConvTensor(torch.zeros(rb.out_indices.shape[0], 1, device='cuda'), rb.out_indices, rb.out_shape, frames)
nbr = spconv.build_subm_rulebook(x).nbr
m = nbr.shape[0]
n_tiles = (m + 127) // 128


def popc(t):
    t = t.long() & 0x7ffffff
    c = torch.zeros_like(t)
    for b in range(27):
        c += (t >> b) & 1
    return c.float().mean().item()


nbr_t = torch.empty((27, n_tiles * 128), dtype=torch.int32, device='cuda')
tm = torch.empty(n_tiles, dtype=torch.int32, device='cuda')
_lib.call('os3d_kernel_map_tiles', nbr, m, None, nbr_t, tm)
print('storage order: offsets per tile', popc(tm))
import ctypes  # noqa: E402
scratch = torch.empty((3, m), dtype=torch.int32, device='cuda')
perm = torch.empty(m, dtype=torch.int32, device='cuda')
nbytes = ctypes.c_int64(0)
_lib.lib().os3d_kernel_map_order_scratch(m, ctypes.byref(nbytes))
temp = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device='cuda')
args = (nbr, m, scratch[0], scratch[1], scratch[2], perm, temp, nbytes.value)
for _ in range(2):
    _lib.call('os3d_kernel_map_order', *args)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
_lib.call('os3d_kernel_map_order', *args)
e1.record()
torch.cuda.synchronize()
print('order build ms', e0.elapsed_time(e1), 'is permutation', bool((torch.sort(perm.long()).values == torch.arange(m, device='cuda')).all()))
_lib.call('os3d_kernel_map_tiles', nbr, m, perm, nbr_t, tm)
print('grouped order: offsets per tile', popc(tm))
