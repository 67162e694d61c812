"""N > 1 host logic on CPU: two gloo ranks shard frames and aggregate units / time the way bench.py does on NCCL."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from openseg3d_b200 import synthetic
    from openseg3d_b200.utils.sharding import frame_seeds, shard_frames, throughput
    seeds = frame_seeds(rank, 2)
    pts, _ = synthetic.make_batch(seeds, 1, False, 8, 100)          # tiny frames: 8 beams x 100 columns
    ms = 10.0 * (rank + 1)                                          # rank 1 is the slow one
    value, max_ms = throughput(pts.shape[0], ms, steps=3)
    gathered = [None] * world
    dist.all_gather_object(gathered, (seeds, pts.shape[0], shard_frames(7, rank, world)))
    if rank == 0:
        torch.save({'value': value, 'max_ms': max_ms, 'gathered': gathered}, out)
    dist.destroy_process_group()


def test_two_rank_frame_sharding_and_aggregation(tmp_path):
    out = str(tmp_path / 'r0.pt')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out, weights_only=False)
    (s0, n0, f0), (s1, n1, f1) = r['gathered']
    assert s0 == [0, 1] and s1 == [2, 3]                            # disjoint contiguous blocks of frames
    assert sorted(f0 + f1) == list(range(7)) and not set(f0) & set(f1)
    assert r['max_ms'] == 20.0                                       # slowest rank
    assert abs(r['value'] - (n0 + n1) * 3 / 20e-3) < 1e-6           # whole-job units / slowest time


def test_single_process_totals_need_no_group():
    from openseg3d_b200.utils.sharding import job_totals
    assert job_totals(5, 2.5) == (5.0, 2.5)
