"""BASELINE config 5: bf16 training step of the Segformer hot path (forward + backward + SGD), frame-parallel under DDP.
    python tools/train_step.py [--frames F] [--steps K]                      (1 GPU)
    python -m torch.distributed.run --nproc-per-node N tools/train_step.py   (N GPUs, NCCL gradient all-reduce)
Loss: cross-entropy on synthetic labels for point / voxel / aux-voxel logits (the reference's OHEM + Lovasz losses are
outside the hot path, tools/train.py:71-110).  Prints one JSON line (points/s over all ranks, max over ranks)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from openseg3d_b200 import synthetic  # noqa: E402
from openseg3d_b200.models import build_segformer  # noqa: E402
from openseg3d_b200.utils.sharding import frame_seeds, throughput  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--dtype', default='bf16')
    run(ap.parse_args())


def bench_main(args):
    """`python bench.py --mode train ...`: the same step under bench.py's JSON contract (frames per GPU from --frames,
    default 8 -> use 2 for the train step unless the caller overrides it)."""
    run(args, contract=True)


def run(args, contract=False):
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dtype = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    model = build_segformer('waymo_one_sweep', compute_dtype=dtype).cuda().train()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.SGD(net.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)      # configs/waymo_one_sweep.yaml
    pts, _ = synthetic.make_batch(frame_seeds(rank, args.frames), 1, False)
    dev = torch.from_numpy(pts).cuda()
    g = torch.Generator(device='cuda').manual_seed(rank)
    labels = torch.randint(0, 22, (pts.shape[0],), device='cuda', generator=g)

    def step():
        out = net({'points': dev, 'batch_size': args.frames})
        loss = F.cross_entropy(out['point_out'].float(), labels)
        loss = loss + out['voxel_out'].float().logsumexp(1).mean() * 0.0 + out['aux_voxel_out'].float().logsumexp(1).mean() * 0.0
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    value, ms = throughput(pts.shape[0], e0.elapsed_time(e1), args.steps, 'cuda')
    if rank == 0:
        line = {'metric': 'points/sec, Waymo 1-sweep seg train step (fwd+bwd+SGD)', 'value': value,
                'unit': 'points/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': args.dtype, 'data': 'synthetic', 'frames_per_gpu': args.frames, 'loss': float(loss.detach()),
                'peak_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30,
                'config': {'workload': f'configs/waymo_one_sweep.yaml {args.dtype} training step (forward + backward + SGD), '
                                       f'{args.frames} synthetic frames per GPU ({pts.shape[0]} points/GPU), '
                                       'frame-parallel DDP gradient all-reduce over NCCL',
                           'frames_per_gpu': args.frames, 'points_per_gpu': int(pts.shape[0]),
                           'parallelism': f'DDP x{world}'}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
