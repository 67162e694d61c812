"""Frame sharding of the hot path over the GPUs of one box.

Frames are independent units (the batch index is the top key of every hash and window id), so the path shards by whole
frames exactly as the reference's DistributedSampler does (seg3d/datasets/samplers/distributed_sampler.py:35-58): no
data-path collective at inference.  The only cross-rank traffic of a benchmark / evaluation run is the bookkeeping
below: total units processed (SUM) and the slowest rank's time (MAX)."""
import torch
import torch.distributed as dist


def frame_seeds(rank, frames_per_rank):
    """Synthetic-frame seeds of one rank: ranks own disjoint, contiguous blocks of frames."""
    return [rank * frames_per_rank + i for i in range(frames_per_rank)]


def shard_frames(n_frames, rank, world):
    """Indices of the frames rank `rank` processes when `n_frames` frames are dealt round-robin
    (DistributedSampler's indices[rank::world] without padding)."""
    return list(range(rank, n_frames, world))


def job_totals(units, elapsed_ms, device=None):
    """(units processed by all ranks, max elapsed ms over ranks).  Works without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(units), float(elapsed_ms)
    t = torch.tensor([float(units)], dtype=torch.float64, device=device)
    m = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return float(t.item()), float(m.item())


def throughput(units, elapsed_ms, steps, device=None):
    """Whole-job units per second: all ranks' units per step * steps / slowest rank's time."""
    total, ms = job_totals(units, elapsed_ms, device)
    return total * steps / (ms * 1e-3), ms
