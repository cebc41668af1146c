"""Frame sharding for the multi-GPU driver: frames are independent, so rank r of `world` takes a contiguous
block of the batch and no collective touches the data path.  Only the timing reduction
(max over ranks) and an optional gather of per-frame counts go through torch.distributed."""
from __future__ import annotations


def shard_range(nframes: int, rank: int, world: int) -> tuple[int, int]:
    """[begin, end) of the frames owned by `rank`; blocks differ by at most one frame."""
    if world < 1 or not (0 <= rank < world) or nframes < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(nframes, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def reduce_max_time(ms: float, backend_device=None) -> float:
    """max over ranks of a device-measured time (every multi-GPU number is the slowest rank's)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=backend_device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(counts: list[int], backend_device=None) -> list[int]:
    """per-frame keypoint counts of every rank, in global frame order (rank-major blocks)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(counts)
    world = dist.get_world_size()
    n = torch.tensor([len(counts)], dtype=torch.int64, device=backend_device or "cpu")
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    m = int(max(s.item() for s in sizes))
    buf = torch.full((m,), -1, dtype=torch.int64, device=backend_device or "cpu")
    buf[:len(counts)] = torch.tensor(counts, dtype=torch.int64)
    outs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    return [int(v) for o, s in zip(outs, sizes) for v in o[: int(s.item())].tolist()]
