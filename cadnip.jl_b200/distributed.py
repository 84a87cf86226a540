"""Multi-GPU sharding of a CircuitSweep.

Lanes are independent, so the sweep is block-partitioned over ranks in iteration
order -- rank g owns lanes [g*ceil(P/G), (g+1)*ceil(P/G)) -- with the structure
(pattern, maps, LU schedule) replicated.  There is NO collective on the hot path;
the only exchange is one end-of-run gather of per-lane status and saved waveforms
(``torch.distributed`` over NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Tuple

import numpy as np


def shard_slice(P: int, rank: int, world: int) -> slice:
    """Contiguous block of lanes owned by ``rank`` (may be empty for trailing ranks)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    block = math.ceil(P / world) if P > 0 else 0
    lo = min(P, rank * block)
    hi = min(P, lo + block)
    return slice(lo, hi)


def dist_info() -> Tuple[int, int]:
    """(rank, world) of the default process group, (0, 1) when not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def gather_lanes(local: np.ndarray, P: int, lane_axis: int = -1, dst: int = 0,
                 device: Optional[str] = None) -> Optional[np.ndarray]:
    """Gather per-rank blocks (lane axis = ``lane_axis``) into the full-sweep array on
    ``dst``; other ranks return None.  Blocks are padded to the common block size so a
    single fixed-size gather suffices."""
    import torch
    import torch.distributed as dist
    rank, world = dist_info()
    if world == 1:
        return local
    block = math.ceil(P / world)
    loc = np.moveaxis(np.asarray(local), lane_axis, 0)
    pad = np.zeros((block,) + loc.shape[1:], dtype=loc.dtype)
    pad[:loc.shape[0]] = loc
    backend = dist.get_backend()
    dev = device or ("cuda" if backend == "nccl" else "cpu")
    t = torch.from_numpy(np.ascontiguousarray(pad)).to(dev)
    if rank == dst:
        bufs = [torch.empty_like(t) for _ in range(world)]
        dist.gather(t, bufs, dst=dst)
        full = torch.cat(bufs, dim=0)[:P].cpu().numpy()
        return np.moveaxis(full, 0, lane_axis)
    dist.gather(t, None, dst=dst)
    return None


def run_sharded(P: int, run_local: Callable[[slice], np.ndarray], lane_axis: int = -1,
                gather: bool = True):
    """Run ``run_local(lanes)`` on this rank's block and (optionally) gather on rank 0."""
    rank, world = dist_info()
    sl = shard_slice(P, rank, world)
    local = run_local(sl)
    if not gather:
        return local
    return gather_lanes(local, P, lane_axis)
