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


def pin_to_gpu_numa_node(device: int) -> Optional[str]:
    """Restrict this process to the CPUs of the NUMA node the GPU hangs off, so that the pinned
    host buffers it allocates next (first touch) and its copy-engine traffic stay on that socket.
    Returns the CPU list used, or None when the topology is not readable (nothing changed)."""
    import os
    try:
        import torch
        bus = torch.cuda.get_device_properties(device).pci_bus_id
        dom = torch.cuda.get_device_properties(device).pci_domain_id
        dev = torch.cuda.get_device_properties(device).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/local_cpulist"
        with open(path) as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return txt
    except Exception:
        return None


class SharedSweepBuffer:
    """One host array shared by all ranks of a node: the destination of the end-of-run gather.

    Rank 0 creates a POSIX shared-memory segment, every rank maps it and page-locks the mapping
    (cudaHostRegister), and each rank's GPU then copies ITS lane block straight into its columns
    of the one full-sweep array (cb200_tran_fetch_ld) -- no staging buffer, no host-side
    concatenation, every GPU on its own PCIe link.  With world == 1 this is a plain pinned array.
    ``array`` has the given shape on every rank; rank 0 owns the result."""

    def __init__(self, shape, dtype=np.float64, pin: bool = True):
        import torch
        self.shape = tuple(int(x) for x in shape)
        self.dtype = np.dtype(dtype)
        nbytes = max(1, int(np.prod(self.shape)) * self.dtype.itemsize)
        rank, world = dist_info()
        self._shm = None
        self._registered = None
        if world == 1:
            if pin and torch.cuda.is_available():
                self._t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            else:
                self._t = torch.empty(nbytes, dtype=torch.uint8)
            buf = self._t.numpy()
        else:
            import torch.distributed as dist
            from multiprocessing import shared_memory
            name = [None]
            if rank == 0:
                self._shm = shared_memory.SharedMemory(create=True, size=nbytes)
                name[0] = self._shm.name
            dist.broadcast_object_list(name, src=0)
            if rank != 0:
                self._shm = shared_memory.SharedMemory(name=name[0])
                try:                                         # rank 0 unlinks; stop the tracker double-freeing
                    from multiprocessing import resource_tracker
                    resource_tracker.unregister(self._shm._name, "shared_memory")
                except Exception:
                    pass
            buf = np.frombuffer(self._shm.buf, dtype=np.uint8, count=nbytes)
            if pin and torch.cuda.is_available():
                ptr = buf.ctypes.data
                rc = torch.cuda.cudart().cudaHostRegister(ptr, nbytes, 0)
                if int(rc) != 0:
                    raise RuntimeError(f"cudaHostRegister failed: {rc}")
                self._registered = ptr
        self.array = buf[:int(np.prod(self.shape)) * self.dtype.itemsize].view(self.dtype).reshape(self.shape)

    def lane_block(self, sl: slice) -> np.ndarray:
        """This rank's columns: a view [..., sl] whose rows are ``shape[-1]`` elements apart."""
        return self.array[..., sl]

    def close(self):
        import torch
        rank, world = dist_info()
        if self._registered is not None:
            torch.cuda.cudart().cudaHostUnregister(self._registered)
            self._registered = None
        self.array = None
        if self._shm is not None:
            if world > 1:
                import torch.distributed as dist
                dist.barrier()
            try:
                self._shm.close()
            except BufferError:
                # a caller still holds a view of the array: leave the mapping to process exit (and keep
                # SharedMemory.__del__ from retrying the close and reporting the same error)
                self._shm._mmap = None
            if rank == 0:
                self._shm.unlink()
            self._shm = None


def run_sharded(P: int, run_local: Callable[[slice], np.ndarray], lane_axis: int = -1,
                gather: bool = True):
    """Run ``run_local(lanes)`` on this rank's block and (optionally) gather on rank 0."""
    rank, world = dist_info()
    sl = shard_slice(P, rank, world)
    local = run_local(sl)
    if not gather:
        return local
    return gather_lanes(local, P, lane_axis)
