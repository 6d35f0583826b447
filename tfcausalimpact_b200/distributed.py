"""Series sharding across the GPUs of one box.

Series are statistically independent and all chains of a series stay on one GPU (they share the
TMA-staged design-matrix chunk), so the data path needs NO collective: every rank fits its own
contiguous block of series.  The only exchange is the final gather of per-series summaries
(21 floats per series).  One process per GPU, `torch.distributed` (NCCL on GPUs; gloo in the CPU
tests).
"""
import os
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `n` series owned by `rank`: the first n % world ranks hold one
    extra series."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment; initialises the process group when
    world > 1 (NCCL if CUDA is present, else gloo)."""
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def gather_series(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather per-series rows: `local` is [n_local, ...] for this rank's `shard_range`; returns
    [n_total, ...] on every rank (blocks padded to a common size for the collective)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert local.shape[0] == n_total
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_range(n_total, rank, world)
    assert local.shape[0] == hi - lo, f'rank {rank} holds {local.shape[0]} series, expected {hi - lo}'
    width = (n_total + world - 1) // world
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:hi - lo] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad.contiguous())
    out = []
    for r in range(world):
        a, b = shard_range(n_total, r, world)
        out.append(parts[r][:b - a])
    return torch.cat(out, dim=0)


def max_over_ranks(value: float, device=None) -> float:
    """Max of a scalar over ranks (timing)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bind_to_gpu_numa(local_rank: int) -> bool:
    """Pin this process to the CPUs NVML reports as closest to its GPU, so that pinned host buffers (first
    touched by this process) and the threads feeding the PCIe copies sit on the GPU's NUMA node.  With 8 ranks
    uploading from one host this decides the host-to-device bandwidth every rank gets.  Returns False (and
    changes nothing) when NVML or the affinity call is unavailable."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False
