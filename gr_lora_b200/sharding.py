"""Multi-GPU plumbing for the one place the path shards: independent
(channel, SF) streams are dealt to ranks round-robin (gpu = stream_id mod G; every stream costs
the same bytes/s whatever its SF, so equal counts = equal load), and the chirp/twiddle table blob is broadcast ONCE at
init so every rank demodulates against bit-identical tables.  No per-symbol collective."""
from __future__ import annotations

import numpy as np


def shard_streams(n_streams: int, world: int, rank: int) -> np.ndarray:
    """Global stream ids owned by `rank`: stream_id mod world == rank."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return np.arange(rank, n_streams, world, dtype=np.int64)


def owner_of(stream_id: int, world: int) -> int:
    return int(stream_id) % int(world)


def broadcast_tables(dec, dist, device=None, src: int = 0) -> None:
    """One collective at init: rank `src`'s device table blob overwrites everyone's, in place
    (NCCL on GPUs; on CPU-only test runs `dec` is a host blob ndarray and gloo moves it)."""
    import torch
    if isinstance(dec, np.ndarray):                      # host blob (CPU tests)
        t = torch.from_numpy(dec)
        dist.broadcast(t, src=src)
        return
    view = torch.as_tensor(dec.tables_device_view(), device=device)
    dist.broadcast(view, src=src)
    torch.cuda.synchronize()
    dec.tables_commit()


def aggregate_rate(units_per_rank: int, ms_local: float, dist=None, device=None) -> float:
    """Whole-job units/s: all ranks' units divided by the slowest rank's time."""
    import torch
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    t = torch.tensor([ms_local], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return world * units_per_rank / (float(t.item()) * 1e-3)
