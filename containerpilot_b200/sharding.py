"""Multi-GPU host logic: subscribers partition into contiguous shards,
one process per GPU; the only exchange is the event stream, broadcast from the publisher's rank.

Nothing here moves subscriber state; per-subscriber sequences (and digests) are independent of
the shard count because records carry GLOBAL subscriber ids (`cpbus_config.sub_id_base`).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, near-even partition: returns (first_global_id, count) of `rank`'s shard."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_total, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def owner_of(sub_id: int, n_total: int, world: int) -> int:
    """Rank that holds global subscriber `sub_id` (inverse of shard_range)."""
    base, extra = divmod(n_total, world)
    cut = extra * (base + 1)
    return sub_id // (base + 1) if sub_id < cut else extra + (sub_id - cut) // max(base, 1)


def broadcast_events(dist, events_u8, src: int = 0):
    """Broadcast a batch (uint8 tensor [n, 32], on the backend's device) from `src` to every rank.
    `dist` is torch.distributed (NCCL over NVLink on GPUs, gloo in CPU tests)."""
    dist.broadcast(events_u8, src=src)
    return events_u8


def stamp_trace(codes: np.ndarray, sources: np.ndarray, dt_ns: int, first_seq: int = 0) -> np.ndarray:
    """Complete 32-byte records for a device-resident trace: seq = publish ordinal, ts = (seq+1)*dt."""
    from .bus import EVENT_DTYPE
    n = len(codes)
    ev = np.zeros(n, dtype=EVENT_DTYPE)
    ev["seq"] = first_seq + np.arange(n, dtype=np.uint64)
    ev["ts_ns"] = (first_seq + 1 + np.arange(n, dtype=np.uint64)) * np.uint64(dt_ns)
    ev["code"], ev["source_id"], ev["target"] = codes, sources, 0xFFFFFFFF
    return ev
