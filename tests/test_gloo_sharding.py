"""N>1 host path on CPU: world_size-2 gloo.  Rank 0 owns the event stream and broadcasts it;
every rank feeds its contiguous shard (here: the CPU oracle standing in for the GPU, which is
absent in this container) and the per-subscriber digests must equal a single-shard run —
the shard-count invariance the multi-GPU design relies on."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_binding as ob
from containerpilot_b200 import sharding

N_TOTAL, N_EVENTS, DT, PERIOD = 37, 3000, 10_000, 250_000


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _trace():
    rng = np.random.default_rng(0xC0DEB204)
    masks = np.where(rng.random(N_TOTAL) < 0.5, 0x1FFFF, rng.integers(0, 1 << 17, N_TOTAL)).astype(np.uint32)
    codes = rng.integers(0, 17, N_EVENTS).astype(np.uint32)
    srcs = rng.integers(0, 4096, N_EVENTS).astype(np.uint32)
    srcs[::3] = rng.integers(0, 8, len(srcs[::3]))             # a hot set of sources, so that exact {code, source} cases match
    return masks, codes, srcs


def _pairs(i):
    """every third subscriber also pushes down exact {code, source} cases (second-level filter)"""
    if i % 3:
        return None
    rng = np.random.default_rng(1000 + i)
    return [(int(rng.integers(0, 17)), int(rng.integers(0, 8))) for _ in range(int(rng.integers(1, 9)))]


def _run_shard(first, count, masks, codes, srcs):
    orc = ob.Oracle(max(count, 1), timers_per_sub=1, keep_window=64, sub_id_base=first)
    for i in range(count):
        orc.subscribe(int(masks[first + i]), _pairs(first + i))
        orc.timer_add(first + i, PERIOD, 9000 + first + i, False)
    assert orc.publish_many(codes, srcs, dt_ns=DT) == 0
    return np.array([[orc.count(first + i), orc.digest(first + i)] for i in range(count)], dtype=np.uint64).reshape(count, 2)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        masks, codes, srcs = _trace()
        ev = sharding.stamp_trace(codes, srcs, DT)
        stream = torch.from_numpy(ev.view(np.uint8).reshape(-1, 32).copy()) if rank == 0 else torch.zeros((N_EVENTS, 32), dtype=torch.uint8)
        for lo in range(0, N_EVENTS, 256):                      # batch by batch, like bench.py
            sharding.broadcast_events(dist, stream[lo:lo + 256], src=0)
        got = np.frombuffer(stream.numpy().tobytes(), dtype=ob.EVENT_DTYPE)
        assert (got["code"] == codes).all() and (got["source_id"] == srcs).all() and (got["seq"] == np.arange(N_EVENTS)).all()
        first, count = sharding.shard_range(N_TOTAL, world, rank)
        res = _run_shard(first, count, masks, got["code"].copy(), got["source_id"].copy())
        gathered = [None] * world
        dist.all_gather_object(gathered, (first, res))
        if rank == 0:
            full = np.zeros((N_TOTAL, 2), dtype=np.uint64)
            for f, r in gathered:
                full[f:f + len(r)] = r
            np.save(out, full)
    finally:
        dist.destroy_process_group()


def test_shard_ranges_partition_everything():
    for n in (1, 7, 37, 65_536, 8_388_608):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            for r in range(world):
                first, count = sharding.shard_range(n, world, r)
                assert first == covered
                covered += count
                if count:
                    assert sharding.owner_of(first, n, world) == r and sharding.owner_of(first + count - 1, n, world) == r
            assert covered == n


@pytest.mark.timeout(180)
def test_world2_broadcast_stream_gives_shard_invariant_digests(tmp_path):
    out = str(tmp_path / "digests.npy")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    sharded = np.load(out)
    masks, codes, srcs = _trace()
    single = _run_shard(0, N_TOTAL, masks, codes, srcs)
    assert (sharded == single).all()
    assert single[:, 0].min() > 0
