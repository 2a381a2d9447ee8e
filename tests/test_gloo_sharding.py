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


# ---- ShardedBus's host logic at world_size 2 on CPU/gloo: the construction handshake (IPC handle from rank 0 to the others,
#      all-or-nothing fallback), SPMD put/fanout routing and the cross-rank digest fold.  The GPU side is replaced by a
#      recording stand-in (`bus_factory`); the real thing runs in tests/test_gpu_multi.py on hardware. ----
class _FakeBus:
    fail_create = False

    def __init__(self, n, **kw):
        self.n, self.kw, self.calls = n, kw, []

    def stream_create(self, slots, n_consumers):
        if _FakeBus.fail_create:
            from containerpilot_b200 import _native as nat
            raise nat.CpbusError.__new__(nat.CpbusError)
        self.calls.append(("create", slots, n_consumers))
        return "st0", b"H" * 64

    def stream_open(self, handle, idx):
        self.calls.append(("open", handle, idx))
        return f"st{idx}"

    def stream_put(self, st, ev, now, raw=False, nowait=False):
        self.calls.append(("put", st, len(ev), now)); return 0

    def stream_fanout(self, st, n, now):
        self.calls.append(("fanout", st, n, now)); return 0

    def stream_close(self, st):
        self.calls.append(("close", st))

    def digest_fold(self, first, count):
        return (count * 10, (first + 1) * 7, 1 << (first % 60), count)

    def close(self):
        self.calls.append(("destroy",))


def _sb_worker(rank, world, port, out, fail):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _FakeBus.fail_create = fail
        sb = sharding.ShardedBus(37, dist=dist, rank=rank, world=world, bus_factory=_FakeBus, batch_cap=64, stream_slots=8)
        ev = np.zeros(5, dtype=ob.EVENT_DTYPE)
        if sb.stream_ok:
            assert sb.publish(ev, 1000) == 0
        fold = sb.digest_fold_all()
        calls = list(sb.bus.calls)
        sb.close()
        torch.save({"first": sb.first, "count": sb.count, "ok": sb.stream_ok, "ingest": sb.ingest, "calls": calls + sb.bus.calls[len(calls):],
                    "fold": fold}, f"{out}.{rank}")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("fail", [False, True])
def test_sharded_bus_handshake_and_routing_at_world2(tmp_path, fail):
    out = str(tmp_path / "sb")
    mp.spawn(_sb_worker, args=(2, _free_port(), out, fail), nprocs=2, join=True)
    r0, r1 = (torch.load(f"{out}.{r}", weights_only=False) for r in (0, 1))
    assert (r0["first"], r0["count"], r1["first"], r1["count"]) == (0, 19, 19, 18)          # contiguous shards
    assert r0["fold"] == r1["fold"] == (370, (7 + 140) & (2**64 - 1), 1 ^ (1 << 19), 37)    # sums / xor over both shards
    if fail:                                                                                 # rank 0 could not create the ring:
        assert not r0["ok"] and not r1["ok"] and r1["ingest"] == "local"                     # EVERY rank falls back together
        assert not any(c[0] == "open" for c in r1["calls"])
        return
    assert r0["ok"] and r1["ok"] and "nvlink-stream" in r1["ingest"]
    assert ("create", 8, 2) in r0["calls"] and ("open", b"H" * 64, 1) in r1["calls"]          # the 64-byte handle travelled
    assert ("put", "st0", 5, 1000) in r0["calls"] and not any(c[0] == "put" for c in r1["calls"])   # only the publisher puts
    assert ("fanout", "st0", 5, 1000) in r0["calls"] and ("fanout", "st1", 5, 1000) in r1["calls"]  # everyone fans out
    assert r1["calls"].index(("close", "st1")) < r1["calls"].index(("destroy",))             # importers unmap before destroy
