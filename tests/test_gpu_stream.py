"""The publisher's stream (cpbus_stream_*): flagged ring + in-kernel pull + ack + device-managed prefetch, driven on
whatever GPUs this box has (all shards on one GPU when there is only one; real peer access when there are more), plus the
bounded waits, the kernel-side publish accounting of device batches and the prefetch-cache rules.  Bit-exact vs the oracle."""
import os

import numpy as np
import pytest

import oracle_binding as ob
import trace as tr
from containerpilot_b200 import _native as nat
from containerpilot_b200.bus import Bus, EVENT_DTYPE
from containerpilot_b200.sharding import LocalShardedBus, shard_range

pytestmark = pytest.mark.gpu


def _devices(g):
    import torch
    nd = torch.cuda.device_count()
    return [i % nd for i in range(g)]


def _batches(seed, n_batches, B, ragged=True):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_batches):
        n = int(rng.integers(0, B + 1)) if ragged and rng.random() < 0.3 else B
        ev = np.zeros(n, dtype=EVENT_DTYPE)
        ev["code"] = rng.integers(0, 17, n)
        ev["source_id"] = rng.integers(0, 64, n)
        out.append(ev)
    return out


def _oracle_for_shard(first, count, masks, batches, dt, period, K):
    orc = ob.Oracle(max(count, 1), timers_per_sub=K, keep_window=1024, sub_id_base=first)
    for i in range(count):
        orc.subscribe(int(masks[first + i]))
        if K:
            orc.timer_add(first + i, period, 7000 + first + i, False)
    now = 0
    for ev in batches:
        now += dt
        assert orc.advance(now) == 0
        for c, s_ in zip(ev["code"], ev["source_id"]):
            assert orc.publish(int(c), int(s_)) == 0
    return orc


@pytest.mark.parametrize("G,K,lookahead", [(1, 0, 0), (2, 1, 0), (2, 1, 3), (4, 1, 5), (3, 0, 2)])
def test_stream_shards_match_oracle_and_each_other(G, K, lookahead):
    """Every shard count gives every subscriber the same (count, digest, window) as the oracle; the publisher may run
    `lookahead` batches ahead of the fan-outs (device-managed prefetch of batch q+2 kicks in from lookahead >= 2)."""
    N, B, dt, period = 157, 64, 40_000, 90_000
    rng = np.random.default_rng(42 + G)
    masks = np.where(rng.random(N) < 0.5, nat.MASK_ALL, rng.integers(0, 1 << 17, N)).astype(np.uint32)
    batches = _batches(7 + K, 40, B)
    sb = LocalShardedBus(N, _devices(G), ring_cap=1024, batch_cap=B, timers_per_sub=K, stream_slots=8)
    try:
        sb.subscribe_many(masks)
        if K:
            sb.timer_add_many(period, source_id0=7000)
        put = 0
        for j, ev in enumerate(batches):
            while put < len(batches) and put <= j + lookahead:
                rc = sb.put(batches[put], (put + 1) * dt)
                if rc == nat.EAGAIN:        # the ring is full of batches the (asynchronous) fan-outs have not pulled yet
                    if put > j:
                        break               # running ahead is optional
                    sb.sync()               # the batch about to be fanned out must go in: let the queued launches finish
                    continue
                nat.check(rc, "cpbus_stream_put")
                put += 1
            sb.fanout(len(ev), (j + 1) * dt)
        sb.sync()
        for g, (first, count, bus) in enumerate(sb.shards):
            assert bus.stream_status(sb._st[g]) == nat.OK
            orc = _oracle_for_shard(first, count, masks, batches, dt, period, K)
            tr.compare(bus, orc, count, sub_id_base=first)
    finally:
        sb.close()


def test_stream_ring_wraps_and_put_backpressure():
    """n_slots = 4: the publisher cannot overwrite a slot before every consumer acknowledged it (EAGAIN), and the ring
    wraps many times without losing or duplicating a batch."""
    N, B, dt = 40, 32, 1000
    batches = _batches(99, 64, B, ragged=False)
    sb = LocalShardedBus(N, _devices(2), ring_cap=256, batch_cap=B, stream_slots=4)
    try:
        sb.subscribe_many(np.full(N, nat.MASK_ALL, dtype=np.uint32))
        for q in range(4):
            nat.check(sb.put(batches[q], (q + 1) * dt), "put")
        assert sb.put(batches[4], 5 * dt) == nat.EAGAIN         # slot of batch 1 not acknowledged yet (NOWAIT: one thread drives both sides)
        put = 4
        for j, ev in enumerate(batches):
            sb.fanout(len(ev), (j + 1) * dt)
            sb.sync()
            while put < len(batches) and put < j + 1 + 4:
                rc = sb.put(batches[put], (put + 1) * dt)
                if rc == nat.EAGAIN:
                    break
                nat.check(rc, "put"); put += 1
            assert put > j + 1 or put == len(batches)          # after a sync at least the next batch always fits
        sb.sync()
        for first, count, bus in sb.shards:
            orc = _oracle_for_shard(first, count, np.full(N, nat.MASK_ALL, dtype=np.uint32), batches, dt, 0, 0)
            tr.compare(bus, orc, count, sub_id_base=first, window=256)
    finally:
        sb.close()


def test_consumer_that_is_not_told_the_shapes_polls_the_headers():
    """cpbus_stream_poll: a consumer follows the publisher knowing nothing but the stream — ragged batches, clock from the header."""
    N, B = 50, 64
    batches = _batches(5, 12, B)
    masks = np.full(N, nat.MASK_ALL, dtype=np.uint32)
    sb = LocalShardedBus(N, _devices(2), ring_cap=1024, batch_cap=B, stream_slots=8)
    try:
        sb.subscribe_many(masks)
        pub_bus, pub_st = sb.shards[0][2], sb._st[0]
        con_bus, con_st = sb.shards[1][2], sb._st[1]
        assert con_bus.stream_poll(con_st) is None                  # nothing released yet
        for q, ev in enumerate(batches):
            nat.check(sb.put(ev, (q + 1) * 7000), "put")
            nat.check(pub_bus.stream_fanout(pub_st, len(ev), (q + 1) * 7000), "fanout")     # the publisher's own shard knows
            shape = None                                            # the other one asks (the release is an asynchronous copy)
            for _ in range(100_000):
                shape = con_bus.stream_poll(con_st)
                if shape is not None:
                    break
            assert shape == (len(ev), (q + 1) * 7000)
            nat.check(con_bus.stream_fanout(con_st, *shape), "fanout")
        assert con_bus.stream_poll(con_st) is None
        sb.sync()
        for first, count, bus in sb.shards:
            orc = _oracle_for_shard(first, count, masks, batches, 7000, 0, 0)
            tr.compare(bus, orc, count, sub_id_base=first)
    finally:
        sb.close()


def test_stream_wait_is_bounded_and_reports_timeout():
    """A consumer launched for a batch the publisher never released gives up after the configured bound: the kernel ends,
    nothing is delivered, the bus reports CPBUS_ETIMEDOUT from then on."""
    with Bus(64, ring_cap=256, batch_cap=32, digest=True) as bus:
        bus.subscribe_many(np.full(64, nat.MASK_ALL, dtype=np.uint32))
        st, _ = bus.stream_create(4, 1)
        bus.stream_set_timeout(st, 20_000)                       # 20 ms
        nat.check(bus.stream_fanout(st, 32, 1000), "fanout")    # nothing was put
        bus.sync()
        assert bus.stream_status(st) == nat.ETIMEDOUT
        assert bus.stream_fanout(st, 32, 2000) == nat.ETIMEDOUT
        assert bus.stats()["deliveries"] == 0
        assert int(bus.digests(0, 64)["count"].max()) == 0
        bus.stream_close(st)


def test_stream_shape_mismatch_aborts_the_launch():
    with Bus(8, ring_cap=256, batch_cap=32) as bus:
        bus.subscribe_many(np.full(8, nat.MASK_ALL, dtype=np.uint32))
        st, _ = bus.stream_create(4, 1)
        ev = np.zeros(16, dtype=EVENT_DTYPE); ev["code"] = 3
        nat.check(bus.stream_put(st, ev, 1000), "put")
        nat.check(bus.stream_fanout(st, 17, 1000), "fanout")    # the header says 16
        bus.sync()
        assert bus.stream_status(st) == nat.ETIMEDOUT and bus.stats()["deliveries"] == 0
        bus.stream_close(st)


@pytest.mark.parametrize("K", [0, 1])
def test_every_cta_can_build_the_descriptor_itself(K, monkeypatch):
    """CPBUS_HINTS=2 is the test hook for the bounded-spin fallback: no CTA waits for CTA 0, each one hashes the batch and
    builds the descriptor from its own staged copy.  Results must be identical."""
    monkeypatch.setenv("CPBUS_HINTS", "2")
    ops, n_total = tr.random_ops(4242 + K, 300, 1500, timers_per_sub=K, p_member=0.0)
    with Bus(n_total + 16, ring_cap=1024, batch_cap=128, timers_per_sub=K, digest=True) as bus:
        tr.run_bus(bus, ops)
        orc = tr.run_oracle(ops, n_total + 16, timers_per_sub=K, keep_window=1024)
        tr.compare(bus, orc, n_total)


def test_device_published_batches_are_accounted_by_the_kernel():
    """published_by_code, the {code, source} publish counts (events/bus.go:130-132) and the debug ring (bus.go:139) see
    batches that arrive in device memory exactly as if each event had gone through cpbus_publish."""
    import torch
    B, n_b = 64, 5
    rng = np.random.default_rng(5)
    codes = rng.integers(0, 17, B * n_b).astype(np.uint32)
    srcs = rng.integers(0, 6, B * n_b).astype(np.uint32)
    ev = np.zeros(B * n_b, dtype=EVENT_DTYPE)
    ev["seq"] = np.arange(B * n_b); ev["ts_ns"] = (np.arange(B * n_b) + 1) * 10
    ev["code"], ev["source_id"], ev["target"] = codes, srcs, nat.TARGET_ALL
    ev["target"][5::17] = 3; ev["flags"][5::17] = nat.F_UNICAST        # some unicast records: not "published"
    bcast = ev["target"] == nat.TARGET_ALL
    dev = torch.from_numpy(ev.view(np.uint8).reshape(-1, 32).copy()).cuda()
    with Bus(16, ring_cap=1024, batch_cap=B, digest=True) as bus, Bus(16, ring_cap=1024, batch_cap=B, digest=True) as ref:
        for b_ in (bus, ref):
            b_.subscribe_many(np.full(16, nat.MASK_ALL, dtype=np.uint32))
        # host publishes before and after, so that the debug ring has to interleave the two kinds in order
        nat.check(bus.publish(14, 1), "publish"); nat.check(ref.publish(14, 1), "publish")
        for i in range(n_b):
            nat.check(bus.publish_device(dev.data_ptr() + i * B * 32, B, int(ev["ts_ns"][(i + 1) * B - 1])), "publish_device")
        for e in ev[bcast]:
            nat.check(ref.advance(int(e["ts_ns"])), "advance")
            nat.check(ref.publish(int(e["code"]), int(e["source_id"])), "publish")
        nat.check(bus.publish(15, 2), "publish"); nat.check(ref.publish(15, 2), "publish")
        nat.check(bus.flush(), "flush"); nat.check(ref.flush(), "flush")
        sa, sr = bus.stats(), ref.stats()
        assert sa["published_by_code"] == sr["published_by_code"]
        assert bus.publish_counts() == ref.publish_counts()
        want = {}
        for c, s_ in zip(codes[bcast], srcs[bcast]):
            if c != 13:
                want[(int(c), int(s_))] = want.get((int(c), int(s_)), 0) + 1
        want[(14, 1)] = want.get((14, 1), 0) + 1; want[(15, 2)] = want.get((15, 2), 0) + 1
        assert bus.publish_counts() == want
        da, dr = bus.debug_events(), ref.debug_events()
        assert [(int(r["code"]), int(r["source_id"])) for r in da] == [(int(r["code"]), int(r["source_id"])) for r in dr]
        assert len(dr) > 0


def test_prefetch_cache_never_serves_a_stale_copy():
    """Two peer buffers used alternately (A, B, A, B ... with d_next = the other one), rewritten between uses: a
    prefetched copy is good for one use only, so every launch must deliver the CURRENT content."""
    import torch
    B, rounds = 32, 12
    rng = np.random.default_rng(11)
    bufs = [torch.zeros((B, 32), dtype=torch.uint8, device="cuda") for _ in range(2)]
    with Bus(8, ring_cap=1024, batch_cap=B, digest=True) as bus:
        bus.subscribe_many(np.full(8, nat.MASK_ALL, dtype=np.uint32))
        orc = ob.Oracle(8, keep_window=1024)
        for _ in range(8):
            orc.subscribe()
        seq = 0
        def fill(k):
            nonlocal seq
            ev = np.zeros(B, dtype=EVENT_DTYPE)
            ev["seq"] = seq + np.arange(B); ev["ts_ns"] = (seq + 1 + np.arange(B)) * 10
            ev["code"] = rng.integers(1, 17, B); ev["source_id"] = rng.integers(0, 99, B); ev["target"] = nat.TARGET_ALL
            seq += B
            bufs[k].copy_(torch.from_numpy(ev.view(np.uint8).reshape(B, 32).copy()))
            torch.cuda.synchronize()
            return ev
        cur = fill(0)
        for r in range(rounds):
            k = r % 2
            bus.sync()                                           # the previous launch (which prefetched buffer k's OLD... nothing yet) is done
            nxt = fill(1 - k)                                    # the other buffer gets NEW content before it is named as d_next
            nat.check(bus.publish_device_staged(bufs[k].data_ptr(), B, int(cur["ts_ns"][-1]), bufs[1 - k].data_ptr(), B), "staged")
            for e in cur:
                orc.advance(int(e["ts_ns"])); orc.publish(int(e["code"]), int(e["source_id"]))
            cur = nxt
        bus.sync()
        tr.compare(bus, orc, 8)


def test_ephemeral_intern_region_is_bounded_and_recycles():
    """Metric payloads ("key|value", control/endpoints.go:125) come from a bounded region: ids stay resolvable for at
    least CPBUS_EPHEMERAL_SLOTS newer payloads, then are recycled; permanent names are untouched."""
    with Bus(4, ring_cap=64, batch_cap=32) as bus:
        name = bus.intern("myjob")
        a = bus.intern_ephemeral("mymetric|1")
        assert a & nat.EPHEMERAL_BIT and bus.source(a) == "mymetric|1"
        assert bus.intern_ephemeral("mymetric|1") == a           # equal strings that are both live share an id
        assert bus.intern_ephemeral("myjob") == name             # already a name: one id per string
        st0 = bus.stats()
        ids = [bus.intern_ephemeral(f"m|{i}") for i in range(nat.EPHEMERAL_SLOTS + 10)]
        st1 = bus.stats()
        assert st1["intern_entries"] == st0["intern_entries"]    # the permanent table did not grow
        assert st1["ephemeral_live"] == nat.EPHEMERAL_SLOTS and st1["ephemeral_recycled"] >= 10
        assert bus.source(ids[-1]) == f"m|{nat.EPHEMERAL_SLOTS + 9}"
        with pytest.raises(nat.CpbusError) as ei:
            bus.source(a)                                        # recycled long ago
        assert ei.value.status == nat.ENOENT
        assert bus.source(name) == "myjob"
