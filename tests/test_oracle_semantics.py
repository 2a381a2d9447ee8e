"""Properties of the reference bus that the oracle must hold (code reading of
events/bus.go, subscriber.go, timer.go), plus known answers for the digest spec.  CPU-only."""
import numpy as np

import oracle_binding as ob
import trace as tr

M64 = 0xFFFFFFFFFFFFFFFF


def py_record_hash(seq, ts, code, src, target, flags):
    """Independent restatement of the digest spec (include/cpbus.h: cpbus_record_hash)."""
    K0, K1, K2, K3, K4 = 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB, 0xD6E8FEB86659FD93, 0xA0761D6478BD642F
    w2, w3 = code | (src << 32), target | (flags << 32)
    x = ((seq + K4) * K0) & M64; x ^= x >> 32
    x = ((x + ts) * K1) & M64; x ^= x >> 32
    x = ((x + w2) * K2) & M64; x ^= x >> 32
    x = ((x + w3) * K3) & M64; x ^= x >> 29
    return x


def test_record_hash_known_answers():
    rng = np.random.default_rng(7)
    recs = np.zeros(64, dtype=ob.EVENT_DTYPE)
    recs["seq"] = rng.integers(0, 1 << 62, 64); recs["ts_ns"] = rng.integers(0, 1 << 62, 64)
    recs["code"] = rng.integers(0, 17, 64); recs["source_id"] = rng.integers(0, 1 << 32, 64)
    recs["target"] = rng.integers(0, 1 << 32, 64); recs["flags"] = rng.integers(0, 4, 64)
    for r in recs:
        want = py_record_hash(*(int(r[k]) for k in ("seq", "ts_ns", "code", "source_id", "target", "flags")))
        assert ob.lib().orc_record_hash(r.tobytes()) == want
    assert ob.lib().orc_digest_multiplier() == 0x9E3779B97F4A7C15
    zero = np.zeros(1, dtype=ob.EVENT_DTYPE)
    assert ob.lib().orc_record_hash(zero.tobytes()) == py_record_hash(0, 0, 0, 0, 0, 0)


def test_every_subscriber_sees_global_publish_order():
    """exclusive lock + FIFO channels => per-subscriber order == publish order."""
    orc = ob.Oracle(8)
    subs = [orc.subscribe() for _ in range(8)]
    rng = np.random.default_rng(1)
    codes, srcs = rng.integers(1, 17, 10_000).astype(np.uint32), rng.integers(0, 64, 10_000).astype(np.uint32)
    assert orc.publish_many(codes, srcs) == 0
    ref = orc.mailbox(subs[0])
    assert len(ref) == 10_000 and (ref["seq"] == np.arange(10_000)).all()
    assert (ref["code"] == codes).all() and (ref["source_id"] == srcs).all()
    h = 0
    for r in ref:
        h = (h * 0x9E3779B97F4A7C15 + py_record_hash(*(int(r[k]) for k in ("seq", "ts_ns", "code", "source_id", "target", "flags")))) & M64
    for s in subs:
        assert orc.mailbox(s).tobytes() == ref.tobytes()
        assert orc.digest(s) == h and orc.count(s) == 10_000


def test_filter_is_the_consumer_switch_pushed_down():
    """with a mask, a mailbox holds exactly the events its consumer would not have dropped."""
    orc = ob.Oracle(3)
    a, m, j = orc.subscribe(), orc.subscribe(1 << 13), orc.subscribe((1 << 8) | (1 << 15))
    rng = np.random.default_rng(2)
    codes = rng.integers(0, 17, 5000).astype(np.uint32); srcs = rng.integers(0, 9, 5000).astype(np.uint32)
    orc.publish_many(codes, srcs)
    full = orc.mailbox(a)
    assert orc.mailbox(m).tobytes() == full[full["code"] == 13].tobytes()
    assert orc.mailbox(j).tobytes() == full[(full["code"] == 8) | (full["code"] == 15)].tobytes()


def test_timers_unicast_under_virtual_time():
    """events/timer.go: periodic fires every period, one-shot once; unicast to the owner; never on the bus."""
    orc = ob.Oracle(2, timers_per_sub=2)
    a, b = orc.subscribe(), orc.subscribe()
    orc.timer_add(a, 1000, 77, False)          # NewEventTimer
    orc.timer_add(a, 2500, 78, True)           # NewEventTimeout
    orc.publish(14, 1)
    orc.advance(999); orc.publish(5, 2)        # nothing due yet
    orc.advance(3000); orc.publish(6, 3)       # ticks at 1000, 2000, 2500(one-shot), 3000 come first
    box = orc.mailbox(a)
    assert [(int(r["code"]), int(r["source_id"]), int(r["ts_ns"]), int(r["flags"])) for r in box] == [
        (14, 1, 0, 0), (5, 2, 999, 0), (8, 77, 1000, 1), (8, 77, 2000, 1), (8, 78, 2500, 1), (8, 77, 3000, 1), (6, 3, 3000, 0)]
    assert [int(r["seq"]) for r in box if r["flags"] == 1] == [0, 1, 0, 2]       # firing ordinal per timer
    assert [int(r["code"]) for r in orc.mailbox(b)] == [14, 5, 6]
    assert len(orc.debug_events()) == 3 and orc.total_ticks() == 4
    orc.advance(10_000)
    assert orc.count(a) == 7 + 7                                                  # periodic kept firing, one-shot did not


def test_full_mailbox_blocks_publisher():
    """events/subscriber.go:30-32: lossless and blocking."""
    orc = ob.Oracle(2, mailbox_cap=4)
    a, b = orc.subscribe(), orc.subscribe(1 << 1)
    for i in range(4):
        assert orc.publish(2, i) == 0
    assert orc.publish(2, 9) == ob.EAGAIN and orc.count(a) == 4
    assert orc.publish(1, 9) == ob.EAGAIN      # code 1 targets `a` (full) and `b`: blocks, delivers to nobody
    assert orc.count(b) == 0
    assert len(orc.consume(a, 2)) == 2
    assert orc.publish(2, 9) == 0 and orc.count(a) == 5


def test_metric_not_counted_everything_else_is():
    """events/bus.go:130-132"""
    orc = ob.Oracle(1); orc.subscribe()
    for c in (13, 13, 14, 1, 1, 1):
        orc.publish(c, 0)
    assert orc.published_by_code(13) == 0 and orc.published_by_code(14) == 1 and orc.published_by_code(1) == 3


def test_random_trace_is_deterministic_and_shard_invariant():
    """a subscriber's sequence does not depend on which shard (sub_id_base) holds it."""
    ops, n_total = tr.random_ops(11, 12, 3000, timers_per_sub=2)
    o1 = tr.run_oracle(ops, 64, timers_per_sub=2)
    o2 = tr.run_oracle(ops, 64, timers_per_sub=2)
    for s in range(n_total):
        assert o1.digest(s) == o2.digest(s) and o1.count(s) == o2.count(s)
    assert o1.total_deliveries() > 1000


def test_publish_records_equals_event_at_a_time_publishing():
    """orc_publish_records (complete records, as cpbus_publish_device / CPBUS_PUT_RAW take them) is the same bus as
    advance + publish / receive one event at a time: identical mailboxes, digests, debug ring and counters."""
    import numpy as np
    rng = np.random.default_rng(77)
    n_subs, n_ev = 9, 600
    masks = [0x1FFFF if rng.random() < 0.4 else int(rng.integers(0, 1 << 17)) for _ in range(n_subs)]
    a = ob.Oracle(n_subs, timers_per_sub=2); b = ob.Oracle(n_subs, timers_per_sub=2)
    for o in (a, b):
        for s, m in enumerate(masks):
            o.subscribe(m)
            o.timer_add(s, 700 + 13 * s, 500 + s, False)
            if s % 3 == 0:
                o.timer_add(s, 2000, 900 + s, True)
    rec = np.zeros(n_ev, dtype=ob.EVENT_DTYPE)
    rec["seq"] = np.arange(n_ev); rec["ts_ns"] = np.cumsum(rng.integers(0, 40, n_ev)) + 1
    rec["code"] = rng.integers(0, 17, n_ev); rec["source_id"] = rng.integers(0, 30, n_ev); rec["target"] = 0xFFFFFFFF
    uni = rng.random(n_ev) < 0.05
    rec["target"][uni] = rng.integers(0, n_subs, int(uni.sum())); rec["flags"][uni] = 2
    for lo in range(0, n_ev, 50):
        chunk = rec[lo:lo + 50]
        assert a.publish_records(chunk, int(chunk["ts_ns"][-1])) == 0
        for r in chunk:
            assert b.advance(int(r["ts_ns"])) == 0
            if r["target"] == 0xFFFFFFFF:
                assert b.publish(int(r["code"]), int(r["source_id"])) == 0
            else:
                assert b.receive(int(r["target"]), int(r["code"]), int(r["source_id"])) == 0
        assert b.advance(int(chunk["ts_ns"][-1])) == 0
    for s in range(n_subs):
        assert a.count(s) == b.count(s) and a.digest(s) == b.digest(s)
        assert a.mailbox(s).tobytes() == b.mailbox(s).tobytes()
    assert a.total_ticks() == b.total_ticks() > 0
    assert [a.published_by_code(c) for c in range(17)] == [b.published_by_code(c) for c in range(17)]
    assert a.debug_events().tobytes() == b.debug_events().tobytes()
