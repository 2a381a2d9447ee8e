"""Lossless mode (the reference's only semantics: `sub.Rx <- event` blocks on a full channel, events/subscriber.go:30-32)
at speed: the admission pass and its host sync run only when a batch does not provably fit; the exact refusal point is
unchanged.  Plus the device-side consumer and the bulk bridge."""
import numpy as np
import pytest

import oracle_binding as ob
import trace as tr
from containerpilot_b200 import _native as nat
from containerpilot_b200.bus import Bus, EVENT_DTYPE

pytestmark = pytest.mark.gpu


def test_admission_is_skipped_while_batches_provably_fit_and_refusal_stays_exact():
    """Mailbox cap 256, batches of 32 events to 3 subscribers (one filtered).  With an oracle of the same capacity: the
    GPU bus must refuse (EAGAIN) exactly when a targeted mailbox lacks the room — per batch, all-or-nothing, as before — and
    most flushes must have gone straight to the fan-out."""
    R, B = 256, 32
    masks = [nat.MASK_ALL, 1 << 2, nat.MASK_ALL]
    orc = ob.Oracle(3, keep_window=0)
    for m in masks:
        orc.subscribe(m)
    rng = np.random.default_rng(8)
    with Bus(3, ring_cap=R, batch_cap=B, lossless=True) as bus:
        bus.subscribe_many(np.array(masks, dtype=np.uint32))
        got = [[], [], []]
        n_refused = 0
        for step in range(200):
            ev = np.zeros(B, dtype=EVENT_DTYPE)
            ev["code"] = rng.integers(1, 4, B); ev["source_id"] = step * B + np.arange(B)
            nat.check(bus.publish_many(ev), "publish")
            while True:
                rc = bus.flush()
                if rc == nat.EAGAIN:                      # someone is full: consumers run (drain everything), publisher retries
                    n_refused += 1
                    used = [int(c) - sum(len(x) for x in got[s]) for s, c in enumerate(bus.digests(0, 3)["count"])]
                    assert max(u + (B if s != 1 else int((ev["code"] == 2).sum())) for s, u in enumerate(used)) > R
                    for s in range(3):
                        got[s].append(bus.drain(s, cap=R))
                    continue
                nat.check(rc, "flush"); break
            for c, s_ in zip(ev["code"], ev["source_id"]):
                orc.publish(int(c), int(s_))
            if step % 7 == 6:                             # a consumer that sometimes keeps up
                got[0].append(bus.drain(0, cap=R))
        for s in range(3):
            got[s].append(bus.drain(s, cap=R))
            assert np.concatenate(got[s]).tobytes() == orc.mailbox(s).tobytes()
        st = bus.stats()
        assert st["overwritten"] == 0 and n_refused > 3
        assert st["admit_skipped"] > st["admit_passes"] > 0       # the fast path carried most flushes


def test_device_consumer_keeps_lossless_mode_on_the_fast_path():
    """cpbus_consume_all after every flush: every mailbox is empty again, so no flush ever needs the admission pass, and
    the delivered sequences (count, digest) equal the oracle's."""
    N, B = 4096, 256
    rng = np.random.default_rng(9)
    masks = np.where(rng.random(N) < 0.5, nat.MASK_ALL, rng.integers(0, 1 << 17, N)).astype(np.uint32)
    orc = ob.Oracle(N, timers_per_sub=1, keep_window=8)
    for m in masks:
        orc.subscribe(int(m))
    with Bus(N, ring_cap=1024, batch_cap=B, lossless=True, timers_per_sub=1) as bus:
        bus.subscribe_many(masks)
        bus.timer_add_many(0, N, 50_000, source_id0=100)
        for s in range(N):
            orc.timer_add(s, 50_000, 100 + s, False)
        for step in range(30):
            now = (step + 1) * 100_000
            nat.check(bus.advance(now), "advance"); orc.advance(now)
            ev = np.zeros(B, dtype=EVENT_DTYPE)
            ev["code"] = rng.integers(0, 17, B); ev["source_id"] = rng.integers(0, 50, B)
            nat.check(bus.publish_many(ev), "publish"); nat.check(bus.flush(), "flush")
            bus.consume_all()
            orc.publish_many(ev["code"], ev["source_id"])
        bus.sync()
        got = bus.digests(0, N)
        assert (got["count"] == np.array([orc.count(s) for s in range(N)], dtype=np.uint64)).all()
        assert (got["digest"] == np.array([orc.digest(s) for s in range(N)], dtype=np.uint64)).all()
        st = bus.stats()
        assert st["admit_passes"] == 0 and st["admit_skipped"] == 30
        assert len(bus.drain(5)) == 0                               # everything was consumed on the device


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_flush_blocks_per_event_like_the_go_bus(seed):
    """The Go publisher stalls at the FIRST event a full channel cannot take (events/subscriber.go:30-32); everything in front
    of it has been delivered.  The oracle models exactly that (mailbox_cap: orc_publish refuses one event at a time).  The
    GPU bus publishes whole batches: when cpbus_flush returns EAGAIN, every mailbox must hold exactly what the oracle's
    mailboxes hold at ITS stall point — the longest prefix every targeted mailbox could take — and the next flush continues
    with the first undelivered event."""
    R, B, N = 128, 64, 6
    rng = np.random.default_rng(100 + seed)
    masks = [nat.MASK_ALL, 1 << 2, (1 << 3) | (1 << 2), nat.MASK_ALL, 1 << 5, 0]
    orc = ob.Oracle(N, keep_window=0, mailbox_cap=R)
    for m_ in masks:
        orc.subscribe(m_)
    with Bus(N, ring_cap=R, batch_cap=B, lossless=True) as bus:
        bus.subscribe_many(np.array(masks, dtype=np.uint32))
        got = [[] for _ in range(N)]
        n_partial = 0
        for step in range(60):
            ev = np.zeros(B, dtype=EVENT_DTYPE)
            ev["code"] = rng.integers(1, 7, B); ev["source_id"] = step * B + np.arange(B)
            nat.check(bus.publish_many(ev), "publish")
            i = 0
            while True:
                rc = bus.flush()
                while i < B:                                          # the oracle publishes event by event until it blocks
                    r = orc.publish(int(ev["code"][i]), int(ev["source_id"][i]))
                    if r == ob.EAGAIN:
                        break
                    assert r == 0
                    i += 1
                counts = bus.digests(0, N)["count"]
                assert [int(c) for c in counts] == [orc.count(s) for s in range(N)], (step, i, rc)   # same delivered sets at the stall point
                if rc == nat.OK:
                    assert i == B
                    break
                assert rc == nat.EAGAIN and i < B
                n_partial += 1 if i > 0 else 0
                for s in rng.permutation(N)[:3]:                      # some consumers run (not necessarily the full one)
                    take = int(rng.integers(1, R + 1))
                    g = bus.drain(int(s), cap=take)
                    o = orc.consume(int(s), take)
                    assert g.tobytes() == o.tobytes()
                    got[int(s)].append(g)
        for s in range(N):
            g = bus.drain(s, cap=R); o = orc.consume(s, R)
            assert g.tobytes() == o.tobytes()
        st = bus.stats()
        assert st["admit_partial"] > 0 and n_partial > 0 and st["overwritten"] == 0
