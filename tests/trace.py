"""Seeded synthetic op traces, and runners that feed the SAME trace to the CPU
oracle (checker) and to the CUDA bus through the C-ABI.  Test infrastructure."""
from __future__ import annotations

import numpy as np

import oracle_binding as ob

MASK_ALL = 0x1FFFF


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def zipf_masks(n_subs: int, s: float, seed: int) -> np.ndarray:
    """BASELINE config 5: each of the 16 codes 1..16 is included with p_c ∝ 1/rank(c)^s (rank = fixed permutation)."""
    rng = np.random.default_rng(seed)
    perm = np.random.default_rng(0xC0DEB205).permutation(16)      # fixed across seeds
    p = 1.0 / (np.arange(1, 17, dtype=np.float64) ** s)
    p = p / p.max()
    masks = np.zeros(n_subs, dtype=np.uint32)
    for r in range(16):
        code = 1 + int(perm[r])
        masks |= (rng.random(n_subs) < p[r]).astype(np.uint32) << np.uint32(code)
    return masks


def zipf_codes(n: int, s: float, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    perm = np.random.default_rng(0xC0DEB205).permutation(16)
    p = 1.0 / (np.arange(1, 17, dtype=np.float64) ** s)
    p /= p.sum()
    ranks = rng.choice(16, size=n, p=p)
    return (1 + perm[ranks]).astype(np.uint32)


def random_ops(seed: int, n_subs0: int, n_ops: int, timers_per_sub: int = 0, p_filter: float = 0.5,
               p_send: float = 0.02, p_adv: float = 0.3, p_member: float = 0.01, p_timer: float = 0.02,
               max_subs: int | None = None, n_sources: int = 64, dt_max: int = 5000, period_min: int = 2000,
               period_max: int = 40000, p_flush: float = 0.01, p_pairs: float = 0.0):
    """A mixed trace: subscribes/unsubscribes, publishes, direct sends, clock advances, timers."""
    rng = np.random.default_rng(seed)
    max_subs = max_subs or n_subs0 + 16
    ops = []
    active, n_total, now = [], 0, 0
    timers = []   # (sub, handle index in order of creation)
    n_timer_handles = 0
    slots = {}    # sub -> armed periodic count (to respect timers_per_sub)

    def new_mask():
        return MASK_ALL if rng.random() > p_filter else int(rng.integers(0, 1 << 17))

    def new_sub():
        m = new_mask()
        if p_pairs and rng.random() < p_pairs:   # second-level filter: exact {code, source} cases on top of a narrower mask
            m &= int(rng.integers(0, 1 << 17))
            pairs = [(int(rng.integers(0, 17)), int(rng.integers(0, n_sources))) for _ in range(int(rng.integers(1, 17)))]
            return ("sub", m, pairs)
        return ("sub", m)

    for _ in range(n_subs0):
        ops.append(new_sub())
        active.append(n_total); n_total += 1
    for _ in range(n_ops):
        r = rng.random()
        if r < p_member and n_total < max_subs:
            ops.append(new_sub()); active.append(n_total); n_total += 1
        elif r < 2 * p_member and len(active) > 1:
            s = active.pop(int(rng.integers(0, len(active))))
            ops.append(("unsub", s))
            timers = [t for t in timers if t[0] != s]
            slots.pop(s, None)
        elif r < 2 * p_member + p_timer and timers_per_sub and active:
            s = active[int(rng.integers(0, len(active)))]
            if slots.get(s, 0) < timers_per_sub:
                oneshot = bool(rng.random() < 0.3)
                period = int(rng.integers(period_min, period_max))
                ops.append(("tadd", s, period, 1000 + n_timer_handles, oneshot))
                if not oneshot:   # one-shots free their slot by themselves; keep the model simple
                    slots[s] = slots.get(s, 0) + 1
                    timers.append((s, n_timer_handles))
                else:
                    slots[s] = slots.get(s, 0) + 1   # conservative: never reuse in the generator
                n_timer_handles += 1
        elif r < 2 * p_member + 1.5 * p_timer and timers:
            s, h = timers.pop(int(rng.integers(0, len(timers))))
            ops.append(("tcancel", h))
        elif r < 2 * p_member + 1.5 * p_timer + p_send and active:
            s = active[int(rng.integers(0, len(active)))]
            ops.append(("send", s, int(rng.integers(0, 17)), int(rng.integers(0, n_sources))))
        elif r < 2 * p_member + 1.5 * p_timer + p_send + p_adv:
            now += int(rng.integers(1, dt_max))
            ops.append(("adv", now))
        elif r < 2 * p_member + 1.5 * p_timer + p_send + p_adv + p_flush:
            ops.append(("flush",))
        else:
            ops.append(("pub", int(rng.integers(0, 17)), int(rng.integers(0, n_sources))))
    return ops, n_total


def run_oracle(ops, n_max_subs, timers_per_sub=0, keep_window=0, sub_id_base=0, mailbox_cap=0):
    orc = ob.Oracle(n_max_subs, timers_per_sub=timers_per_sub, keep_window=keep_window, mailbox_cap=mailbox_cap,
                    sub_id_base=sub_id_base)
    handles = []
    for op in ops:
        k = op[0]
        if k == "sub":
            orc.subscribe(op[1], op[2] if len(op) > 2 else None)
        elif k == "unsub":
            assert orc.unsubscribe(sub_id_base + op[1]) == 0
        elif k == "pub":
            assert orc.publish(op[1], op[2]) == 0
        elif k == "send":
            assert orc.receive(sub_id_base + op[1], op[2], op[3]) == 0
        elif k == "adv":
            assert orc.advance(op[1]) == 0
        elif k == "tadd":
            handles.append(orc.timer_add(sub_id_base + op[1], op[2], op[3], op[4]))
        elif k == "tcancel":
            rc = orc.timer_cancel(handles[op[1]])
            assert rc in (0, ob.ENOENT)
        elif k == "flush":
            pass
    return orc


def run_bus(bus, ops, sub_id_base=0):
    """Drive the CUDA bus through the C-ABI with the same ops."""
    from containerpilot_b200 import _native as nat
    handles = []
    for op in ops:
        k = op[0]
        if k == "sub":
            if len(op) > 2:
                bus.subscribe_pairs(op[1], op[2])
            else:
                bus.subscribe(op[1])
        elif k == "unsub":
            bus.unsubscribe(sub_id_base + op[1])
        elif k == "pub":
            nat.check(bus.publish(op[1], op[2]), "publish")
        elif k == "send":
            nat.check(bus.send(sub_id_base + op[1], op[2], op[3]), "send")
        elif k == "adv":
            nat.check(bus.advance(op[1]), "advance")
        elif k == "tadd":
            handles.append(bus.timer_add(sub_id_base + op[1], op[2], op[3], op[4]))
        elif k == "tcancel":
            try:
                bus.timer_cancel(handles[op[1]])
            except nat.CpbusError as e:
                assert e.status == nat.ENOENT
        elif k == "flush":
            nat.check(bus.flush(), "flush")
    nat.check(bus.flush(), "flush")
    bus.sync()


def compare(bus, orc, n_total, sub_id_base=0, window=None):
    """Bit-exact comparison of every mailbox: count, order-sensitive digest, retained records."""
    got = bus.digests(sub_id_base, n_total)
    for s in range(n_total):
        gid = sub_id_base + s
        assert int(got["count"][s]) == orc.count(gid), f"count mismatch at subscriber {gid}: {int(got['count'][s])} vs {orc.count(gid)}"
        w = bus.peek_window(gid)
        o = orc.mailbox(gid)
        if window is not None:
            o = o[-window:] if len(o) > window else o
        if len(o) > len(w):
            o = o[len(o) - len(w):]
        assert w.tobytes() == o.tobytes(), f"mailbox mismatch at subscriber {gid}\n gpu={w[:8]}\n orc={o[:8]}"
        assert int(got["digest"][s]) == orc.digest(gid), f"digest mismatch at subscriber {gid}"
    st = bus.stats()
    assert st["deliveries"] == orc.total_deliveries(), (st["deliveries"], orc.total_deliveries())
    assert st["ticks"] == orc.total_ticks(), (st["ticks"], orc.total_ticks())
    return st
