"""Pin the CPU oracle against every known-answer vector the reference's own
tests hold for the bus (tests/golden/reference_vectors.json, each entry cites
its file:line in /root/reference).  CPU-only."""
import json
import os
from collections import Counter

import pytest

import oracle_binding as ob

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_vectors.json")) as f:
    G = json.load(f)


class Interner:
    def __init__(self):
        self.ids, self.names = {"": 0}, [""]

    def __call__(self, s):
        if s not in self.ids:
            self.ids[s] = len(self.names)
            self.names.append(s)
        return self.ids[s]


def test_code_names_match_stringer_table():
    names, blob, idx = G["code_names"]["names"], G["code_names"]["stringer_blob"], G["code_names"]["stringer_index"]
    assert len(names) == 17
    for i, n in enumerate(names):
        assert blob[idx[i]:idx[i + 1]] == n                      # the generated table itself
        assert ob.lib().orc_code_name(i).decode() == n
    assert ob.lib().orc_code_name(17) is None and ob.lib().orc_code_name(-1) is None


def test_from_string_vectors():
    for name, code in G["from_string"]["accepted"].items():
        assert ob.lib().orc_code_from_string(name.encode()) == code, name
    for name in G["from_string"]["rejected"]:
        assert ob.lib().orc_code_from_string(name.encode()) == -1, name


def play(vec, n_subs=1):
    it = Interner()
    orc = ob.Oracle(8)
    subs = [orc.subscribe() for _ in range(vec.get("subscribers", n_subs))]
    for s, (code, src) in zip(subs * 8, vec.get("direct_receives", [])):
        assert orc.receive(s, code, it(src)) == 0
    for e in vec["published"]:
        if e == "UNSUBSCRIBE_ALL":
            for s in subs:
                assert orc.unsubscribe(s) == 0
            continue
        assert orc.publish(e[0], it(e[1])) == 0
    dbg = [(int(r["code"]), it.names[int(r["source_id"])]) for r in orc.debug_events()]
    return orc, it, subs, dbg


@pytest.mark.parametrize("vec", G["ordered"], ids=lambda v: v["name"])
def test_ordered_debug_events(vec):
    orc, it, subs, dbg = play(vec)
    assert dbg == [tuple(e) for e in vec["debug_events"]]
    if "then_publish_must_not_panic" in vec:                       # jobs/jobs_test.go:33-38
        c, s = vec["then_publish_must_not_panic"]
        assert orc.publish(c, it(s)) == 0


@pytest.mark.parametrize("vec", G["multiset"], ids=lambda v: v["name"])
def test_multiset_debug_events(vec):
    orc, it, subs, dbg = play(vec)
    got = Counter(f"{c}|{s}" for c, s in dbg)
    assert dict(got) == vec["debug_events"]
    # direct Receive()s never reach the debug ring but do reach the mailbox (watches_test.go:48-50)
    if "direct_receives" in vec:
        box = orc.mailbox(subs[0])
        assert [(int(r["code"]), it.names[int(r["source_id"])]) for r in box[:3]] == [tuple(e) for e in vec["direct_receives"]]


def test_debug_ring_keeps_last_ten_and_stops_at_nonevent():
    """events/bus.go:24-54 followed to the letter (unpinned by reference tests: they publish <= 10 events)."""
    orc = ob.Oracle(1)
    for i in range(1, 24):
        orc.publish(1 + i % 16, i)
    dbg = orc.debug_events()
    assert [int(r["source_id"]) for r in dbg] == list(range(14, 24))     # last 10, oldest first
    assert len(orc.debug_events()) == 0                                  # drained: head=-1, tail=0
    orc.publish(1, 5); orc.publish(0, 0); orc.publish(2, 6)              # NonEvent in the middle
    dbg = orc.debug_events()
    assert [(int(r["code"]), int(r["source_id"])) for r in dbg] == [(1, 5)]   # early stop at NonEvent (bus.go:48-50)
    dbg = orc.debug_events()                                             # ...which leaves the rest for the next call
    assert [(int(r["code"]), int(r["source_id"])) for r in dbg] == [(2, 6)]


def test_waitgroup_accounting():
    """bus.go:91-122,164-169: Register/Subscribe add, Unregister/Unsubscribe done, Wait returns reload."""
    orc = ob.Oracle(4)
    orc.register(); s = orc.subscribe()
    assert orc.wait() == -1
    assert orc.unsubscribe(s) == 0 and orc.unregister() == 0
    assert orc.wait() == 0
    orc.set_reload()
    assert orc.wait() == 1
    assert orc.unsubscribe(s) == ob.ECLOSED        # second Unsubscribe: negative WaitGroup => Go panics
