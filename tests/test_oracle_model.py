"""Cross-check of the C oracle against an independent pure-Python restatement of the reference bus (tests/py_model.py)
on random op traces: two restatements written separately from the Go source must agree record for record.  CPU-only."""
import numpy as np
import pytest

import oracle_binding as ob
import trace as tr
from py_model import PyBus


def run_model(ops, K):
    bus = PyBus(K)
    handles = []
    for op in ops:
        k = op[0]
        if k == "sub":
            bus.subscribe(op[1], op[2] if len(op) > 2 else ())
        elif k == "unsub":
            bus.unsubscribe(op[1])
        elif k == "pub":
            bus.publish(op[1], op[2])
        elif k == "send":
            bus.receive(op[1], op[2], op[3])
        elif k == "adv":
            bus.advance(op[1])
        elif k == "tadd":
            handles.append(bus.timer_add(op[1], op[2], op[3], op[4]))
        elif k == "tcancel":
            s, slot = handles[op[1]]
            if bus.subs[s]["timers"][slot] is not None:
                bus.timer_cancel(handles[op[1]])
    return bus


@pytest.mark.parametrize("seed", range(12))
def test_c_oracle_agrees_with_python_model(seed):
    K = (0, 1, 2, 4, 8)[seed % 5]
    ops, n_total = tr.random_ops(seed + 500, 10, 2500, timers_per_sub=K, max_subs=20, p_filter=0.6, p_send=0.05, dt_max=7000)
    orc = tr.run_oracle(ops, 20, timers_per_sub=K)
    model = run_model(ops, K)
    for s in range(n_total):
        got = [tuple(int(x) for x in r) for r in orc.mailbox(s)]
        assert got == model.subs[s]["box"], f"seed {seed} subscriber {s}"
        assert orc.count(s) == len(model.subs[s]["box"])
    dbg = [(int(r["code"]), int(r["source_id"])) for r in orc.debug_events()]
    assert dbg == model.debug_events()
    assert orc.total_deliveries() == sum(len(x["box"]) for x in model.subs)


@pytest.mark.parametrize("seed", range(8))
def test_pair_filter_agrees_with_python_model(seed):
    """second-level filter (exact {code, source} cases of the consumer switch, jobs/jobs.go:188-231)"""
    K = (0, 2)[seed % 2]
    ops, n_total = tr.random_ops(seed + 900, 12, 3000, timers_per_sub=K, max_subs=24, p_filter=0.8, p_send=0.05,
                                 n_sources=5, p_pairs=0.7)
    assert any(len(op) > 2 for op in ops if op[0] == "sub")
    orc = tr.run_oracle(ops, 24, timers_per_sub=K)
    model = run_model(ops, K)
    for s in range(n_total):
        got = [tuple(int(x) for x in r) for r in orc.mailbox(s)]
        assert got == model.subs[s]["box"], f"seed {seed} subscriber {s}"
    assert orc.total_deliveries() == sum(len(x["box"]) for x in model.subs)


def test_pair_filter_semantics():
    orc = ob.Oracle(4)
    a = orc.subscribe(0, pairs=[(2, 7), (3, 9)])          # only {ExitSuccess,7} and {ExitFailed,9}
    b = orc.subscribe(1 << 2, pairs=[(3, 9)])            # every code-2 event, plus {3,9}
    c = orc.subscribe(0)                                 # nothing
    for code, src in [(2, 7), (2, 8), (3, 9), (3, 7), (4, 9)]:
        assert orc.publish(code, src) == 0
    assert [(int(r["code"]), int(r["source_id"])) for r in orc.mailbox(a)] == [(2, 7), (3, 9)]
    assert [(int(r["code"]), int(r["source_id"])) for r in orc.mailbox(b)] == [(2, 7), (2, 8), (3, 9)]
    assert orc.count(c) == 0
    assert orc.l.orc_subscribe_pairs(orc.h, 0, None, None, 17, None) == -1          # too many pairs


@pytest.mark.parametrize("seed", range(6))
def test_pair_filter_is_sandwiched_between_mask_and_code_superset(seed):
    """what makes the second level a pure optimisation: for any subscriber, mailbox(mask) is a subsequence of
    mailbox(mask, pairs), which is a subsequence of mailbox(mask | codes of the pairs) — and the middle one is exactly the
    events of the outer one whose code is in the mask or whose {code, source} is a listed case."""
    rng = np.random.default_rng(4000 + seed)
    n_subs, n_events = 12, 4000
    subs = []
    for _ in range(n_subs):
        m = int(rng.integers(0, 1 << 17)) & int(rng.integers(0, 1 << 17))
        pr = [(int(rng.integers(0, 17)), int(rng.integers(0, 6))) for _ in range(int(rng.integers(0, 17)))]
        subs.append((m, pr))
    codes = rng.integers(0, 17, n_events).astype(np.uint32); srcs = rng.integers(0, 6, n_events).astype(np.uint32)
    lo, mid, hi = ob.Oracle(n_subs), ob.Oracle(n_subs), ob.Oracle(n_subs)
    for m, pr in subs:
        sup = m
        for c, _ in pr:
            sup |= 1 << c
        lo.subscribe(m); mid.subscribe(m, pr); hi.subscribe(sup)
    for o in (lo, mid, hi):
        assert o.publish_many(codes, srcs) == 0
    for s, (m, pr) in enumerate(subs):
        key = lambda box: [(int(r["seq"]), int(r["code"]), int(r["source_id"])) for r in box]
        a, b, c = key(lo.mailbox(s)), key(mid.mailbox(s)), key(hi.mailbox(s))
        want = [r for r in c if (m >> r[1]) & 1 or (r[1], r[2]) in set(pr)]
        assert b == want
        it = iter(b)
        assert all(r in it for r in a)          # a is a subsequence of b
