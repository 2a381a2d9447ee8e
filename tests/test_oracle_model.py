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
            bus.subscribe(op[1])
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
