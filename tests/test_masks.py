"""masks derived from the consumers' switches are semantics-preserving — a masked mailbox
contains every event the consumer would have handled, in the same order.  CPU-only (oracle as the bus)."""
import numpy as np

import oracle_binding as ob
from containerpilot_b200 import events as ev
from containerpilot_b200 import masks


def test_masks_cover_every_code_the_switch_can_match():
    names = ["global", "closed", "SIGHUP", "SIGUSR2", "myjob", "check.myjob", "other", "check.other", "db", "watch.backend",
             "myjob.heartbeat", "myjob.run-every", "myjob.wait-timeout", "mymetric|1.5", ""]
    src_id = {s: i for i, s in enumerate(sorted(set(names), key=lambda s: (s != "", s)))}   # "" -> 0
    id_src = {i: s for s, i in src_id.items()}
    consumers = [masks.JobSwitch("myjob"),
                 masks.JobSwitch("myjob", start_event=ev.Event(ev.StatusChanged, "watch.backend")),
                 masks.JobSwitch("myjob", start_event=ev.Event(ev.Stopped, "db"), stopping_wait_event=ev.Event(ev.Stopped, "other")),
                 masks.JobSwitch("other", start_event=ev.Event(ev.StatusHealthy, "watch.backend"), health_check_name="check.other"),
                 masks.MetricSwitch()]
    orc_all = ob.Oracle(len(consumers)); orc_masked = ob.Oracle(len(consumers))
    for c in consumers:
        orc_all.subscribe(0x1FFFF); orc_masked.subscribe(c.mask())
    rng = np.random.default_rng(4)
    n = 20_000
    codes = rng.integers(0, 17, n).astype(np.uint32); srcs = rng.integers(0, len(src_id), n).astype(np.uint32)
    orc_all.publish_many(codes, srcs); orc_masked.publish_many(codes, srcs)
    total_all = total_masked = 0
    for i, c in enumerate(consumers):
        def handled(box):
            return [(int(r["seq"]), int(r["code"]), int(r["source_id"])) for r in box
                    if c.handles(ev.Event(int(r["code"]), id_src[int(r["source_id"])]))]
        full, masked = orc_all.mailbox(i), orc_masked.mailbox(i)
        assert handled(full) == handled(masked), f"consumer {i} would miss events behind its mask"
        assert len(handled(full)) > 0
        total_all += len(full); total_masked += len(masked)
    assert total_masked < 0.66 * total_all          # and the filter removes real work (jobs keep 10-11 of 17 codes, metrics 3)
    assert masks.job_mask() & (1 << ev.Startup) and not masks.job_mask() & (1 << ev.Metric)
    assert masks.METRIC_MASK == (1 << ev.Metric) | (1 << ev.Shutdown) | (1 << ev.Quit)


def test_exact_cases_deliver_exactly_what_the_switch_handles():
    """with the switch's exact {code, source} cases pushed down, a mailbox holds precisely the events the
    consumer reacts to (same order), nothing else."""
    names = ["global", "closed", "SIGHUP", "SIGUSR2", "myjob", "check.myjob", "other", "check.other", "db", "watch.backend",
             "myjob.heartbeat", "myjob.run-every", "myjob.wait-timeout", "mymetric|1.5", ""]
    consumers = [masks.JobSwitch("myjob"),
                 masks.JobSwitch("myjob", start_event=ev.Event(ev.Stopped, "db"), stopping_wait_event=ev.Event(ev.Stopped, "other")),
                 masks.JobSwitch("other", start_event=ev.Event(ev.StatusHealthy, "watch.backend"), health_check_name="check.other"),
                 masks.MetricSwitch()]
    names += [e.Source for c in consumers for e in c.cases()[1]]
    src_id = {s: i for i, s in enumerate(sorted(set(names), key=lambda s: (s != "", s)))}
    id_src = {i: s for s, i in src_id.items()}
    orc_all = ob.Oracle(len(consumers)); orc_exact = ob.Oracle(len(consumers))
    for c in consumers:
        m, cases = c.cases()
        assert len(cases) <= 16
        orc_all.subscribe(0x1FFFF)
        orc_exact.subscribe(m, [(e.Code, src_id[e.Source]) for e in cases])
    rng = np.random.default_rng(5)
    n = 20_000
    codes = rng.integers(0, 17, n).astype(np.uint32); srcs = rng.integers(0, len(src_id), n).astype(np.uint32)
    orc_all.publish_many(codes, srcs); orc_exact.publish_many(codes, srcs)
    for i, c in enumerate(consumers):
        want = [(int(r["seq"]), int(r["code"]), int(r["source_id"])) for r in orc_all.mailbox(i)
                if c.handles(ev.Event(int(r["code"]), id_src[int(r["source_id"])]))]
        got = [(int(r["seq"]), int(r["code"]), int(r["source_id"])) for r in orc_exact.mailbox(i)]
        assert got == want and len(got) > 0, f"consumer {i}"
    assert orc_exact.total_deliveries() < 0.15 * orc_all.total_deliveries()
