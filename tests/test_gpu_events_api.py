"""The reference's own bus tests, re-stated against the mirror of the `events`
API running on the CUDA bus (containerpilot_b200.events).  Each test cites the
Go test it follows."""
import json
import os
from collections import Counter

import pytest

from containerpilot_b200 import events
from containerpilot_b200.events import Event

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))


class TestPublisher(events.Publisher):              # events/events_test.go:13-21
    __test__ = False

    def __init__(self, bus):
        super().__init__()
        self.Register(bus)


class TestSubscriber(events.Subscriber):            # events/events_test.go:23-37
    __test__ = False

    def __init__(self):
        super().__init__(events.Chan(100))
        self.results = []

    def Run(self, bus):                             # events_test.go:39-61 (the goroutine body, run on demand)
        self.Subscribe(bus)

    def Stop(self):
        self.results += self.Received()
        self.Unsubscribe()
        self.Rx.close()


def test_pub_sub_interfaces():
    """events/events_test.go:64-89"""
    bus = events.NewEventBus()
    tp = TestPublisher(bus)
    ts = TestSubscriber()
    ts.Run(bus)
    expected = [Event(events.Startup, "serviceA")]
    for e in expected:
        tp.Publish(e)
    ts.Stop()
    assert bus.DebugEvents() == expected
    assert ts.results == expected               # the Go test collects but never asserts this; we do
    tp.Unregister()
    assert bus.Wait() is False
    bus.close()


def test_publish_signal():
    """events/events_test.go:91-113"""
    bus = events.NewEventBus()
    ts = TestSubscriber()
    ts.Run(bus)
    signals = ["SIGHUP", "SIGUSR2"]
    expected = [Event(events.Signal, s) for s in signals]
    for s in signals:
        bus.PublishSignal(s)
    ts.Stop()
    assert bus.DebugEvents() == expected and ts.results == expected
    bus.close()


def test_job_run_safe_close_sequence():
    """jobs/jobs_test.go:15-48: publish after the only subscriber unsubscribed must not panic."""
    bus = events.NewEventBus()
    job_sub, job_pub = events.Subscriber(events.Chan(1000)), events.Publisher()
    job_sub.Subscribe(bus); job_pub.Register(bus)
    bus.Publish(events.GlobalStartup)
    job_pub.Publish(Event(events.Stopping, "myjob"))      # jobs/jobs.go:390
    job_sub.Unsubscribe(); job_pub.Unregister()           # jobs/jobs.go:411-412
    bus.Publish(Event(events.Stopped, "myjob"))           # jobs/jobs.go:415
    assert bus.Wait() is False
    assert bus.DebugEvents() == [events.GlobalStartup, Event(events.Stopping, "myjob"), Event(events.Stopped, "myjob")]
    job_pub.Bus.Publish(events.GlobalStartup)             # must not raise
    assert job_sub.Received() == [events.GlobalStartup, Event(events.Stopping, "myjob")]
    bus.close()


@pytest.mark.parametrize("vec", G["multiset"], ids=lambda v: v["name"])
def test_multiset_vectors(vec):
    bus = events.NewEventBus()
    sub = events.Subscriber(events.Chan(1000)); sub.Subscribe(bus)
    for code, src in vec.get("direct_receives", []):
        sub.Receive(Event(code, src))                     # watches/watches_test.go:48-50
    for code, src in vec["published"]:
        bus.Publish(Event(code, src))
    got = Counter(f"{e.Code}|{e.Source}" for e in bus.DebugEvents())
    assert dict(got) == vec["debug_events"]
    want = [Event(c, s) for c, s in vec.get("direct_receives", [])] + [Event(c, s) for c, s in vec["published"]]
    assert sub.Received() == want
    bus.close()


def test_closed_mailbox_panics_and_double_unsubscribe_panics():
    """events/bus.go:135-137 (send on closed Rx panics) and :121 (negative WaitGroup)."""
    bus = events.NewEventBus()
    sub = events.Subscriber(events.Chan(10)); sub.Subscribe(bus)
    sub.Rx.close()
    with pytest.raises(events.BusPanic):
        bus.Publish(events.GlobalStartup)
    sub.Unsubscribe()
    with pytest.raises(events.BusPanic):
        sub.Unsubscribe()
    with pytest.raises(events.BusPanic):
        bus.Subscribe(object())                           # bus.go:108 type assertion
    bus.close()


def test_timers_deliver_to_owner_only():
    """events/timer.go:12-71 under the virtual clock; names as in jobs/jobs.go:147-158."""
    bus = events.NewEventBus()
    job = events.Subscriber(events.Chan(1000)); job.Subscribe(bus)
    other = events.Subscriber(events.Chan(1000)); other.Subscribe(bus)
    ctx, cancel = events.WithCancel()
    events.NewEventTimer(ctx, job.Rx, 1_000_000_000, "myjob.heartbeat")
    events.NewEventTimeout(ctx, job.Rx, 2_500_000_000, "myjob.wait-timeout")
    bus.Publish(events.GlobalStartup)
    bus.Advance(3_000_000_000)
    bus.Publish(Event(events.StatusHealthy, "myjob"))
    hb, to = Event(events.TimerExpired, "myjob.heartbeat"), Event(events.TimerExpired, "myjob.wait-timeout")
    assert job.Received() == [events.GlobalStartup, hb, hb, to, hb, Event(events.StatusHealthy, "myjob")]
    assert other.Received() == [events.GlobalStartup, Event(events.StatusHealthy, "myjob")]
    cancel()                                              # ctx.Done(): timer.go:57-58
    bus.Advance(10_000_000_000)
    bus.Publish(events.GlobalShutdown)
    assert job.Received() == [events.GlobalShutdown]
    assert bus.DebugEvents() == [events.GlobalStartup, Event(events.StatusHealthy, "myjob"), events.GlobalShutdown]
    bus.close()


def test_from_string_and_names():
    for name, code in G["from_string"]["accepted"].items():
        assert events.FromString(name) == (code, None)
    code, err = events.FromString("bogus")
    assert code == events.None_ and "bogus is not a valid event code" in str(err)      # events.go:85
    assert [events.CodeString(i) for i in range(17)] == G["code_names"]["names"]
    assert events.CodeString(42) == "EventCode(42)"                                    # eventcode_string.go:11


def test_cpp_mirror_restates_reference_tests():
    """The compiled-language host side (containerpilot_b200/csrc/host/events.hpp, C++17 over the C-ABI):
    events_test.cc restates events/events_test.go + the jobs/control/watches vectors, with real
    bounded channels and consumer threads."""
    import subprocess
    exe = os.path.join(os.path.dirname(HERE), "containerpilot_b200", "csrc", "host", "events_test")
    assert os.path.exists(exe), "build with __graft_entry__.build()"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("PASS"), r.stdout + r.stderr


@pytest.mark.parametrize("body,status,expected", [
    ("{{\n", 422, {}),
    ('{"mymetric": 1.0}', 200, {Event(events.Metric, "mymetric|1"): 1}),
    ('{"mymetric": 1.5, "myothermetric": 2}', 200,
     {Event(events.Metric, "mymetric|1.5"): 1, Event(events.Metric, "myothermetric|2"): 1}),
])
def test_post_metric_as_one_batch(body, status, expected):
    """control/endpoints_test.go:104-145 through the batched handler (containerpilot_b200/ingest.py)"""
    from containerpilot_b200 import ingest
    bus = events.NewEventBus()
    sensor = events.Subscriber(events.Chan(1000))
    sensor.Subscribe(bus, 1 << events.Metric)
    _, got_status = ingest.post_metric(bus, body)
    assert got_status == status
    got = Counter(e for e in bus.DebugEvents() if e != events.GlobalStartup)
    assert dict(got) == expected
    assert Counter(sensor.Received()) == Counter(expected)
    sensor.Unsubscribe()
    bus.close()


def test_post_metric_burst_is_one_fan_out():
    """a 200-key request costs one staging pass and one launch instead of 200 publishes"""
    import json as _json
    from containerpilot_b200 import ingest
    bus = events.NewEventBus(n_max_subs=8, batch_cap=256)
    subs = [events.Subscriber(events.Chan(1000)) for _ in range(4)]
    for s in subs:
        s.Subscribe(bus)
    body = _json.dumps({f"m{i}": i / 4 for i in range(200)})
    before = bus._bus.stats()["batches"]
    assert ingest.post_metric(bus, body) == (None, 200)
    bus.Flush()
    assert bus._bus.stats()["batches"] - before == 1
    want = [Event(events.Metric, f"m{i}|{ingest.go_sprint_v(i / 4)}") for i in range(200)]
    for s in subs:
        assert s.Received() == want
        s.Unsubscribe()
    bus.close()
