"""Subscription masks derived from the consumers' own switches.

The reference bus delivers every event to every subscriber and each consumer drops what its
`switch` does not match.  Pushing that switch down as a code mask must not change behaviour,
so a mask has to cover EVERY code a consumer can react to:

  Job     jobs/jobs.go:187-234 (processEvent), :163-184 (Run: QuitByTest), :388-406 (cleanup)
  Metric  telemetry/metrics.go:92-106
  Watch   not a bus subscriber (private rx + its own timer, watches/watches.go:65-100)

Source matching (`{ExitSuccess, "check.myjob"}` vs any other ExitSuccess): the mask is a superset
filter on the code only.  Every case of these switches compares a whole Event value, so the exact
filter is a short list of {code, source} cases (row N3): `JobSwitch.cases()` / `MetricSwitch.cases()`
feed `EventBus.Subscribe(sub, mask, cases)` -> `cpbus_subscribe_pairs`, and the mailbox then holds
only what the consumer handles.  Timer ticks and direct sends are unicast and bypass both levels,
exactly like a direct channel write.
"""
from __future__ import annotations

from . import events as ev


def _bits(*codes: int) -> int:
    m = 0
    for c in codes:
        m |= 1 << c
    return m


# codes every Job's processEvent can match, whatever its config (jobs/jobs.go:206-227, :174)
JOB_BASE_MASK = _bits(ev.ExitFailed, ev.ExitSuccess,      # health check + own exec exit
                      ev.Quit, ev.Shutdown,               # {Quit, name}, QuitByTest, GlobalShutdown
                      ev.EnterMaintenance, ev.ExitMaintenance,
                      ev.Signal)                          # SIGHUP / SIGUSR2

METRIC_MASK = _bits(ev.Metric, ev.Shutdown, ev.Quit)      # telemetry/metrics.go:97-104


def job_mask(start_event_code: int | None = None, stopping_wait_code: int | None = None,
             has_stopping_timeout: bool = False, has_timers: bool = True) -> int:
    """Mask for a Job.  `start_event_code` is the code of its `when` event (any name FromString accepts,
    jobs/config.go:225-244; None = starts on GlobalStartup => Startup); `stopping_wait_code` is the code
    of `stoppingWaitEvent` (jobs/jobs.go:391-406), if configured."""
    m = JOB_BASE_MASK
    if has_timers:
        # heartbeat / run-every / wait-timeout ticks arrive unicast (they bypass the mask), but processEvent would
        # also match a BROADCAST {TimerExpired, "<job>.heartbeat"}; nobody publishes one (events/timer.go writes rx
        # directly), so keeping the bit costs nothing and keeps the mask a strict superset of the switch
        m |= 1 << ev.TimerExpired
    m |= 1 << (ev.Startup if start_event_code is None else start_event_code)
    if stopping_wait_code is not None:
        m |= 1 << stopping_wait_code
    if has_stopping_timeout:
        m |= 1 << ev.Stopping     # cleanup also waits for {Stopping, "<job>.stopping-timeout"} (jobs/jobs.go:403)
    return m


class JobSwitch:
    """The matching half of Job.processEvent + Run + cleanup, as a predicate: would this Job react to `event`?
    Used by the tests to show that a masked mailbox loses nothing the consumer would have handled."""

    def __init__(self, name: str, start_event: ev.Event = ev.GlobalStartup, health_check_name: str | None = None,
                 stopping_wait_event: ev.Event | None = None, has_stopping_timeout: bool = False):
        self.name = name
        self.start_event = start_event
        self.health = health_check_name or f"check.{name}"
        self.stopping_wait_event = stopping_wait_event
        # cleanup also selects on {Stopping, "<job>.stopping-timeout"} (jobs/jobs.go:403) when a stopping timeout is
        # configured.  (The timer that is meant to produce it emits code TimerExpired, events/timer.go:31 — the
        # reference's own quirk, SURVEY §3.4 — but the switch case exists, so the filter must let such an event through.)
        self.stopping_timeout_event = ev.Event(ev.Stopping, f"{name}.stopping-timeout") if has_stopping_timeout else None

    def mask(self) -> int:
        return job_mask(self.start_event.Code, self.stopping_wait_event.Code if self.stopping_wait_event else None,
                        has_stopping_timeout=self.stopping_timeout_event is not None)

    def cases(self) -> tuple[int, list]:
        """(mask, cases) for the exact second-level filter: every case of the switch as a whole Event value (at most
        16 = CPBUS_MAX_PAIRS).  The three TimerExpired cases normally arrive unicast; they are listed so that a
        broadcast one would still be delivered, like the switch would match it."""
        n = self.name
        cs = [ev.Event(ev.TimerExpired, f"{n}.heartbeat"), ev.Event(ev.TimerExpired, f"{n}.run-every"),
              ev.Event(ev.TimerExpired, f"{n}.wait-timeout"),
              ev.Event(ev.ExitFailed, self.health), ev.Event(ev.ExitSuccess, self.health),
              ev.Event(ev.Quit, n), ev.GlobalShutdown, ev.QuitByTest,
              ev.GlobalEnterMaintenance, ev.GlobalExitMaintenance,
              ev.Event(ev.ExitSuccess, n), ev.Event(ev.ExitFailed, n),
              ev.Event(ev.Signal, "SIGHUP"), ev.Event(ev.Signal, "SIGUSR2"), self.start_event]
        if self.stopping_wait_event is not None:
            cs.append(self.stopping_wait_event)
        if self.stopping_timeout_event is not None:
            cs.append(self.stopping_timeout_event)
        out = []
        for c in cs:
            if c not in out:
                out.append(c)
        mask = 0
        while len(out) > 16:            # CPBUS_MAX_PAIRS: a case that does not fit widens the code mask instead (still a superset)
            mask |= 1 << out.pop().Code
        return mask, [c for c in out if not (mask >> c.Code) & 1]

    def handles(self, e: ev.Event) -> bool:
        n = self.name
        return (e in (ev.Event(ev.TimerExpired, f"{n}.heartbeat"), ev.Event(ev.TimerExpired, f"{n}.run-every"),
                      ev.Event(ev.TimerExpired, f"{n}.wait-timeout"),
                      ev.Event(ev.ExitFailed, self.health), ev.Event(ev.ExitSuccess, self.health),
                      ev.Event(ev.Quit, n), ev.GlobalShutdown, ev.QuitByTest,
                      ev.GlobalEnterMaintenance, ev.GlobalExitMaintenance,
                      ev.Event(ev.ExitSuccess, n), ev.Event(ev.ExitFailed, n),
                      ev.Event(ev.Signal, "SIGHUP"), ev.Event(ev.Signal, "SIGUSR2"), self.start_event)
                or (self.stopping_wait_event is not None and e == self.stopping_wait_event)
                or (self.stopping_timeout_event is not None and e == self.stopping_timeout_event))


class MetricSwitch:
    def mask(self) -> int:
        return METRIC_MASK

    def cases(self) -> tuple[int, list]:
        return 1 << ev.Metric, [ev.GlobalShutdown, ev.QuitByTest]     # any Metric source; the two exact stop events

    def handles(self, e: ev.Event) -> bool:
        return e.Code == ev.Metric or e in (ev.GlobalShutdown, ev.QuitByTest)
