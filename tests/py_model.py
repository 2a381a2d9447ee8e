"""A second, independent restatement of the reference bus (pure Python, written from events/*.go without looking at
oracle/cpbus_oracle.c's structure) — used only to cross-check the C oracle on random traces (tests/test_oracle_model.py).
It models exactly what the parity tests compare: per-subscriber delivered sequences and the 10-slot debug ring."""
from __future__ import annotations

TIMER_EXPIRED, METRIC, TARGET_ALL, F_TICK, F_UNICAST = 8, 13, 0xFFFFFFFF, 1, 2


class PyBus:
    def __init__(self, timers_per_sub: int = 0):
        self.K = timers_per_sub
        self.subs = []            # dicts: active, mask, box (list of record tuples), timers (list of K slots or None)
        self.now = 0
        self.seq = 0
        self.buf = [(0, 0)] * 10  # events/bus.go:18-21,72-88
        self.head, self.tail = -1, 0

    # events/bus.go:105-111
    def subscribe(self, mask, pairs=()):
        self.subs.append({"active": True, "mask": mask, "pairs": set(pairs), "box": [], "timers": [None] * self.K})
        return len(self.subs) - 1

    # events/bus.go:114-122
    def unsubscribe(self, s):
        self.subs[s]["active"] = False
        self.subs[s]["timers"] = [None] * self.K

    # events/bus.go:24-31
    def _enqueue(self, ev):
        self.buf[(self.head + 1) % 10] = ev
        old = self.head
        self.head = (self.head + 1) % 10
        if old != -1 and self.head == self.tail:
            self.tail = (self.tail + 1) % 10

    # events/bus.go:34-54
    def debug_events(self):
        out = []
        while self.head != -1:
            ev = self.buf[self.tail % 10]
            if self.tail == self.head:
                self.head, self.tail = -1, 0
            else:
                self.tail = (self.tail + 1) % 10
            if ev == (0, 0):
                break
            out.append(ev)
        return out

    # events/bus.go:125-140 (+ the pushed-down consumer switch: a code mask and the switch's exact Event cases,
    # jobs/jobs.go:188-231)
    def publish(self, code, src):
        rec = (self.seq, self.now, code, src, TARGET_ALL, 0)
        for sub in self.subs:
            if sub["active"] and ((sub["mask"] >> code) & 1 or (code, src) in sub["pairs"]):
                sub["box"].append(rec)
        self._enqueue((code, src))
        self.seq += 1

    # events/subscriber.go:30-32 called directly (jobs/jobs.go:262)
    def receive(self, s, code, src):
        self.subs[s]["box"].append((self.seq, self.now, code, src, s, F_UNICAST))
        self.seq += 1

    # events/timer.go:12-71
    def timer_add(self, s, period, src, oneshot):
        slots = self.subs[s]["timers"]
        k = slots.index(None)
        slots[k] = {"due": self.now + period, "period": period, "src": src, "oneshot": oneshot, "fired": 0}
        return (s, k)

    def timer_cancel(self, handle):
        s, k = handle
        self.subs[s]["timers"][k] = None

    def advance(self, now):
        for s, sub in enumerate(self.subs):
            while True:
                due = [(t["due"], k) for k, t in enumerate(sub["timers"]) if t is not None and t["due"] <= now]
                if not due:
                    break
                _, k = min(due)
                t = sub["timers"][k]
                sub["box"].append((t["fired"], t["due"], TIMER_EXPIRED, t["src"], s, F_TICK))
                t["fired"] += 1
                if t["oneshot"]:
                    sub["timers"][k] = None
                else:
                    t["due"] += t["period"]
        self.now = now
