"""Small-scale behaviour (BASELINE config 1 shape): latency of one Publish through the bus, 8 subscribers."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from containerpilot_b200 import _native as nat
from containerpilot_b200.bus import Bus
for lossless in (False, True):
    with Bus(8, ring_cap=1024, batch_cap=256, lossless=lossless) as bus:
        for _ in range(8):
            bus.subscribe()
        lib, h = bus._lib, bus._h
        def one(i):
            nat.check(bus.publish(1 + i % 16, i % 64), "publish"); nat.check(bus.flush(), "flush"); bus.sync()
        for i in range(200): one(i)
        t = []
        for i in range(2000):
            t0 = time.perf_counter(); one(i); t.append(time.perf_counter() - t0)
            if lossless and i % 500 == 499:
                for s in range(8): bus.drain(s)
        t = np.array(t) * 1e6
        print(f"lossless={lossless}: publish+flush+sync of ONE event to 8 mailboxes: median {np.median(t):.1f} us, p99 {np.percentile(t, 99):.1f} us")
        t0 = time.perf_counter()
        for i in range(10_000):
            rc = bus.publish(1 + i % 16, i % 64)
            if rc == nat.EAGAIN:
                for s in range(8): bus.drain(s)
                nat.check(bus.publish(1 + i % 16, i % 64), "publish")
        while bus.flush() == nat.EAGAIN:
            for s in range(8): bus.drain(s)
        bus.sync()
        dt = time.perf_counter() - t0
        print(f"lossless={lossless}: config 1 (10,000 events x 8 subscribers, staged in 256-event batches): {dt*1e3:.1f} ms = {8e4/dt:.3e} deliveries/s")
