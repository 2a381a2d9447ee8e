"""Diagnostic: per-chunk timing of a long device-resident run + NVML power/clock samples."""
import os, sys, time, threading, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, pynvml
from containerpilot_b200 import _native as nat
from containerpilot_b200.bus import Bus, EVENT_DTYPE
n_subs, B, steps, chunk = 65536, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 24000, 1000
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
bus = Bus(n_subs, ring_cap=1024, batch_cap=B, digest=True, stream=stream.cuda_stream)
bus.subscribe_many(np.full(n_subs, nat.MASK_ALL, dtype=np.uint32))
ev = np.zeros(B * 64, dtype=EVENT_DTYPE); ev["target"] = nat.TARGET_ALL; ev["code"] = 1 + np.arange(B * 64) % 16
dev = torch.from_numpy(ev.view(np.uint8).reshape(-1, 32)).cuda()
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples = []; stop = False
def samp():
    while not stop:
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, 0), pynvml.nvmlDeviceGetClockInfo(h, 2),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetTemperature(h, 0),
                        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.05)
th = threading.Thread(target=samp, daemon=True); th.start()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps // chunk + 1)]
torch.cuda.synchronize()
t0 = time.perf_counter(); host = []
evs[0].record(stream)
for c in range(steps // chunk):
    h0 = time.perf_counter()
    for i in range(chunk):
        nat.check(bus.publish_device(dev.data_ptr() + (i % 64) * B * 32, B, 0), "pd")
    host.append((time.perf_counter() - h0) / chunk * 1e6)
    evs[c + 1].record(stream)
torch.cuda.synchronize(); stop = True
print("chunk gpu us/step:", [round(evs[c].elapsed_time(evs[c + 1]) / chunk * 1e3, 1) for c in range(steps // chunk)])
print("chunk host us/step:", [round(x, 1) for x in host])
t_start = t0
print("nvml (t, sm, mem, W, C, reasons):", [(round(s[0] - t_start, 2), s[1], s[2], round(s[3]), s[4], hex(s[5])) for s in samples[::4]])
