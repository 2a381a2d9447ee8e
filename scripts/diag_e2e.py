"""Diagnostic: where does the e2e step time go? (host buffers -> cpbus_publish -> flush -> result read)"""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from containerpilot_b200 import _native as nat
from containerpilot_b200.bus import Bus, EVENT_DTYPE
n_subs, B, K = 65536, 256, 3000
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
bus = Bus(n_subs, ring_cap=1024, batch_cap=B, digest=True, stream=stream.cuda_stream)
bus.subscribe_many(np.full(n_subs, nat.MASK_ALL, dtype=np.uint32))
host = np.zeros(B * 64, dtype=EVENT_DTYPE); host["code"] = 1 + np.arange(B * 64) % 16
lib, h = bus._lib, bus._h
ptrs = [host[i * B:(i + 1) * B].ctypes.data for i in range(64)]
def run(name, body):
    for i in range(50): body(i)
    bus.sync(); t0 = time.perf_counter()
    for i in range(K): body(50 + i)
    t_host = time.perf_counter() - t0; bus.sync(); t = time.perf_counter() - t0
    print(f"{name:50s} host {t_host / K * 1e6:7.1f} us/step   total {t / K * 1e6:7.1f} us/step")
now = [0]
def adv():
    now[0] += 2_560_000; lib.cpbus_advance(h, now[0])
def a(i): adv(); bus.publish_many(host[(i % 64) * B:(i % 64 + 1) * B]); bus.flush()
def b(i): adv(); lib.cpbus_publish(h, ptrs[i % 64], B); lib.cpbus_flush(h)
tk = []
def c(i):
    adv(); lib.cpbus_publish(h, ptrs[i % 64], B); lib.cpbus_flush(h)
    tk.append(bus.digest_fold_begin(0, n_subs))
    if len(tk) > 2: bus.digest_fold_end(tk.pop(0))
def d(i):
    adv(); lib.cpbus_publish(h, ptrs[i % 64], B); lib.cpbus_flush(h); bus.digest_fold(0, n_subs)
run("publish_many(numpy slice)+flush", a)
run("cpbus_publish(ptr)+flush", b)
run("  + digest_fold_begin/_end (2 deep)", c)
while tk: bus.digest_fold_end(tk.pop(0))
run("  + digest_fold (sync every step)", d)
