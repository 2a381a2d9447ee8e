#!/usr/bin/env python
"""Kernel A/B on one box: device-resident fan-out timing of configs 2/3/5 through the OLDEST common C-ABI subset
(cpbus_create / subscribe_many / timer_add_many / publish_device / stats), so that library builds of different ABI
versions (incl. round 1's) can be compared.  usage: ab_bench.py <lib.so> [<lib.so> ...]   (prints one line per lib x config)"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import trace as tr  # noqa: E402


class Config(C.Structure):
    _fields_ = [("n_max_subs", C.c_uint32), ("ring_cap", C.c_uint32), ("batch_cap", C.c_uint32), ("timers_per_sub", C.c_uint32),
                ("flags", C.c_uint32), ("device", C.c_int32), ("sub_id_base", C.c_uint32), ("store_path", C.c_uint32),
                ("stream", C.c_void_p), ("grid_ctas", C.c_uint32), ("reserved", C.c_uint32 * 5)]


def run(lib_path, name, n_subs, timers, zipf, steps=60, warm=30, B=512):
    lib = C.CDLL(lib_path)
    lib.cpbus_publish_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64]
    lib.cpbus_timer_add_many.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint32, C.c_int]
    lib.cpbus_subscribe_many.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    cfg = Config(); cfg.n_max_subs, cfg.ring_cap, cfg.batch_cap, cfg.timers_per_sub, cfg.flags, cfg.device = n_subs, 1024, B, timers, 2, 0
    cfg.stream = C.c_void_p(stream.cuda_stream)
    cfg.store_path = int(os.environ.get('AB_STORE', '0'))
    h = C.c_void_p()
    assert lib.cpbus_create(C.byref(cfg), C.byref(h)) == 0
    masks = tr.zipf_masks(n_subs, zipf, 0xC0DEB205) if zipf else np.full(n_subs, 0x1FFFF, dtype=np.uint32)
    assert lib.cpbus_subscribe_many(h, masks.ctypes.data, n_subs, None) == 0
    if timers:
        assert lib.cpbus_timer_add_many(h, 0, n_subs, 1_000_000, None, 1_000_000, 0) == 0
    nb = warm + steps + 4
    n_ev = nb * B
    ev = np.zeros(n_ev, dtype=[("seq", "<u8"), ("ts", "<u8"), ("code", "<u4"), ("src", "<u4"), ("target", "<u4"), ("flags", "<u4")])
    ev["seq"] = np.arange(n_ev); ev["ts"] = (np.arange(n_ev) + 1) * 10_000
    rng = np.random.default_rng(2)
    ev["code"] = tr.zipf_codes(n_ev, zipf, 0xC0DEB205) if zipf else rng.integers(1, 17, n_ev)
    ev["src"] = rng.integers(0, 4096, n_ev); ev["target"] = 0xFFFFFFFF
    dev = torch.from_numpy(ev.view(np.uint8).reshape(-1, 32)).cuda()
    step = [0]

    def go(k):
        for _ in range(k):
            i = step[0]
            assert lib.cpbus_publish_device(h, C.c_void_p(dev.data_ptr() + i * B * 32), B, (i + 1) * B * 10_000) == 0
            step[0] += 1
    go(warm); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); go(steps); e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    lib.cpbus_destroy.argtypes = [C.c_void_p]; lib.cpbus_destroy(h)
    print(f"{os.path.basename(lib_path):34s} {name:8s} {ms * 1e3:9.2f} us/step", flush=True)
    del dev; torch.cuda.empty_cache()


if __name__ == "__main__":
    for rep in range(int(os.environ.get("AB_REPS", "2"))):
        which = os.environ.get("AB_CONFIGS", "config2,config3,config5").split(",")
        for lib in sys.argv[1:]:
            if "config2" in which:
                run(os.path.abspath(lib), "config2", 65_536, 0, None, steps=200, warm=100)
            if "config3" in which:
                run(os.path.abspath(lib), "config3", 1_048_576, 1, None)
            if "config5" in which:
                run(os.path.abspath(lib), "config5", int(os.environ.get("AB_SUBS5", "1048576")), 0, 1.0)
