"""ctypes binding of oracle/libcpbus_oracle.so — the CHECKER.  Test infrastructure
only: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "libcpbus_oracle.so")
EVENT_DTYPE = np.dtype([("seq", "<u8"), ("ts_ns", "<u8"), ("code", "<u4"), ("source_id", "<u4"),
                        ("target", "<u4"), ("flags", "<u4")])
EAGAIN, ENOSPC, ENOENT, ECLOSED = -4, -5, -6, -7
_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(LIB)
        vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
        sig = {
            "orc_new": (vp, [u32, u32, u32, u32, u32]), "orc_free": (None, [vp]),
            "orc_subscribe": (C.c_int, [vp, u32, C.POINTER(u32)]),
            "orc_subscribe_pairs": (C.c_int, [vp, u32, vp, vp, u32, C.POINTER(u32)]), "orc_unsubscribe": (C.c_int, [vp, u32]),
            "orc_register": (C.c_int, [vp]), "orc_unregister": (C.c_int, [vp]), "orc_set_reload": (None, [vp]),
            "orc_wait": (C.c_int, [vp]), "orc_publish": (C.c_int, [vp, u32, u32]),
            "orc_publish_many": (C.c_int, [vp, vp, vp, C.c_size_t, u64]),
            "orc_publish_records": (C.c_int, [vp, vp, C.c_size_t, u64]),
            "orc_receive": (C.c_int, [vp, u32, u32, u32]), "orc_advance": (C.c_int, [vp, u64]),
            "orc_timer_add": (C.c_int, [vp, u32, u64, u32, C.c_int, C.POINTER(u32)]),
            "orc_timer_cancel": (C.c_int, [vp, u32]), "orc_debug_events": (C.c_size_t, [vp, vp, C.c_size_t]),
            "orc_count": (u64, [vp, u32]), "orc_digest": (u64, [vp, u32]),
            "orc_mailbox": (C.c_size_t, [vp, u32, vp, C.c_size_t]), "orc_consume": (C.c_size_t, [vp, u32, vp, C.c_size_t]),
            "orc_now": (u64, [vp]), "orc_total_deliveries": (u64, [vp]), "orc_total_ticks": (u64, [vp]),
            "orc_published_by_code": (u64, [vp, u32]), "orc_code_name": (C.c_char_p, [C.c_int]),
            "orc_code_from_string": (C.c_int, [C.c_char_p]), "orc_record_hash": (u64, [vp]),
            "orc_digest_multiplier": (u64, []),
            "gobus_bench": (C.c_double, [u32, u32, u32, u32, C.POINTER(u64)]),
            "gobus_bench_steps": (C.c_double, [u32, u32, u32, u32, u32, u32, C.POINTER(C.c_double)]),
            "gobus_bench_steps2": (C.c_double, [u32, u32, u32, u32, u32, u32, C.c_int, vp]),
        }
        for name, (res, args) in sig.items():
            f = getattr(l, name)
            f.restype, f.argtypes = res, args
        _lib = l
    return _lib


class Oracle:
    def __init__(self, n_max_subs, timers_per_sub=0, keep_window=0, mailbox_cap=0, sub_id_base=0):
        self.l = lib()
        self.h = C.c_void_p(self.l.orc_new(n_max_subs, timers_per_sub, keep_window, mailbox_cap, sub_id_base))
        self.keep_window = keep_window

    def __del__(self):
        try:
            if self.h:
                self.l.orc_free(self.h)
                self.h = None
        except Exception:
            pass

    def subscribe(self, mask=0x1FFFF, pairs=None):
        out = C.c_uint32()
        if pairs:
            codes = np.ascontiguousarray([p[0] for p in pairs], dtype=np.uint32)
            srcs = np.ascontiguousarray([p[1] for p in pairs], dtype=np.uint32)
            rc = self.l.orc_subscribe_pairs(self.h, mask, codes.ctypes.data, srcs.ctypes.data, len(pairs), C.byref(out))
        else:
            rc = self.l.orc_subscribe(self.h, mask, C.byref(out))
        assert rc == 0, rc
        return out.value

    def unsubscribe(self, sub):
        return self.l.orc_unsubscribe(self.h, sub)

    def register(self):
        return self.l.orc_register(self.h)

    def unregister(self):
        return self.l.orc_unregister(self.h)

    def set_reload(self):
        self.l.orc_set_reload(self.h)

    def wait(self):
        return self.l.orc_wait(self.h)

    def publish(self, code, source_id=0):
        return self.l.orc_publish(self.h, code, source_id)

    def publish_many(self, codes, sources, dt_ns=0):
        c = np.ascontiguousarray(codes, dtype=np.uint32)
        s = np.ascontiguousarray(sources, dtype=np.uint32)
        return self.l.orc_publish_many(self.h, c.ctypes.data, s.ctypes.data, c.size, dt_ns)

    def publish_records(self, records, watermark_ns=0):
        """complete 32-byte records (cpbus_publish_device / CPBUS_PUT_RAW semantics)"""
        r = np.ascontiguousarray(records, dtype=EVENT_DTYPE)
        return self.l.orc_publish_records(self.h, r.ctypes.data, r.size, watermark_ns)

    def receive(self, sub, code, source_id=0):
        return self.l.orc_receive(self.h, sub, code, source_id)

    def advance(self, now_ns):
        return self.l.orc_advance(self.h, now_ns)

    def timer_add(self, sub, period_ns, source_id, oneshot=False):
        out = C.c_uint32()
        rc = self.l.orc_timer_add(self.h, sub, period_ns, source_id, int(oneshot), C.byref(out))
        assert rc == 0, rc
        return out.value

    def timer_cancel(self, tid):
        return self.l.orc_timer_cancel(self.h, tid)

    def debug_events(self):
        out = np.zeros(16, dtype=EVENT_DTYPE)
        n = self.l.orc_debug_events(self.h, out.ctypes.data, 16)
        return out[:n]

    def count(self, sub):
        return self.l.orc_count(self.h, sub)

    def digest(self, sub):
        return self.l.orc_digest(self.h, sub)

    def mailbox(self, sub):
        cap = int(self.count(sub)) if self.keep_window == 0 else self.keep_window
        out = np.zeros(max(cap, 1), dtype=EVENT_DTYPE)
        n = self.l.orc_mailbox(self.h, sub, out.ctypes.data, cap)
        return out[:n]

    def consume(self, sub, cap):
        out = np.zeros(max(cap, 1), dtype=EVENT_DTYPE)
        n = self.l.orc_consume(self.h, sub, out.ctypes.data, cap)
        return out[:n]

    def now(self):
        return self.l.orc_now(self.h)

    def total_deliveries(self):
        return self.l.orc_total_deliveries(self.h)

    def total_ticks(self):
        return self.l.orc_total_ticks(self.h)

    def published_by_code(self, code):
        return self.l.orc_published_by_code(self.h, code)


def gobus_bench(n_subs, n_events, mailbox_cap=1000, n_threads=1):
    chk = C.c_uint64()
    return lib().gobus_bench(n_subs, n_events, mailbox_cap, n_threads, C.byref(chk))


def gobus_bench_steps(n_subs, events_per_step, steps, warmup, mailbox_cap=1000, n_threads=1):
    """(deliveries/s, seconds) over `steps` timed steps; threads and channels are created once."""
    sec = C.c_double()
    v = lib().gobus_bench_steps(n_subs, events_per_step, steps, warmup, mailbox_cap, n_threads, C.byref(sec))
    return v, sec.value


def gobus_bench_steps2(n_subs, events_per_step, steps, warmup, mailbox_cap=1000, n_threads=1, send_only=False):
    """(deliveries/s over the timed steps, per-step seconds of the timed steps) — pinned threads, NUMA-local mailboxes."""
    sec = np.zeros(steps + warmup, dtype=np.float64)
    v = lib().gobus_bench_steps2(n_subs, events_per_step, steps, warmup, mailbox_cap, n_threads, int(send_only), sec.ctypes.data)
    return v, sec[warmup:].copy()
