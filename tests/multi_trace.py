"""The seeded multi-GPU test case shared by tests/multi_worker.py (GPU ranks) and tests/test_gpu_multi.py (oracle side):
Zipf masks, one periodic timer per subscriber, complete records with a sprinkling of unicast sends to global ids."""
import numpy as np

import trace as tr

EVENT_DTYPE = np.dtype([("seq", "<u8"), ("ts_ns", "<u8"), ("code", "<u4"), ("source_id", "<u4"), ("target", "<u4"), ("flags", "<u4")])


def make_case(n_subs: int, n_batches: int, B: int, seed: int = 0xC0DEB204):
    rng = np.random.default_rng(seed)
    n = n_batches * B
    dt = 10_000
    masks = tr.zipf_masks(n_subs, 1.0, seed)
    masks[::7] = 0x1FFFF                                       # some take everything (the reference's behaviour)
    rec = np.zeros(n, dtype=EVENT_DTYPE)
    rec["seq"] = np.arange(n)
    rec["ts_ns"] = (np.arange(n) + 1) * dt
    rec["code"] = tr.zipf_codes(n, 1.0, seed + 1)
    rec["source_id"] = rng.integers(0, 4096, n)
    rec["target"] = 0xFFFFFFFF
    uni = rng.random(n) < 0.01                                  # `job.Rx <- ev` style direct sends (jobs/jobs.go:262)
    rec["target"][uni] = rng.integers(0, n_subs, int(uni.sum()))
    rec["flags"][uni] = 2
    watermarks = rec["ts_ns"][B - 1::B].copy()
    sampled = np.unique(np.concatenate([[0, n_subs - 1, n_subs // 2, n_subs // 2 - 1], rng.integers(0, n_subs, 12)]))
    return {"masks": masks, "records": rec, "watermarks": watermarks, "period": 37 * dt + 3, "timer_src0": 1_000_000,
            "sampled": sampled, "dt": dt}
