"""Multi-GPU parity on hardware (BASELINE config 4 shape): one process per GPU under
torch.distributed.run, real CUDA-IPC / NVLink peer mappings, Zipf masks + timers + unicast.  Every subscriber's
(count, digest) must equal the oracle's and be identical at G = 1, 2, 4, 8; sampled subscribers' mailbox windows are
compared record for record.  G > device_count is skipped (the driver's 1-GPU box runs the G = 1 leg)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_binding as ob
from multi_trace import make_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_SUBS, N_BATCHES, B = 4096, 48, 128
_oracle_cache = {}


def _oracle():
    if "full" not in _oracle_cache:
        case = make_case(N_SUBS, N_BATCHES, B)
        orc = ob.Oracle(N_SUBS, timers_per_sub=1, keep_window=1024)
        for s in range(N_SUBS):
            orc.subscribe(int(case["masks"][s]))
            orc.timer_add(s, case["period"], case["timer_src0"] + s, False)
        for j in range(N_BATCHES):
            assert orc.publish_records(case["records"][j * B:(j + 1) * B], int(case["watermarks"][j])) == 0
        _oracle_cache["full"] = (case, orc)
    return _oracle_cache["full"]


def _run(G, mode, tmp_path, port):
    out = tmp_path / f"{mode}{G}"
    out.mkdir()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={G}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_worker.py"), "--out", str(out), "--mode", mode,
           "--subs", str(N_SUBS), "--batches", str(N_BATCHES), "--batch", str(B)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return [np.load(out / f"rank{k}.npz") for k in range(G)]


@pytest.mark.timeout(900)
@pytest.mark.parametrize("mode", ["stream", "trace"])
@pytest.mark.parametrize("G", [1, 2, 4, 8])
def test_sharded_bus_equals_oracle_at_every_shard_count(G, mode, tmp_path):
    import torch
    if torch.cuda.device_count() < G:
        pytest.skip(f"needs {G} GPUs")
    case, orc = _oracle()
    ranks = _run(G, mode, tmp_path, 29600 + G + (10 if mode == "trace" else 0))
    count = np.concatenate([r["count"] for r in ranks]); digest = np.concatenate([r["digest"] for r in ranks])
    assert len(count) == N_SUBS
    want_c = np.array([orc.count(s) for s in range(N_SUBS)], dtype=np.uint64)
    want_d = np.array([orc.digest(s) for s in range(N_SUBS)], dtype=np.uint64)
    bad = np.nonzero((count != want_c) | (digest != want_d))[0]
    assert len(bad) == 0, f"{len(bad)} subscribers differ from the oracle, first {bad[:8]}"
    assert sum(int(r["deliveries"]) for r in ranks) == orc.total_deliveries()
    assert sum(int(r["ticks"]) for r in ranks) == orc.total_ticks()
    for r in ranks:                                             # sampled mailboxes, record for record
        for key in r.files:
            if key.startswith("w"):
                s = int(key[1:])
                assert r[key].tobytes() == orc.mailbox(s)[-len(r[key]):].tobytes(), f"window of subscriber {s}"
    folds = {tuple(int(x) for x in r["fold"]) for r in ranks}   # every rank reduced the same global fold
    assert len(folds) == 1
    if G > 1:
        assert "nvlink" in str(ranks[1]["ingest"])              # the fused path, not the NCCL fallback
    _oracle_cache.setdefault("folds", {})[(G, mode)] = folds.pop()
    assert len(set(_oracle_cache["folds"].values())) == 1       # ... and the fold does not depend on G or the ingest mode
