"""Worker of tests/test_gpu_multi.py, launched as `python -m torch.distributed.run --nproc-per-node G tests/multi_worker.py ...`:
one process per GPU, each owning a contiguous shard of the subscribers (containerpilot_b200.sharding.ShardedBus).  Writes
every subscriber's (count, digest) and the mailbox windows of sampled subscribers to <out>/rank<r>.npz."""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from multi_trace import make_case  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--mode", default="stream", choices=["stream", "trace"])
    ap.add_argument("--subs", type=int, default=4096)
    ap.add_argument("--batches", type=int, default=48)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--lookahead", type=int, default=2)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from containerpilot_b200 import _native as nat
    from containerpilot_b200.sharding import ShardedBus

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    case = make_case(args.subs, args.batches, args.batch)
    sb = ShardedBus(args.subs, dist=dist, rank=rank, world=world, device=local, ring_cap=1024, batch_cap=args.batch,
                    timers_per_sub=1, digest=True, stream_slots=8)
    try:
        sb.subscribe_many(case["masks"][sb.first:sb.first + sb.count])
        sb.timer_add_many(case["period"], source_id0=case["timer_src0"])
        B, recs, wms = args.batch, case["records"], case["watermarks"]
        nb = args.batches
        if args.mode == "stream":
            assert sb.stream_ok, "stream handshake failed"
            put = 0
            for j in range(nb):
                while put < nb and put <= j + args.lookahead:
                    # (waits, bounded, for consumers that are merely behind: their launches are queued on busy GPUs)
                    nat.check(sb.put(recs[put * B:(put + 1) * B], int(wms[put]), raw=True), "cpbus_stream_put"); put += 1
                nat.check(sb.fanout(B, int(wms[j])), "cpbus_stream_fanout")
        else:
            ptr = sb.attach_trace(nb * B * 32)
            assert sb.trace_ok, "peer mapping of the trace failed"
            if rank == 0:
                class _Raw:
                    __cuda_array_interface__ = {"shape": (nb * B, 32), "typestr": "|u1", "data": (ptr, False), "version": 2}
                view = torch.as_tensor(_Raw(), device=torch.device("cuda", local))
                view.copy_(torch.from_numpy(recs.view(np.uint8).reshape(-1, 32).copy()))
                torch.cuda.synchronize()
            sb.barrier()
            for j in range(nb):
                nxt = (j + 2) * B * 32 if j + 2 < nb else None
                nat.check(sb.fanout_trace(j * B * 32, B, int(wms[j]), nxt, B), "fanout_trace")
        sb.bus.sync()
        assert sb.bus.stream_status(sb._st) == nat.OK
        dg = sb.digests()
        windows = {}
        for s in case["sampled"]:
            if sb.first <= s < sb.first + sb.count:
                windows[f"w{s}"] = sb.bus.peek_window(int(s))
        st = sb.bus.stats()
        fold = sb.digest_fold_all()
        np.savez(os.path.join(args.out, f"rank{rank}.npz"), first=sb.first, count=dg["count"], digest=dg["digest"],
                 deliveries=st["deliveries"], ticks=st["ticks"], fold=np.array(fold, dtype=np.uint64),
                 ingest=np.array(sb.ingest), **windows)
        sb.barrier()
    finally:
        sb.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
