#!/usr/bin/env python
"""bench.py — events/sec through EventBus.Publish on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: `batch` published events
fanned out to every subscriber mailbox of every GPU's shard (one fan-out kernel
launch per GPU).  Headline unit: deliveries/s = 32-byte records
landed in mailboxes per second, whole job; publishes/s is reported beside it.

  python bench.py [--gpus N --steps K --warmup W] [--workload config2|config3|config5]
  python bench.py --impl reference ...      # the reference's CPU path (restated Go bus) on host cores

One JSON line on stdout (rank 0).  Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "config2": dict(subs=65_536, events=10_000_000, timers=0, zipf=None,
                    desc="1xB200: 65,536 subscribers, 10M-event synthetic trace, 32-byte records, all-ones masks"),
    # configs[2]: 1,048,576 subscribers, 1 kHz timer per subscriber (a stated prefix of the 100M-event trace)
    "config3": dict(subs=1_048_576, events=100_000_000, timers=1, zipf=None,
                    desc="1xB200: 1,048,576 subscribers, 100M events (prefix timed), Timer ticks interleaved at 1 kHz"),
    # configs[4]: Zipf-skewed masks over 16 codes
    "config5": dict(subs=1_048_576, events=10_000_000, timers=0, zipf=1.0,
                    desc="filter sweep: 1,048,576 subscribers, 16 event codes, Zipf(s=1.0) masks and codes"),
}
DT_NS = 10_000            # virtual time per publish: 1e5 publishes per virtual second
TICK_NS = 1_000_000       # 1 kHz


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="batches to time (0 = the workload's whole trace, capped)")
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="cpbus", choices=["cpbus", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=512, help="events per step (<= ring/2); 512 measured best: config 2 +1.5 %, config 3 +8 % over 256")
    ap.add_argument("--ring", type=int, default=1024)
    ap.add_argument("--subs", type=int, default=0, help="override subscribers per GPU")
    ap.add_argument("--store", type=int, default=0, help="0 auto, 1 v4, 2 v8, 3 TMA bulk")
    ap.add_argument("--grid", type=int, default=0)
    ap.add_argument("--no-digest", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--max-steps", type=int, default=40_000)
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (NVML, in-process)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def sample(self):
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("hw_thermal_slowdown", 0x40),
                              ("sw_thermal_slowdown", 0x20), ("hw_power_brake", 0x80), ("sync_boost", 0x10)):
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        if self.nv is None:
            return
        while not self.stop_flag:
            self.sample()
            time.sleep(0.02)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_baseline(n_subs: int, seconds_target: float = 12.0):
    """Restated Go bus (oracle/gobus_baseline.c) on this box's host cores: a reported baseline."""
    import oracle_binding as ob
    cores = os.cpu_count() or 1
    probe = ob.gobus_bench(n_subs, 20, 1000, 1)                       # ~0.1 s probe to size the sample
    n_events = int(max(50, min(20_000, probe * seconds_target / 2 / n_subs)))
    single = ob.gobus_bench(n_subs, n_events, 1000, 1)
    multi = ob.gobus_bench(n_subs, n_events * min(cores, 8), 1000, cores) if cores > 1 else single
    return {"value": multi, "unit": "deliveries/s", "cores": cores, "kind": "port",
            "single_thread_value": single,
            "sample": f"{n_subs} subscribers x {n_events} events single-threaded (GOMAXPROCS(1)-faithful) and "
                      f"x {n_events * min(cores, 8)} events sharded over {cores} threads; mailbox cap 1000, "
                      f"restated Go bus (no Go toolchain in this image), value = the sharded all-cores run"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Go bus cannot be
    built here, so this is the oracle port of its cost model, with all the host threads it can use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_binding as ob
    wl = WORKLOADS[args.workload]
    n_subs = args.subs or wl["subs"]
    cores = os.cpu_count() or 1
    steps = args.steps or 8
    warmup = max(0, args.warmup)
    probe = ob.gobus_bench(n_subs, 4, 1000, cores)                    # deliveries/s estimate, to size the per-step sample
    budget_s = 45.0                                                   # the whole --steps K --warmup W run ends within about a minute
    per_step_events = int(max(1, min(50_000, probe * budget_s / n_subs / (steps + warmup))))
    value, dt = ob.gobus_bench_steps(n_subs, per_step_events, steps, warmup, 1000, cores)
    line = {"impl": "reference", "metric": "events/sec through Bus.Publish (deliveries/s)", "value": value,
            "unit": "deliveries/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": wl["desc"], "subscribers": n_subs, "events_per_step": per_step_events,
                       "mailbox_cap": 1000},
            "cpu_baseline": {"value": value, "unit": "deliveries/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} steps x {per_step_events} events x {n_subs} subscribers, sharded over {cores} threads"},
            "e2e": {"value": value, "unit": "deliveries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from containerpilot_b200 import _native as nat
    from containerpilot_b200.bus import Bus, EVENT_DTYPE
    import trace as tr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; cpbus has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    wl = WORKLOADS[args.workload]
    n_subs = args.subs or wl["subs"]                  # per GPU (weak scaling: contiguous shards)
    B, R = args.batch, args.ring
    n_batches_trace = wl["events"] // B
    steps = args.steps or min(n_batches_trace, args.max_steps)
    warmup = max(args.warmup, 3)
    K_timers = 1 if wl["timers"] else 0

    # ---- synthetic trace, resident in HBM (rank 0 is the publisher's GPU) ----
    # trace length: long enough that the timed region of a short run never crosses a re-stamp boundary, and the default
    # full run crosses at most one (32768 batches x 16 KiB = 512 MiB of HBM)
    n_trace_batches = min(max(steps + warmup, 2048), 32768)
    n_ev = n_trace_batches * B
    g = torch.Generator(device="cpu"); g.manual_seed(0xC0DEB200 + 2)
    if wl["zipf"]:
        codes = torch.from_numpy(tr.zipf_codes(n_ev, wl["zipf"], 0xC0DEB205).astype(np.int64))
    else:
        codes = torch.randint(1, 17, (n_ev,), generator=g)
    srcs = torch.randint(0, 4096, (n_ev,), generator=g)

    def make_records(first_seq: int) -> np.ndarray:
        ev = np.zeros(n_ev, dtype=EVENT_DTYPE)
        ev["seq"] = first_seq + np.arange(n_ev, dtype=np.uint64)
        ev["ts_ns"] = (first_seq + 1 + np.arange(n_ev, dtype=np.uint64)) * DT_NS
        ev["code"], ev["source_id"], ev["target"] = codes.numpy(), srcs.numpy(), nat.TARGET_ALL
        return ev

    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    bus = Bus(n_subs, ring_cap=R, batch_cap=B, timers_per_sub=K_timers, digest=not args.no_digest, device=local,
              sub_id_base=rank * n_subs, store_path=args.store, stream=stream.cuda_stream, grid_ctas=args.grid)
    if wl["zipf"]:
        masks = tr.zipf_masks(n_subs, wl["zipf"], 0xC0DEB205 + rank)
    else:
        masks = np.full(n_subs, nat.MASK_ALL, dtype=np.uint32)
    bus.subscribe_many(masks)
    if K_timers:
        bus.timer_add_many(rank * n_subs, n_subs, TICK_NS, source_id0=1_000_000 + rank * n_subs)

    # The device trace is re-stamped per cycle on the GPU so seq/ts stay monotonic over any number of steps.
    base = make_records(0)
    trace_dev = torch.from_numpy(base.view(np.uint8).reshape(n_ev, 32)).to(dev) if rank == 0 else \
        torch.empty((n_ev, 32), dtype=torch.uint8, device=dev)
    trace_q = trace_dev.view(torch.int64).view(n_ev, 4)               # words: seq, ts, code|src, target|flags
    seq0 = torch.arange(n_ev, dtype=torch.int64, device=dev)

    def restamp(cycle: int):
        if rank == 0:
            trace_q[:, 0] = seq0 + cycle * n_ev
            trace_q[:, 1] = (seq0 + 1 + cycle * n_ev) * DT_NS

    CHUNK = 64                                                         # batches per NCCL broadcast (512 KiB), fallback ingest only
    state = {"step": 0}
    chunk_events = []                                                  # CPBUS_BENCH_TRACE=1: per-1000-step device timing

    # Multi-GPU ingest of the HBM-resident stream.  Preferred: the publisher's trace is peer-mapped (CUDA IPC over
    # NVLink) and every shard's fan-out kernel pulls its batch itself (cpbus_publish_device_staged: CTA 0 reads the
    # 8 KiB across the link, stages it locally) — no collective call on the data path.  Fallback: NCCL broadcast.
    ingest_mode, peer_trace = "local", None
    if world > 1:
        ingest_mode = "nccl-broadcast"
        if not os.environ.get("CPBUS_BENCH_NCCL_INGEST"):
            # the publisher's stream lives in a cpbus_shared_alloc buffer; the other ranks map it (CUDA IPC, NVLink).
            # Every rank reaches both collectives below whatever happens locally, so a failure cannot hang the job.
            local_ok, handle = 1, None
            if rank == 0:
                try:
                    shared_ptr, handle = bus.shared_alloc(n_ev * 32)

                    class _Raw:                                        # zero-copy torch view of the shared buffer
                        __cuda_array_interface__ = {"shape": (n_ev, 32), "typestr": "|u1", "data": (shared_ptr, False), "version": 2}
                    shared_view = torch.as_tensor(_Raw(), device=dev)
                    shared_view.copy_(trace_dev)
                    trace_dev = shared_view
                    trace_q = trace_dev.view(torch.int64).view(n_ev, 4)
                    torch.cuda.synchronize()
                except Exception as ex:                                # pragma: no cover - depends on the box
                    print(f"[bench] shared stream buffer unavailable ({ex!r})", file=sys.stderr)
                    local_ok, handle = 0, None
            box = [handle]
            dist.broadcast_object_list(box, src=0)
            if rank != 0:
                if box[0] is None:
                    local_ok = 0
                else:
                    try:
                        peer_trace = bus.shared_open(box[0])
                    except Exception as ex:                            # pragma: no cover
                        print(f"[bench] peer mapping unavailable ({ex!r})", file=sys.stderr)
                        local_ok, peer_trace = 0, None
            ok_t = torch.tensor([local_ok], device=dev)
            dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
            if int(ok_t.item()) == 1:
                ingest_mode = "nvlink-peer-pull (fused into the fan-out kernel)"
            elif rank == 0:
                print("[bench] falling back to NCCL broadcast of the event stream", file=sys.stderr)
        src_ptr = peer_trace if (peer_trace is not None and ingest_mode.startswith("nvlink")) else trace_dev.data_ptr()
    else:
        src_ptr = trace_dev.data_ptr()
    fused = ingest_mode.startswith("nvlink")

    def run_steps(k: int, trace_chunks: bool = False):
        """k fan-out steps from the HBM-resident trace."""
        for j in range(k):
            if trace_chunks and j % 1000 == 0:
                ev_ = torch.cuda.Event(enable_timing=True); ev_.record(stream); chunk_events.append((j, ev_, time.perf_counter()))
            i = state["step"]
            slot, cycle = i % n_trace_batches, i // n_trace_batches
            if slot == 0 and cycle > 0:
                if fused:                                              # nobody may be reading while the publisher re-stamps
                    torch.cuda.synchronize(); dist.barrier()
                restamp(cycle)
                if fused:
                    torch.cuda.synchronize(); dist.barrier()
            if world > 1 and not fused and slot % CHUNK == 0:
                hi = min(slot + CHUNK, n_trace_batches)
                dist.broadcast(trace_dev[slot * B: hi * B], src=0)
            wm = (i + 1) * B * DT_NS
            if fused:
                nslot = slot + 2                                       # the batch after next: pulled by THIS launch, hidden under its stores
                nxt_ptr = src_ptr + nslot * B * 32 if nslot < n_trace_batches else 0   # (not across a re-stamp boundary)
                nat.check(bus.publish_device_staged(src_ptr + slot * B * 32, B, wm, nxt_ptr, B if nxt_ptr else 0), "cpbus_publish_device_staged")
            else:
                nat.check(bus.publish_device(src_ptr + slot * B * 32, B, wm), "cpbus_publish_device")
            state["step"] = i + 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: `value` ----
    sampler = ClockSampler(local)
    sampler.start()                                                    # started early so NVML is warm before the timed region
    run_steps(warmup)
    # settle: a fresh box pages in driver/library code lazily; keep warming (untimed) for ~0.3 s of wall clock
    settle, t_settle = 0, time.perf_counter()
    while time.perf_counter() - t_settle < 0.3 and settle < 20_000:
        run_steps(100); torch.cuda.synchronize(); settle += 100
    # if the timed region would straddle the end of the trace, start it at the next cycle instead (re-stamp outside the timing)
    pos = state["step"] % n_trace_batches
    if steps <= n_trace_batches and pos + steps > n_trace_batches:
        cycle = state["step"] // n_trace_batches + 1
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        restamp(cycle)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        state["step"] = cycle * n_trace_batches
        nat.check(bus.advance(state["step"] * B * DT_NS), "cpbus_advance")   # armed timers catch up in bounded windows
    barrier()
    st0 = bus.stats()
    sampler.samples.clear(); sampler.reasons.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    run_steps(steps, trace_chunks=bool(os.environ.get("CPBUS_BENCH_TRACE")))
    e1.record(stream)
    sampler.sample()                                                   # launches are asynchronous: the GPU is still inside the timed region here
    barrier()
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    if chunk_events and rank == 0:
        print("per-chunk device us/step:", [round(a[1].elapsed_time(b[1]) / (b[0] - a[0]) * 1e3, 1) for a, b in zip(chunk_events, chunk_events[1:])], file=sys.stderr)
        print("per-chunk host us/step:", [round((b[2] - a[2]) / (b[0] - a[0]) * 1e6, 1) for a, b in zip(chunk_events, chunk_events[1:])], file=sys.stderr)
    st1 = bus.stats()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    deliv = torch.tensor([st1["deliveries"] - st0["deliveries"], st1["ticks"] - st0["ticks"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(deliv, op=dist.ReduceOp.SUM)
    deliveries, ticks = float(deliv[0].item()), float(deliv[1].item())
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    value = deliveries / (ms * 1e-3)
    publishes_per_s = steps * B / (ms * 1e-3)

    # ---- roofline for the dominant kernel (fan-out): algorithmic bytes per launch / avg launch duration ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    d_local = (st1["deliveries"] - st0["deliveries"]) / steps        # records per launch on this GPU
    # algorithmic bytes per launch: every delivered record is one 32-byte sector; every mailbox's 32-byte
    # control block is read once and written once; an armed timer slot costs its 16-byte hot half read plus, when it fires,
    # the cold half read and both halves written (~64 B); the batch itself is read once from HBM
    per_sub_state = 64 + (64 if K_timers else 0)
    alg_bytes = 32.0 * d_local + n_subs * per_sub_state + B * 32
    kernel_ms = float(e0.elapsed_time(e1)) / steps                    # this rank's launches are back to back on the stream
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = (json.load(open(tp)).get(args.workload) or {}).get(str(B))
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "cpbus_dev::fanout_kernel", "alg_bytes_per_launch": alg_bytes,
                "record_bytes_per_launch": 32.0 * d_local, "state_bytes_per_launch": float(n_subs * per_sub_state + B * 32),
                "kernel_ms": kernel_ms, "peak_source": peak_src}

    # ---- end to end through the public C-ABI with HOST buffers: `e2e` ----
    e2e = None
    if not args.no_e2e:
        k2 = min(steps, 4000)
        host = base.copy()                                            # events as a caller holds them (host memory)
        fold = None
        nxt = state["step"]
        tickets = []                                                  # result reads are pipelined two steps deep
        primed = [False]
        if world > 1:
            comm = torch.cuda.Stream(device=dev)
            ingest = torch.empty((4, B, 32), dtype=torch.uint8, device=dev)
            ready = [torch.cuda.Event() for _ in range(4)]
            used = [torch.cuda.Event() for _ in range(4)]

            def issue_ingest(i: int):
                kslot, slot = i % 4, i % n_trace_batches
                with torch.cuda.stream(comm):
                    comm.wait_event(used[kslot])
                    if rank == 0:
                        ingest[kslot].copy_(pinned[slot * B: (slot + 1) * B], non_blocking=True)
                    dist.broadcast(ingest[kslot], src=0)
                    ready[kslot].record(comm)

        def e2e_step(j: int):
            nonlocal fold
            i = nxt + j
            if world == 1:
                nat.check(bus.advance((i + 1) * B * DT_NS), "cpbus_advance")
                lo = (i % n_trace_batches) * B
                nat.check(bus.publish_many(host[lo: lo + B]), "cpbus_publish")   # pinned staging + H2D inside
                nat.check(bus.flush(), "cpbus_flush")
            else:
                # ingest of step i+1 (H2D on rank 0 + NCCL broadcast over NVLink) runs one step ahead on a side
                # stream, overlapped with the fan-out of step i; a 4-slot ring of staging buffers, ordered by events
                if not primed[0]:
                    issue_ingest(i); primed[0] = True
                issue_ingest(i + 1)
                kslot = i % 4
                stream.wait_event(ready[kslot])
                nat.check(bus.publish_device(ingest[kslot].data_ptr(), B, (i + 1) * B * DT_NS), "cpbus_publish_device")
                used[kslot].record(stream)
            tickets.append(bus.step_result_begin())                   # 256-byte D2H of the step's result (written by the fan-out kernel)
            if len(tickets) > 2:
                fold = bus.step_result_end(tickets.pop(0))            # ...read two steps later: the GPU never idles

        if world > 1:
            # ts must keep increasing: restamp host copy for the e2e region
            pinned = torch.from_numpy(make_records(0).view(np.uint8).reshape(n_ev, 32)).pin_memory() if rank == 0 else None
            if rank == 0:
                pq = pinned.view(torch.int64).view(n_ev, 4)
                off = (nxt // n_trace_batches + 1) * n_ev
                pq[:, 0] += off
                pq[:, 1] = (pq[:, 0] + 1) * DT_NS
            nxt = (nxt // n_trace_batches + 1) * n_trace_batches
            nat.check(bus.advance(nxt * B * DT_NS), "cpbus_advance")   # the clock jumps to the next trace cycle: armed timers catch up in bounded windows
            k2 = min(k2, n_trace_batches - 8)
        for j in range(3):
            e2e_step(j)
        while tickets:
            bus.step_result_end(tickets.pop(0))
        nxt += 3
        barrier()
        s0 = bus.stats()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        f0.record(stream)
        for j in range(k2):
            e2e_step(j)
        while tickets:
            fold = bus.step_result_end(tickets.pop(0))                # every step's result has reached the host
        f1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - w0) * 1e3
        s1 = bus.stats()
        t2 = torch.tensor([max(f0.elapsed_time(f1), wall_ms)], dtype=torch.float64, device=dev)
        d2 = torch.tensor([s1["deliveries"] - s0["deliveries"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX); dist.all_reduce(d2, op=dist.ReduceOp.SUM)
        e2e = {"value": float(d2.item()) / (float(t2.item()) * 1e-3), "unit": "deliveries/s",
               "h2d_bytes_per_step": B * 32, "d2h_bytes_per_step": 256, "steps": k2,
               "ms_per_step": float(t2.item()) / k2,
               "api": "cpbus_advance+cpbus_publish(host events)+cpbus_flush+cpbus_step_result_begin/_end (deliveries + digest checksum of the step, read 2 steps later)" if world == 1 else
                      "pinned host batch -> H2D on rank 0 -> NCCL broadcast (side stream, one step ahead) -> cpbus_publish_device + cpbus_step_result_begin/_end"}
        launches_e2e = s1["kernel_launches"] - s0["kernel_launches"]
    else:
        launches_e2e = 0

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(min(n_subs, 65_536))

    if world > 1:
        # importers unmap the publisher's stream before the publisher frees it
        torch.cuda.synchronize()
        if peer_trace is not None:
            try:
                bus.shared_close(peer_trace)
            except Exception as ex:                                    # pragma: no cover - teardown only
                print(f"[bench] peer unmap: {ex!r}", file=sys.stderr)
        dist.barrier()
    bus.close()
    if rank == 0:
        line = {
            "metric": "events/sec through Bus.Publish (deliveries/s = 32-byte records landed in subscriber mailboxes)",
            "value": value, "unit": "deliveries/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": wl["desc"], "subscribers_per_gpu": n_subs, "subscribers_total": n_subs * world,
                       "events_per_step": B, "ring_cap": R, "record_bytes": 32, "mode": "overwrite-oldest throughput mode",
                       "digest": not args.no_digest, "timers_per_sub": K_timers, "warmup_settle_steps": settle, "store_path": args.store,
                       "parallelism": f"subscriber shards x{world}" + (f", ingest: {ingest_mode}" if world > 1 else ""),
                       "l2": f"inputs larger than L2: {n_subs * R * 32 / 2**30:.1f} GiB of rings per GPU, "
                             f"{d_local * 32 / 2**20:.0f} MiB written per step vs 126 MB L2",
                       "trace": f"splitmix-seeded uniform codes 1..16, 4096 sources, {n_trace_batches} batches cycled with re-stamped seq/ts"},
            "publishes_per_s": publishes_per_s, "deliveries": deliveries, "ticks": ticks,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "gpu_launches_e2e": int(launches_e2e), "clocks": sampler.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
