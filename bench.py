#!/usr/bin/env python
"""bench.py — filter + GROUP BY throughput of the B200-native path (BASELINE.json metric) and of the CPU arm.

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the CPU arm: the reference-shaped C restatement of the
                                                            Go chain (oracle/oracle_ref.c; the Go reference itself
                                                            cannot be built in this image), all host cores

Workload (config.workload = "config2"): BASELINE.json configs[1] - 10 M flat documents
{"id","n" int64 U[0,1e6),"f" float64 U[0,1),"type"} per GPU and step,
  SELECT COUNT(*),COUNT(n),SUM(n),AVG(n),MIN(n),MAX(n),SUM(f) FROM d WHERE n BETWEEN 250000 AND 749999   (50 % selectivity)
A step = one pass of Filter + InitialGroup/IntermediateGroup/FinalGroup over one 10 M-row batch.
  value : rows/s with the shredded columns already resident in HBM (4 table copies are rotated so that no step
          finds its 160 MB of input in the 126 MB L2), timed with CUDA events on the launching stream, max over ranks.
  e2e   : rows/s through the public API from HOST buffers: JSON documents -> shredder (host threads) -> H2D ->
          scan -> result on the host, every step.
  roofline: the scan kernel nq_scan (+ its 1-block partial reduction) - column bytes it must read / its mean
          CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline: oracle_ref.c (kind "port") on a bounded sample of the same documents, all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS = 10_000_000
ALIAS = "d"
WHERE = "((`d`.`n`) between 250000 and 749999)"
AGGS = ["count(*)", "count((`d`.`n`))", "sum((`d`.`n`))", "avg((`d`.`n`))", "min((`d`.`n`))", "max((`d`.`n`))", "sum((`d`.`f`))"]
KEYS = []
METRIC = "filter+GROUP BY rows/sec (columns resident in HBM)"
UNIT = "rows/s"
CONFIG = {"workload": "config2", "rows_per_gpu_per_step": ROWS, "query": "SELECT COUNT(*),COUNT(n),SUM(n),AVG(n),MIN(n),MAX(n),SUM(f) "
          "FROM d WHERE n BETWEEN 250000 AND 749999", "selectivity": 0.5, "partitioning": "row ranges, one per GPU", "pipelining": "independent steps overlap on 4 CUDA streams",
          "l2": "4 rotating table copies per GPU (640 MB) > 126 MB L2"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [l.strip().split(", ") for l in open(self.path) if l.strip()]
            sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
                out["sm_max_mhz"] = float(rows[0][1])
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            seen = set()
            for r in rows:
                for i, nm in enumerate(names):
                    if len(r) > 3 + i and r[3 + i].strip().lower() == "active":
                        seen.add(nm)
            out["reasons"] = sorted(seen)
            out["samples"] = len(rows)
            os.unlink(self.path)
        except Exception:
            pass
        return out


def synth_columns(seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    n = rng.integers(0, 1_000_000, ROWS, dtype=np.int64)
    n[0], n[1] = 0, 999_999  # identical column statistics for every copy -> one compiled kernel serves them all
    f = rng.integers(0, 1_000_000, ROWS).astype(np.float64) / 1e6
    f[f == 0.0] = 0.5        # keep the column purely float64 (an integral value would be an int: value.NewValue)
    return n, f


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import query_b200 as q
    from query_b200 import dist as qd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout (one JSON line only)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    q.init(local)
    K, W = args.steps, max(args.warmup, 3)

    # ---- resident columns: NT rotating copies ------------------------------------------------------------------
    NT = 4
    tables = []
    for c in range(NT):
        n, f = synth_columns(1000 * rank + c + 1)
        t = q.Table(["n", "f"])
        t.set_column("n", n)
        t.set_column("f", f, tags=np.full(ROWS, 5, dtype=np.uint8))
        t.seal()
        tables.append(t)
    NQ = 8
    NS = args.streams  # scans of independent batches overlap on NS CUDA streams (tails hide behind the next scan)
    side = [torch.cuda.Stream() for _ in range(NS)]
    queries = []
    for i in range(NQ):
        qq = q.Query(tables[i % NT], ALIAS, WHERE, KEYS, AGGS)
        qq.set_stream(side[i % NS].cuda_stream)
        qq.set_timing(False)  # the timed region is bracketed by our own events; per-launch events only cost front-end time
        queries.append(qq)
    # N > 1: the Intermediate->Final merge is fused into the scan kernel (peer stores over NVLink into every rank's
    # mailbox + a 1-block fold); --merge nccl switches to the NCCL all_gather of the accumulator words instead
    mailbox = qd.make_mailbox(max_words=1024) if (world > 1 and args.merge == "fused") else None
    dqs = [qd.DistributedQuery(qq, stream=side[i % NS], mailbox=mailbox) for i, qq in enumerate(queries)]
    info = queries[0].info
    bytes_per_row = info["scan_bytes_per_row"]

    def step_sync(i):
        """one step, blocking (used for warm-up and for N > 1 where every step ends in the merge exchange)"""
        return dqs[i % NQ].execute()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(W):
        r = step_sync(i)
    # soak: keep scanning for ~1 s (untimed warm-up) so that clocks are ramped and nvidia-smi has samples under load
    t_soak = time.perf_counter()
    soak_steps = 0
    while args.soak > 0:
        for _ in range(50):
            r = step_sync(soak_steps)
            soak_steps += 1
        done = time.perf_counter() - t_soak >= args.soak
        if world > 1:  # every rank must run the same number of (collective) steps: rank 0 decides
            flag = torch.tensor([1 if done else 0], device="cuda")
            dist.broadcast(flag, 0)
            done = bool(flag.item())
        if done:
            break
    W += soak_steps
    check_rows = r.rows() if rank == 0 else None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = q.launch_count()
    scan_ns = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    main = torch.cuda.current_stream()
    ev0.record(main)
    for s_ in side:
        s_.wait_event(ev0)   # device-side bracket: no scan starts before ev0 ...
    # pipelined: up to NQ steps in flight over the streams; results are collected in order.  For N > 1 a step is
    # scan -> all_gather of the accumulator words (NCCL, stream-ordered) -> merge kernel: no host round trip.
    inflight = []
    for s in range(K):
        h = s % NQ
        if len(inflight) == NQ:
            j = inflight.pop(0)
            dqs[j].collect()
            scan_ns.append(queries[j].last_scan_ns)
        dqs[h].launch()
        inflight.append(h)
    for j in inflight:
        dqs[j].collect()
        scan_ns.append(queries[j].last_scan_ns)
    for s_ in side:
        main.wait_stream(s_)  # ... and ev1 is recorded after every stream drained
    ev1.record(main)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    launches = q.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tms = torch.tensor([ms], device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    value = world * ROWS * K / (ms / 1e3)

    # the dominant kernel alone: serialised steps on one stream, CUDA events around the nq_scan launch itself
    # (launches are queued NQ deep on ONE stream so that the events bracket the kernel, not the launch latency
    # of an idle queue; the kernels themselves run strictly one after another)
    # NK independent query handles are launched back to back on ONE stream with no per-launch events (an event
    # record costs the GPU front-end several microseconds - a tiny kernel between two events measures 8-10 us on
    # this system); two events bracket the whole run, the kernels execute strictly one after another, and the
    # average launch duration is elapsed / launches.  Tables rotate, so no launch finds its input in L2.
    # Measured twice: as shipped (ungrouped scans are launched with programmatic stream serialization, so the
    # ramp-up of launch i+1 fills the SMs that launch i's tail has left idle), and with N1GPU_NO_PDL=1 (every launch
    # waits for the full completion of the one before: the fixed launch/ramp/tail cost shows up in every launch).
    NK = 32
    ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def kernel_alone():
        kq = []
        for i in range(NK):
            qq = q.Query(tables[i % NT], ALIAS, WHERE, KEYS, AGGS)
            qq.set_stream(side[0].cuda_stream)
            qq.set_timing(False)
            kq.append(qq)
        for qq in kq:
            qq.execute()
        out = []
        for rep in range(5):
            torch.cuda.synchronize()
            ka.record(side[0])
            for qq in kq:
                qq.launch()
            kb.record(side[0])
            for qq in kq:
                qq.collect()
            torch.cuda.synchronize()
            out.append(ka.elapsed_time(kb) * 1e6 / NK)
        return out

    scan_ns = kernel_alone()
    os.environ["N1GPU_NO_PDL"] = "1"
    try:
        serial_ns = kernel_alone()
    finally:
        del os.environ["N1GPU_NO_PDL"]

    # ---- end to end from host JSON ---------------------------------------------------------------------------------
    from oracle import cref  # document generator + CPU baseline only (never the measured path of this arm)
    e2e_rows = args.e2e_rows
    buf, offs = cref.gen_docs(2, 42 + rank, rank * e2e_rows, e2e_rows)
    pinned = torch.from_numpy(buf).pin_memory()   # the step's inputs live in pinned host memory
    hbuf = pinned.numpy()
    pinned_offs = torch.from_numpy(offs).pin_memory()
    hoffs = pinned_offs.numpy()
    e2e_steps = max(1, min(args.e2e_steps, K))

    trace = bool(os.environ.get("N1GPU_TRACE"))

    def e2e_step():
        ts = [time.perf_counter()]
        t = q.Table(["n", "f"])
        t.append_json((hbuf, hoffs), threads=args.shred_threads)
        ts.append(time.perf_counter())
        t.seal()
        qq = q.Query(t, ALIAS, WHERE, KEYS, AGGS)
        ts.append(time.perf_counter())
        res = qd.DistributedQuery(qq, mailbox=mailbox).execute()
        rows = res.rows()
        ts.append(time.perf_counter())
        if trace:
            sys.stderr.write("[bench e2e] shred %.2f ms, seal+compile %.2f ms, scan+result %.2f ms\n" % tuple(
                (b - a) * 1e3 for a, b in zip(ts, ts[1:])))
        if args.shred_threads < 0:
            h2d = int(hoffs[-1]) + 8 * (e2e_rows + 1)          # raw JSON + document offsets
        else:
            h2d = sum(t.scan_bytes(c) for c in ("n", "f")) * e2e_rows  # shredded columns
        d2h = info["words"] * 8
        return rows, h2d, d2h

    rows_e2e, h2d, d2h = e2e_step()  # warm-up (JIT cache, allocator)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rows_e2e, h2d, d2h = e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        ts = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        e2e_s = float(ts.item())
    e2e_value = world * e2e_rows * e2e_steps / e2e_s

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------------------------------
    peak, peak_src = peaks()
    mean_ns = sum(scan_ns) / max(1, len(scan_ns))
    achieved = bytes_per_row * ROWS / mean_ns  # bytes/ns == GB/s
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("nq_scan_config2_dram_bytes_per_launch")
    except Exception:
        pass

    # ---- CPU baseline on a bounded sample, and a parity check of the e2e result against it ---------------------------
    cores = os.cpu_count() or 1
    sample = min(args.cpu_sample, e2e_rows)
    sbuf, soffs = cref.gen_docs(2, 42, 0, sample)
    groups, cpu_s, _passed = cref.run(sbuf, soffs, ALIAS, WHERE, KEYS, AGGS, threads=cores)
    parity = "skipped"
    if world == 1:
        full, _s, _p = cref.run(buf, offs, ALIAS, WHERE, KEYS, AGGS, threads=cores)
        exp, got = full[0][1], rows_e2e[0][1]
        ok = all((abs(a - b) <= 1e-12 * max(abs(a), abs(b))) if isinstance(a, float) or isinstance(b, float) else a == b
                 for a, b in zip(exp, got))
        parity = "ok" if ok else "MISMATCH %r vs %r" % (exp, got)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64+f64", "data": "synthetic", "config": dict(CONFIG, kernel_mode=info["mode"], registers=info["registers"],
                                                                   grid=info["grid"], scan_bytes_per_row=bytes_per_row,
                                                                   survey_bytes_per_row=18, merge=("none" if world == 1 else "fused into nq_scan: peer stores over NVLink into every rank's mailbox + 1-block fold"
                                                                          if args.merge == "fused" else "NCCL all_gather of the accumulator words + merge kernel, stream-ordered")),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "rows_per_step": e2e_rows, "includes": ("H2D of the raw JSON + device shredder (shred.cu) + scan + result on the host" if args.shred_threads < 0
                             else "JSON shredding on host threads + column H2D + scan + result on the host"),
                "json_bytes_per_step": int(offs[-1])},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "nq_scan (filter + aggregation + fused final reduction), timed alone", "peak_source": peak_src, "mean_kernel_us": mean_ns / 1e3,
                     "timing": "32 launches back to back on one stream between two CUDA events, elapsed / 32, 5 repetitions; launches use programmatic stream serialization (PDL)",
                     "serialized_kernel_us": sum(serial_ns) / len(serial_ns) / 1e3,
                     "frac_serialized": bytes_per_row * ROWS / (sum(serial_ns) / len(serial_ns)) / peak,
                     "algorithmic_bytes_per_launch": bytes_per_row * ROWS},
        "cpu_baseline": {"value": sample / cpu_s, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d of the same config2 documents, oracle/oracle_ref.c (reference-shaped C restatement), %d threads" % (sample, cores)},
        "parity_vs_cpu_baseline": parity,
        "result_check": check_rows[0][1] if check_rows else None,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """The CPU arm: the reference's own algorithm for the path (document-at-a-time interpreter over raw JSON, string
    group keys, 3-phase merge) restated in C because Go is not in this image; all host threads; bounded sample/step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cref
    cores = os.cpu_count() or 1
    sample = args.cpu_sample
    buf, offs = cref.gen_docs(2, 42, 0, sample)
    K, W = args.steps, max(1, min(args.warmup, 2))
    t0 = time.perf_counter()
    for _ in range(W):
        cref.run(buf, offs, ALIAS, WHERE, KEYS, AGGS, threads=cores)
    rate = sample * W / max(time.perf_counter() - t0, 1e-9)
    # keep the whole run bounded (~90 s) whatever K the caller asks for: a step is a sample of the same workload
    fit = int(rate * 90.0 / max(K, 1))
    if fit < sample:
        sample = max(50_000, fit)
        buf, offs = cref.gen_docs(2, 42, 0, sample)
        cref.run(buf, offs, ALIAS, WHERE, KEYS, AGGS, threads=cores)
    t0 = time.perf_counter()
    for _ in range(K):
        groups, _s, _p = cref.run(buf, offs, ALIAS, WHERE, KEYS, AGGS, threads=cores)
    el = time.perf_counter() - t0
    value = sample * K / el
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": el / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64+f64",
        "data": "synthetic", "config": dict(CONFIG, rows_per_step=sample),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d config2 documents per step, oracle/oracle_ref.c with %d threads (Go reference not buildable here)" % (sample, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "result_check": groups[0][1],
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-rows", type=int, default=ROWS)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=2_000_000)
    ap.add_argument("--shred-threads", type=int, default=-1, help="-1: device shredder (shred.cu); >= 0: host threads (0 = all cores)")
    ap.add_argument("--merge", default="fused", choices=["fused", "nccl"], help="N > 1: how the per-step partial states are merged")
    ap.add_argument("--streams", type=int, default=4, help="CUDA streams the resident-column steps are pipelined over")
    ap.add_argument("--soak", type=float, default=1.0, help="seconds of untimed scanning before the timed region")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: libraries that print banners to fd 1 (NCCL's version line) are sent to
    # stderr for the duration of the run, and the JSON line goes to the real stdout at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
