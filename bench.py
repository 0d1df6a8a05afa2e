#!/usr/bin/env python
"""bench.py — PairHMM GCUPS on N B200s (BASELINE.json metric) with parity, roofline and CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # CPU PairHMM on the box's host cores

One "step" = one pass of the PairHMM forward path over one synthetic batch of the workload
(default: BASELINE config 2 — 100 000 pairs, 150 bp reads x 300 bp haplotypes, uniform quals).
  value      GCUPS, kernels only, inputs already resident in HBM (CUDA events on the library's launching stream,
             steps rotate over resident batches totalling more than L2 (or --l2 flush), max over ranks); the timed
             steps follow an untimed pre-heat of the same kernels (>= 2 s) with NVML sampled through both
  e2e        same metric through the reference-facing call fcs_pairhmm_compute() with HOST buffers:
             host packing + H2D + kernels + D2H + scatter inside the timed region
  parity     the measured batch scored by the CPU oracle afterwards: fallback decisions, raw FP32 sums (bits),
             |dlog10 L| against the double-precision oracle, per-read best haplotype.  Any mismatch fails the run.
  roofline   FP32 FMA pipe: peak GCUPS = n_SM * 128 * f / 8, at max and at the sustained clock
  configs    (N = 1) the other BASELINE configs, each: kernels-only, e2e, roofline fraction, FP64 share, parity
  e2e_dispatcher  ONE handle over all N devices (the in-process multi-GPU dispatcher) fed a fixed config-3 stream and
             config-4 regions through fcs_pairhmm_compute(): strong scaling, host buffers, wall clock
  cpu_baseline  the oracle's AVX/OpenMP port timed on this box's host cores (rank 0, N = 1)

Under torchrun (N > 1) every rank owns one GPU and scores its own batch of the same shape (regions are independent:
no collective on the data path; torch.distributed is used only for barriers and the max/sum reductions of the
timings) -> "scaling": "weak".  The dispatcher leg then runs on rank 0 alone while the other ranks wait.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pairhmm_gcups"
UNIT = "GCUPS"
TOL = 1e-4  # north_star: |dlog10 L| per pair against the double-precision oracle

WORKLOADS = {
    "c2": "config2: 100k pairs, 150bp reads x 300bp haplotypes, uniform quals (q30/i45/d45/c10)",
    "c2b": "config2b: 100k pairs 150x300, base quals U[6,41]",
    "c1": "config1 stand-in: 400 simulated active regions",
    "c3": "config3 chunk: 2000 WGS-shaped regions, reads 100-250 x haps 100-600",
    "c4": "config4 sample: 20 Mutect2-shaped regions",
    "c5": "config5: underflow stress 250bp x 1kb, 20k pairs",
}


def make_workload(name: str, rank: int):
    import _pkg

    _pkg.load()
    from falcon_genome_b200 import synth

    if name == "c2":
        b = synth.config2_uniform(seed=2002 + rank)
    elif name == "c2b":
        b = synth.config2_uniform(seed=2002 + rank, random_quals=True)
    elif name == "c1":
        b = synth.config1_golden(seed=1001 + rank)
    elif name == "c3":
        b = synth.config3_wgs(n_regions=2000, seed=3003, chunk=rank)
    elif name == "c4":
        b = synth.config4_mutect2(n_regions=20, seed=4004 + rank)
    elif name == "c5":
        b = synth.config5_underflow(seed=5005 + rank)
    else:
        raise SystemExit(f"unknown workload {name}")
    return b, WORKLOADS[name]


def _gen_job(job):
    """(worker process) one synthetic batch as plain arrays."""
    name, rank = job
    b, _ = make_workload(name, rank)
    return b


def make_workloads_parallel(jobs, procs):
    """Generate several synthetic batches in worker processes (the numpy generators are single-threaded Python)."""
    if procs <= 1 or len(jobs) <= 1:
        return [_gen_job(j) for j in jobs]
    import multiprocessing as mp

    with mp.get_context("spawn").Pool(min(procs, len(jobs))) as pool:
        return pool.map(_gen_job, jobs)


def shared_config(args, desc, batch):
    """The `config` object — identical in both arms (our arm and --impl reference) for the same command line."""
    n = max(1, args.gpus)
    return {"workload": desc, "pairs_per_step_per_gpu": int(batch.n_pairs), "cells_per_step_per_gpu": int(batch.cells),
            "n_gpus": n, "parallelism": f"{n} x independent region shards, no collective"}


class ClockSampler(threading.Thread):
    """Samples SM clock, power and throttle reasons through NVML while the device works."""

    def __init__(self, index: int, period_s: float = 0.005):
        super().__init__(daemon=True)
        self.index = index
        self.period = period_s
        self.stop_flag = threading.Event()
        self.t = []
        self.sm = []
        self.power = []
        self.reasons = set()
        self.max_mhz = None
        self.ok = False

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            }
            self.ok = True
            while not self.stop_flag.is_set():
                self.t.append(time.perf_counter())
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    self.power.append(float("nan"))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    try:
                        r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    except Exception:
                        r = 0
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(self.period)
        except Exception:
            self.ok = False

    def window(self, t0, t1):
        t = np.asarray(self.t)
        sel = (t >= t0) & (t <= t1)
        sm = np.asarray(self.sm, dtype=np.float64)[sel]
        pw = np.asarray(self.power, dtype=np.float64)[sel]
        if not self.ok or sm.size == 0:
            return {"sm_mhz": None, "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_mhz_min": float(sm.min()), "samples": int(sm.size),
                "power_w_median": float(np.nanmedian(pw)) if np.isfinite(pw).any() else None,
                "power_w_max": float(np.nanmax(pw)) if np.isfinite(pw).any() else None}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def ncu_traffic(workload):
    """dram__bytes_read + dram__bytes_write of ONE launch of the dominant kernel, from the tracked ncu capture
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py --json from the .ncu-rep of the same build)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            e = json.load(f).get(workload)
        if e:
            return int(e["dram_bytes_read"]) + int(e["dram_bytes_write"]), e
    except Exception:
        pass
    return None, None


def host_threads() -> int:
    """All host cores this process may run on.  Not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
    to its ranks, which would make the CPU reference arm single-threaded at N > 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_isa():
    import ctypes

    from oracle import oracle as O

    lib = O.load()
    lib.phmm_cpu_isa.restype = ctypes.c_char_p
    return lib.phmm_cpu_isa().decode()


CPU_LABEL = ("AVX/OpenMP C port of the oracle ({isa}); vectorised ACROSS reads (16 reads of a region per vector), not GKL's "
             "intra-pair anti-diagonal scheme; float first + double rerun below 1e-28, FTZ on as GKL; not GKL itself (no JVM/GATK in the image)")


def cpu_baseline(batch, reps: int = 3, nthreads: int = 0):
    from oracle import oracle as O

    O.load()
    nthreads = nthreads or host_threads()
    O.batch_simd(batch.select(range(min(4, batch.n_regions))), nthreads, True)  # warm the thread pool / tables
    ts = []
    nd = 0
    for _ in range(reps):
        t = time.perf_counter()
        _, _, _, nd = O.batch_simd(batch, nthreads, True)
        ts.append(time.perf_counter() - t)
    t = float(np.median(ts))
    return {"value": batch.cells / t / 1e9, "unit": UNIT, "cores": int(nthreads), "kind": "port",
            "sample": f"full batch ({batch.n_pairs} pairs, {batch.cells / 1e9:.2f} Gcells) x{reps}, median; " + CPU_LABEL.format(isa=cpu_isa()),
            "seconds_per_pass": t, "fp64_pairs": int(nd)}


def parity_check(batch, out, used, raw=None, nthreads=0):
    """Scores `batch` with the CPU oracle and compares: the per-pair fallback decision and the raw FP32 sums against
    the float twin (SIMD port with IEEE subnormals, bit-identical to the scalar twin: tests/test_oracle.py), every
    result against the DOUBLE-precision oracle (the north_star's reference arithmetic, tolerance 1e-4) and the
    per-read best haplotype against the oracle's."""
    from oracle import oracle as O

    t0 = time.perf_counter()
    nthreads = nthreads or host_threads()
    o_ref, u_ref, r_ref, _ = O.batch_simd(batch, nthreads, False)
    dbl = O.batch_double(batch, nthreads)
    fin = np.isfinite(o_ref)
    res = {"pairs": int(batch.n_pairs), "fallback_mismatches": int((used != u_ref).sum()),
           "raw_f32_bit_mismatches": int((raw.view(np.uint32) != r_ref.view(np.uint32)).sum()) if raw is not None else None,
           "finite_mismatches": int((np.isfinite(out) != fin).sum()),
           "max_abs_dlog10": float(np.abs(out[fin] - o_ref[fin]).max()) if fin.any() else 0.0,
           "max_abs_dlog10_vs_double_oracle": float(np.abs(out[fin] - dbl[fin]).max()) if fin.any() else 0.0,
           "tolerance": TOL}
    sel = (u_ref == 1) & fin
    res["fp64_pairs"] = int(u_ref.sum())
    res["max_abs_dlog10_fp64_pairs"] = float(np.abs(out[sel] - o_ref[sel]).max()) if sel.any() else 0.0
    bad = 0
    for g in range(batch.n_regions):
        nr, nh, o0 = int(batch.reg_nreads[g]), int(batch.reg_nhaps[g]), int(batch.reg_out0[g])
        if nr and nh:
            bad += int((out[o0:o0 + nr * nh].reshape(nr, nh).argmax(1) != o_ref[o0:o0 + nr * nh].reshape(nr, nh).argmax(1)).sum())
    res["argmax_mismatches"] = bad
    res["oracle"] = "oracle/pairhmm_cpu_simd.c (float twin, IEEE subnormals) + oracle/pairhmm_oracle.c (double), parity unpinned: see oracle header"
    res["seconds"] = time.perf_counter() - t0
    res["ok"] = (res["fallback_mismatches"] == 0 and res["finite_mismatches"] == 0 and res["argmax_mismatches"] == 0 and
                 (res["raw_f32_bit_mismatches"] in (0, None)) and res["max_abs_dlog10_vs_double_oracle"] <= TOL)
    return res


def run_reference(args, rank, world):
    """--impl reference: the reference path's CPU PairHMM on this box's host cores.  GATK + GKL cannot
    run here (no JVM, no jar), so this is the oracle's AVX/OpenMP port, labelled as such."""
    if rank != 0:
        return
    batch, desc = make_workload(args.workload, 0)
    from oracle import oracle as O

    O.load()
    nthreads = host_threads()
    for _ in range(max(1, min(args.warmup, 2))):
        O.batch_simd(batch, nthreads, True)
    ts = []
    for _ in range(args.steps):
        t = time.perf_counter()
        O.batch_simd(batch, nthreads, True)
        ts.append(time.perf_counter() - t)
    tot = float(np.sum(ts))
    val = batch.cells * args.steps / tot / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (+f64 rerun)", "data": "synthetic",
        "config": shared_config(args, desc, batch),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthreads, "kind": "port",
                         "sample": "one full batch per step on all host threads; " + CPU_LABEL.format(isa=cpu_isa())},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
def measure_config(hmm, hmm_raw, batch, name, n_sm, f_max, steps, want_parity=True):
    """One BASELINE config on one GPU: kernels-only on a resident batch, e2e through fcs_pairhmm_compute, parity."""
    from falcon_genome_b200 import RegionArray

    rb = hmm.resident(batch, 0)
    for _ in range(3):
        rb.run_timed()
    ts = [rb.run_timed() for _ in range(steps)]
    tot = float(np.median([t[0] for t in ts]))
    main = float(np.median([t[1] for t in ts]))
    out_res, used_res = rb.download()
    launches = rb.launches
    rb.close()
    ra = RegionArray(batch)
    for _ in range(2):
        hmm.compute_regions(batch, ra)
    e2 = []
    for _ in range(steps):
        t0 = time.perf_counter()
        hmm.compute_regions(batch, ra)
        e2.append(time.perf_counter() - t0)
    e2e = float(np.median(e2))
    same = bool(np.array_equal(ra.out, out_res) and np.array_equal(ra.used, used_res))
    # DP cells of the pairs that took the FP64 path (for the FP64-pipe fraction)
    rl = batch.rd_len.astype(np.int64)
    hl = batch.hp_len.astype(np.int64)
    c64 = 0
    for g in range(batch.n_regions):
        nr, nh, o0 = int(batch.reg_nreads[g]), int(batch.reg_nhaps[g]), int(batch.reg_out0[g])
        if nr and nh:
            u = used_res[o0:o0 + nr * nh].reshape(nr, nh)
            if u.any():
                r0, h0 = int(batch.reg_read0[g]), int(batch.reg_hap0[g])
                c64 += int((u * np.outer(rl[r0:r0 + nr], hl[h0:h0 + nh])).sum())
    peak32 = n_sm * 128 * f_max / 8.0
    peak64 = n_sm * 64 * f_max / 8.0
    f64_ms = max(tot - main, 0.0)
    res = {"workload": WORKLOADS[name], "pairs": int(batch.n_pairs), "cells": int(batch.cells), "value": batch.cells / (tot * 1e-3) / 1e9,
           "ms_per_step": tot, "fp32_phase_ms": main, "fp64_phase_ms": f64_ms, "launches_per_step": int(launches),
           "roofline": {"frac": batch.cells / (main * 1e-3) / 1e9 / peak32, "whole_step_frac": batch.cells / (tot * 1e-3) / 1e9 / peak32,
                        "bound": "fp32_fma", "peak": peak32},
           "fp64_pairs": int(used_res.sum()), "fp64_rate": float(used_res.mean()), "fp64_cells": int(c64),
           "e2e": {"value": batch.cells / e2e / 1e9, "ms_per_call": e2e * 1e3}, "e2e_equals_resident": same}
    if c64 and f64_ms > 0:
        g64 = c64 / (f64_ms * 1e-3) / 1e9
        res["fp64_kernel"] = {"gcups": g64, "frac_of_fp64_pipe": g64 / peak64, "peak": peak64,
                              "per_unit": "8 FP64-pipe instructions per cell; peak = n_SM * 64 lanes * f / 8"}
    if want_parity:
        out, used, raw = hmm_raw.compute_flat(batch, want_raw=True)
        res["parity"] = parity_check(batch, out, used, raw)
        res["parity"]["timed_results_identical"] = bool(np.array_equal(out, ra.out) and np.array_equal(used, ra.used))
        res["parity"]["ok"] = bool(res["parity"]["ok"] and res["parity"]["timed_results_identical"] and same)
    return res


def dispatcher_leg(n_dev, args, f_max, n_sm):
    """ONE library handle over all n_dev devices (Engine::compute: regions partitioned by cells, no exchange) fed a
    fixed config-3 stream (20 calls of 2000 regions) and config 4 (30 calls of 20 regions over 10 distinct batches) through
    fcs_pairhmm_compute with host buffers; wall clock.  The same stream at every N: strong scaling."""
    from falcon_genome_b200 import PairHMM, RegionArray

    procs = max(1, min(host_threads() - 2, 20))
    t0 = time.perf_counter()
    n3 = 20 if procs >= 8 else 4
    c3 = make_workloads_parallel([("c3", k) for k in range(n3)], procs)
    c4 = make_workloads_parallel([("c4", k) for k in range(10)], procs)
    gen_s = time.perf_counter() - t0
    res = {"devices": n_dev, "generation_s": gen_s, "host_threads": host_threads()}
    with PairHMM(devices=list(range(n_dev))) as hm:
        res["device_count"] = hm.device_count
        for key, batches, calls in (("c3_stream", c3, 20), ("c4", c4, 30)):
            callers = args.dispatcher_callers
            if calls > len(batches):  # a batch's result arrays are written by one call at a time
                callers = min(callers, len(batches))
            ras = [RegionArray(b) for b in batches]
            cells = sum(batches[k % len(batches)].cells for k in range(calls))
            # the calls are issued by `callers` threads, as GATK's native PairHMM threads (HTCWorker.cpp:85) or the JVMs
            # of a stage do; the library coalesces calls that arrive while a batch is on the devices (flat combining) and
            # overlaps a batch's planning with the previous batch's tail (cross-batch pipelining)
            nxt = [0]
            lock = threading.Lock()
            errs = []
            n_calls = [calls]

            def caller():
                try:
                    while True:
                        with lock:
                            k = nxt[0]
                            nxt[0] += 1
                        if k >= n_calls[0]:
                            return
                        hm.compute_regions(batches[k % len(batches)], ras[k % len(batches)])
                except Exception as e:  # noqa: BLE001
                    errs.append(repr(e))

            # warm-up with the same concurrency as the timed pass: merged batches cut larger chunks than single calls, and the
            # slot buffers must have grown to them before the clock starts (re-allocating pinned memory costs milliseconds)
            n_calls[0] = min(calls, 2 * callers)
            ths = [threading.Thread(target=caller) for _ in range(callers)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            if errs:
                raise RuntimeError("dispatcher leg (warm-up): " + errs[0])
            nxt[0] = 0
            n_calls[0] = calls
            hm.reset_stats()

            ths = [threading.Thread(target=caller) for _ in range(callers)]
            t0 = time.perf_counter()
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            dt = time.perf_counter() - t0
            if errs:
                raise RuntimeError("dispatcher leg: " + errs[0])
            st = hm.stats()
            res[key] = {"calls": calls, "callers": callers, "distinct_batches": len(batches), "regions_per_call": int(batches[0].n_regions), "cells": int(cells),
                        "pairs": int(sum(batches[k % len(batches)].n_pairs for k in range(calls))), "seconds": dt, "value": cells / dt / 1e9, "unit": UNIT,
                        "ms_per_call": dt / calls * 1e3, "chunks": int(st["chunks"]), "h2d_bytes": int(st["h2d_bytes"]), "d2h_bytes": int(st["d2h_bytes"]),
                        "host_ms_per_call": {k2: float(st[k2]) / calls for k2 in ("host_plan_ms", "host_pack_ms", "host_wait_ms", "host_scatter_ms")},
                        }
            res[key]["_out0"] = ras[0].out.copy()
            res[key]["_used0"] = ras[0].used.copy()
    # self-check: the N-device result of the first batch of each stream == the 1-device result, bit for bit, and == oracle
    with PairHMM(devices=[0], keep_raw_f32=True) as h1:
        for key, batches in (("c3_stream", c3), ("c4", c4)):
            o1, u1, r1 = h1.compute_flat(batches[0], want_raw=True)
            same = bool(np.array_equal(o1, res[key].pop("_out0")) and np.array_equal(u1, res[key].pop("_used0")))
            res[key]["equals_one_device_bitwise"] = same
            p = parity_check(batches[0], o1, u1, r1)
            res[key]["parity_first_call"] = p
            res[key]["ok"] = bool(same and p["ok"])
    res["ok"] = bool(res["c3_stream"]["ok"] and res["c4"]["ok"])
    res["note"] = ("strong scaling: the same stream at every N; one process, one handle, pack_threads = min(4, cores / N) per device, calls issued by "
                   "`callers` threads after a warm-up pass with the same concurrency; value = cells / wall seconds from the first call to the last return (generation and "
                   "RegionArray construction excluded)")
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config section (c1, c3, c4, c5; N = 1 only)")
    ap.add_argument("--no-dispatcher", action="store_true", help="skip the in-process multi-GPU dispatcher leg")
    ap.add_argument("--dispatcher-callers", type=int, default=4,
                    help="threads issuing the calls of the dispatcher leg (GATK's --native-pair-hmm-threads default is 4)")
    ap.add_argument("--preheat-s", type=float, default=2.0, help="untimed pre-heat of the same kernels right before the timed steps")
    ap.add_argument("--threads", type=int, default=0,
                    help="host packing threads per rank (the library's max_threads, GATK's --native-pair-hmm-threads); "
                         "0 = this rank's share of the host cores, between 1 and 4")
    ap.add_argument("--l2", default="rotate", choices=["rotate", "flush"],
                    help="how the timed steps are kept from re-using inputs out of L2: rotate over resident batches whose "
                         "total size exceeds L2 (default), or write a 192 MiB buffer between steps")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import _pkg

    _pkg.load()
    from falcon_genome_b200 import PairHMM, RegionArray

    if world == 1 and args.gpus > 1:
        # plain `python bench.py --gpus N`: relaunch one rank per GPU exactly as the driver does
        import socket
        import subprocess

        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    dev_index = local_rank if world > 1 else 0
    torch.cuda.set_device(dev_index)
    import datetime

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev_index), timeout=datetime.timedelta(minutes=30))
    # host-side barrier for the phases in which rank 0 alone drives every GPU: an NCCL barrier would leave a spinning
    # kernel on the other ranks' devices, time-sliced against the dispatcher's kernels
    cpu_group = dist.new_group(backend="gloo", timeout=datetime.timedelta(minutes=30)) if world > 1 else None

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    batch, desc = make_workload(args.workload, rank)
    # One rank per GPU shares the box's host cores with the other ranks: at most the rank's share of the cores, and at
    # most the library default of 4 (GATK's --native-pair-hmm-threads).  Measured on a 4-core affinity mask (what a
    # rank gets on an 8-GPU / 32-core box), config 2 end to end: 1 / 2 / 3 / 4 packing threads = 2579 / 2971 / 3262 / 3411
    # GCUPS -- the calling thread is one of the packers and the sampler sleeps, so four threads on four cores do not
    # oversubscribe; with fewer cores the library keeps the same chunk schedule and lets fewer threads pull from it.
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    share = host_threads() // max(1, local_world)
    pack_threads = args.threads if args.threads > 0 else max(1, min(4, share))
    hmm = PairHMM(devices=[dev_index], max_threads=pack_threads)
    res = [hmm.resident(batch, 0)]
    launches_per_step = res[0].launches
    flush = None
    if args.l2 == "flush":
        flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
        l2_note = "flushed between timed steps (192 MiB write)"
    else:
        # same workload, other seeds: step i runs resident batch i mod NB; together the batches are larger than
        # L2, so a batch's inputs have been evicted by the time its turn comes again
        per_batch = batch.input_bytes() + 9 * batch.n_pairs
        nb = int(np.ceil(160e6 / per_batch)) + 1
        for i in range(1, nb):
            other, _ = make_workload(args.workload, rank + 1000 * i)
            res.append(hmm.resident(other, 0))
        l2_note = (f"no flush: the timed steps rotate over {nb} resident batches of this workload (different seeds, "
                   f"{nb * per_batch / 1e6:.0f} MB of inputs+outputs > 126 MB L2)")
    cells_of = [r.cells for r in res]

    # ---- kernel-only steps (inputs resident in HBM) ------------------------------------
    step_no = [0]

    def step():
        r = res[step_no[0] % len(res)]
        step_no[0] += 1
        return r.run_timed()

    for _ in range(max(args.warmup, len(res))):
        if flush is not None:
            flush.fill_(1)
        step()
    sampler = ClockSampler(dev_index)
    sampler.start()
    time.sleep(0.05)
    barrier()
    # untimed pre-heat: the same kernels back to back for >= preheat_s, so that the timed steps run at the clock and
    # power state the device sustains under this load (the timed region itself lasts tens of milliseconds)
    t_heat0 = time.perf_counter()
    heat_steps = 0
    while time.perf_counter() - t_heat0 < args.preheat_s:
        step()
        heat_steps += 1
    t_heat1 = time.perf_counter()
    step_no[0] = 0
    barrier()
    wall0 = time.perf_counter()
    tot_ms = 0.0
    main_ms = 0.0
    cells_rank = 0
    for i in range(args.steps):
        if flush is not None:
            flush.fill_(0)  # L2 flush between timed iterations (untimed; the events bracket only the kernels)
            torch.cuda.synchronize()
        cells_rank += cells_of[i % len(res)]
        t, m = step()
        tot_ms += t
        main_ms += m
    barrier()
    wall1 = time.perf_counter()
    wall = wall1 - wall0
    t_max = reduce_max(tot_ms)
    cells_all = reduce_sum(cells_rank)
    value = cells_all / (t_max * 1e-3) / 1e9
    out_res, used_res = res[0].download()
    fp64_pairs = int(used_res.sum())

    # ---- end to end through the reference-facing C ABI call, host buffers -----------------
    ra = RegionArray(batch)
    for _ in range(3):
        hmm.compute_regions(batch, ra)
    hmm.reset_stats()
    barrier()
    e2e_w0 = time.perf_counter()
    e2e_t = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        hmm.compute_regions(batch, ra)
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    e2e_w1 = time.perf_counter()
    st = hmm.stats()
    e2e_tot = reduce_max(float(np.sum(e2e_t)))
    e2e_cells = batch.cells * (world if world > 1 else 1)
    e2e_value = e2e_cells * args.steps / e2e_tot / 1e9
    e2e_ms = np.array(e2e_t) * 1e3
    e2e_same = bool(np.array_equal(ra.out, out_res) and np.array_equal(ra.used, used_res))
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    # per-rank host-side phases of the e2e calls (names the limiter at N > 1)
    mine = [float(np.sum(e2e_t)) / args.steps * 1e3] + [float(st[k]) / args.steps for k in ("host_plan_ms", "host_pack_ms", "host_wait_ms", "host_scatter_ms")]
    if world > 1:
        tt = torch.tensor(mine, dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(allr, tt)
        per_rank = [[float(x) for x in a.tolist()] for a in allr]
    else:
        per_rank = [mine]

    # ---- parity of the measured batch (every rank checks its own; the line carries rank 0's and the AND of all) ----
    with PairHMM(devices=[dev_index], keep_raw_f32=True, max_threads=pack_threads) as hmm_raw:
        out_p, used_p, raw_p = hmm_raw.compute_flat(batch, want_raw=True)
        par = parity_check(batch, out_p, used_p, raw_p, nthreads=max(1, share))
        par["timed_results_identical"] = bool(np.array_equal(out_p, ra.out) and np.array_equal(used_p, ra.used) and e2e_same)
        par["ok"] = bool(par["ok"] and par["timed_results_identical"])
        all_ok = reduce_sum(0.0 if par["ok"] else 1.0) == 0.0
        par["all_ranks_ok"] = bool(all_ok)

        prop = torch.cuda.get_device_properties(dev_index)
        n_sm = prop.multi_processor_count
        peaks, peaks_src = measured_peaks()
        f_max = float(sampler.max_mhz or peaks.get("sm_max_mhz", 1965.0)) / 1e3

        configs = None
        if rank == 0 and world == 1 and args.gpus == 1 and not args.no_configs:
            configs = {}
            for name in ("c1", "c3", "c4", "c5"):
                if name == args.workload:
                    continue
                b2, _ = make_workload(name, 0)
                configs[name] = measure_config(hmm, hmm_raw, b2, name, n_sm, f_max, steps=5)

    dispatcher = None
    if not args.no_dispatcher:
        host_barrier()
        if rank == 0:
            dispatcher = dispatcher_leg(max(world, 1), args, f_max, n_sm)
        host_barrier()

    if rank == 0:
        heat = sampler.window(t_heat0, t_heat1)
        timed = sampler.window(wall0, wall1)
        load = sampler.window(t_heat0 + min(0.5, args.preheat_s / 2), wall1)  # pre-heat (after its first half second) + timed steps
        e2e_clk = sampler.window(e2e_w0, e2e_w1)
        clocks = {"sm_mhz": load["sm_mhz"], "sm_max_mhz": sampler.max_mhz, "reasons": sorted(sampler.reasons), "samples": load["samples"],
                  "sm_mhz_min": load.get("sm_mhz_min"), "power_w_median": load.get("power_w_median"), "power_w_max": load.get("power_w_max"),
                  "window": f"pre-heat ({t_heat1 - t_heat0:.2f} s, {heat_steps} untimed steps of the same kernels) + the {args.steps} timed steps",
                  "timed_region_only": timed, "e2e_region": e2e_clk, "sampling_period_ms": sampler.period * 1e3}
        if not sampler.ok:
            clocks["note"] = "NVML unavailable"
        peak_gcups = n_sm * 128 * f_max / 8.0  # 8 FMA-pipe instructions per cell
        main_gcups = cells_rank / (main_ms * 1e-3) / 1e9  # dominant kernel: FP32 wavefront, this rank
        f_sus = (clocks.get("sm_mhz") or f_max * 1e3) / 1e3
        alg_bytes = batch.input_bytes() + 9 * batch.n_pairs  # 5 B/read base + 1 B/hap base in, 8 B + 1 B per pair out
        traffic, traffic_src = ncu_traffic(args.workload)
        cfg = shared_config(args, desc, batch)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": max(world, args.gpus), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (+f64 rerun)", "data": "synthetic", "config": cfg,
            "run": {"fp64_rerun_pairs": fp64_pairs, "l2": l2_note, "host_pack_threads_per_rank": pack_threads, "host_threads": host_threads(),
                    "timing": "CUDA events on the library's launching stream around the kernels of each step, summed; max over ranks"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(st["h2d_bytes"] // args.steps),
                    "d2h_bytes_per_step": int(st["d2h_bytes"] // args.steps), "ms_per_step": e2e_tot / args.steps * 1e3,
                    "ms_min_median_max_rank0": [float(e2e_ms.min()), float(np.median(e2e_ms)), float(e2e_ms.max())],
                    "per_rank_ms_per_call": {"columns": ["call", "host_plan", "host_pack", "host_wait", "host_scatter"], "rows": per_rank,
                                             "note": "host phases are summed over the rank's packing threads (they overlap each other and the device)"},
                    "call": "fcs_pairhmm_compute(handle, regions, n_regions): pack from caller pointers -> pinned staging -> H2D -> kernels -> D2H -> scatter"},
            "gpu_launches": int(launches_per_step * args.steps),
            "parity": par,
            "roofline": {"bound": "fp32_fma", "achieved": main_gcups, "peak": peak_gcups, "unit": UNIT, "frac": main_gcups / peak_gcups,
                         "frac_at_sustained_clock": main_gcups / (n_sm * 128 * f_sus / 8.0), "sustained_sm_mhz": clocks.get("sm_mhz"),
                         "whole_step_frac": (cells_rank / (tot_ms * 1e-3) / 1e9) / peak_gcups,
                         "kernel": "phmm_f32a_tier2 on config 2 (all-uniform form, G=4 R=38) / phmm_f32*_tier* (FP32 wavefront, run_task<float,G,R,FORM>)",
                         "per_unit": "8 FMA-pipe instructions (4 FFMA + 4 FMUL, 12 FLOP) per DP cell", "n_sm": n_sm, "f_max_ghz": f_max,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes": alg_bytes, "hbm": {"achieved_gbs": alg_bytes * args.steps / (t_max * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                                                  "peak_source": peaks_src, "note": "HBM is non-binding for this path"}},
            "clocks": clocks, "wall_s_timed_region": wall,
        }
        if configs is not None:
            line["configs"] = configs
        if dispatcher is not None:
            line["e2e_dispatcher"] = dispatcher
        if not args.no_cpu_baseline and world == 1 and args.gpus == 1:
            line["cpu_baseline"] = cpu_baseline(batch)
        print(json.dumps(line), flush=True)
    for r in res:
        r.close()
    hmm.done()
    if world > 1:
        dist.destroy_process_group()
    ok = par["all_ranks_ok"] and (configs is None or all(c["parity"]["ok"] for c in configs.values())) and (dispatcher is None or dispatcher["ok"])
    if not ok:
        sys.stderr.write("bench.py: PARITY FAILURE (see the parity objects of the JSON line)\n")
        raise SystemExit(3)


if __name__ == "__main__":
    main()
