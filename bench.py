#!/usr/bin/env python
"""bench.py — PairHMM GCUPS on N B200s (BASELINE.json metric) with roofline and CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # CPU PairHMM on the box's host cores

One "step" = one pass of the PairHMM forward path over one synthetic batch of the workload
(default: BASELINE config 2 — 100 000 pairs, 150 bp reads x 300 bp haplotypes, uniform quals).
  value      GCUPS, kernels only, inputs already resident in HBM (CUDA events on the library's
             launching stream, steps rotate over resident batches totalling more than L2 (or --l2 flush), max over ranks)
  e2e        same metric through the reference-facing call fcs_pairhmm_compute() with HOST buffers:
             host packing + H2D + kernels + D2H + scatter inside the timed region
  roofline   FP32 FMA pipe: peak GCUPS = n_SM * 128 * f_max / 8
  cpu_baseline  the oracle's AVX/OpenMP port timed on this box's host cores (rank 0, N=1)

Under torchrun (N > 1) every rank owns one GPU and scores its own batch of the same shape
(regions are independent: no collective on the data path; torch.distributed is used only for
the barrier and the max-over-ranks reduction) -> "scaling": "weak".
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

# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed ncu capture
NCU_DRAM_BYTES_PER_LAUNCH = {"c2": 4496896}

METRIC = "pairhmm_gcups"
UNIT = "GCUPS"


def make_workload(name: str, rank: int):
    import _pkg

    _pkg.load()
    from falcon_genome_b200 import synth

    if name == "c2":
        return synth.config2_uniform(seed=2002 + rank), "config2: 100k pairs, 150bp reads x 300bp haplotypes, uniform quals (q30/i45/d45/c10)"
    if name == "c2b":
        return synth.config2_uniform(seed=2002 + rank, random_quals=True), "config2b: 100k pairs 150x300, base quals U[6,41]"
    if name == "c1":
        return synth.config1_golden(seed=1001 + rank), "config1 stand-in: 400 simulated active regions"
    if name == "c3":
        return synth.config3_wgs(n_regions=2000, seed=3003, chunk=rank), "config3 chunk: 2000 WGS-shaped regions, reads 100-250 x haps 100-600"
    if name == "c4":
        return synth.config4_mutect2(n_regions=20, seed=4004 + rank), "config4 sample: 20 Mutect2-shaped regions"
    if name == "c5":
        return synth.config5_underflow(seed=5005 + rank), "config5: underflow stress 250bp x 1kb, 20k pairs"
    raise SystemExit(f"unknown workload {name}")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.sm = []
        self.reasons = set()
        self.max_mhz = None
        self.power = []
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
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
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
                time.sleep(0.02)
        except Exception:
            self.ok = False

    def summary(self):
        if not self.ok or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "power_w_max": max(self.power) if self.power else None}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def host_threads() -> int:
    """All host cores this process may run on.  Not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
    to its ranks, which would make the CPU reference arm single-threaded at N > 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


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
    lib = O.load()
    import ctypes

    lib.phmm_cpu_isa.restype = ctypes.c_char_p
    return {"value": batch.cells / t / 1e9, "unit": UNIT, "cores": int(nthreads), "kind": "port",
            "sample": f"full batch ({batch.n_pairs} pairs, {batch.cells / 1e9:.2f} Gcells) x{reps}, median; "
                      f"AVX/OpenMP C port of the oracle ({lib.phmm_cpu_isa().decode()}), float-first + double rerun, FTZ on as GKL; "
                      "not GKL itself (no JVM/GATK in the image)",
            "seconds_per_pass": t, "fp64_pairs": int(nd)}


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
    import ctypes

    lib = O.load()
    lib.phmm_cpu_isa.restype = ctypes.c_char_p
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": {"workload": desc, "pairs_per_step": batch.n_pairs, "cells_per_step": batch.cells, "host_threads": nthreads},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthreads, "kind": "port",
                         "sample": f"one full batch per step; AVX/OpenMP C port ({lib.phmm_cpu_isa().decode()}), FTZ on as GKL; not GKL itself"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--threads", type=int, default=0,
                    help="host packing threads per rank (the library's max_threads, GATK's --native-pair-hmm-threads); "
                         "0 = this rank's share of the host cores, at most the library default of 4")
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev_index))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batch, desc = make_workload(args.workload, rank)
    # One rank per GPU shares the box's host cores with the other ranks: more packing threads than the rank's
    # share of cores only makes them (and the event waits) fight for the same cores.
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    pack_threads = args.threads if args.threads > 0 else max(1, min(4, host_threads() // max(1, local_world)))
    hmm = PairHMM(devices=[dev_index], max_threads=pack_threads)
    res = [hmm.resident(batch, 0)]
    launches_per_step = res[0].launches
    flush = None
    l2_note = ""
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
    step_no[0] = 0
    sampler = ClockSampler(dev_index)
    sampler.start()
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
    wall = time.perf_counter() - wall0
    t_max = tot_ms
    if world > 1:
        tt = torch.tensor([tot_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_max = float(tt.item())
    cells_all = cells_rank
    if world > 1:
        cc = torch.tensor([float(cells_rank)], dtype=torch.float64, device="cuda")
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        cells_all = float(cc.item())
    value = cells_all / (t_max * 1e-3) / 1e9
    out, used = res[0].download()
    fp64_pairs = int(used.sum())

    # ---- end to end through the reference-facing C ABI call, host buffers -----------------
    ra = RegionArray(batch)
    for _ in range(2):
        hmm.compute_regions(batch, ra)
    hmm.reset_stats()
    barrier()
    e2e_t = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        hmm.compute_regions(batch, ra)
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    st = hmm.stats()
    e2e_tot = float(np.sum(e2e_t))
    if world > 1:
        tt = torch.tensor([e2e_tot], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_tot = float(tt.item())
    e2e_cells = batch.cells * (world if world > 1 else 1)
    e2e_value = e2e_cells * args.steps / e2e_tot / 1e9
    e2e_ms = np.array(e2e_t) * 1e3
    assert np.array_equal(ra.out, out), "e2e and resident results differ"
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    if rank == 0:
        prop = torch.cuda.get_device_properties(dev_index)
        peaks, peaks_src = measured_peaks()
        clocks = sampler.summary()
        f_max = float(clocks.get("sm_max_mhz") or peaks.get("sm_max_mhz", 1965.0)) / 1e3
        n_sm = prop.multi_processor_count
        peak_gcups = n_sm * 128 * f_max / 8.0  # 8 FMA-pipe instructions per cell
        main_gcups = cells_rank / (main_ms * 1e-3) / 1e9  # dominant kernel: FP32 wavefront, this rank
        f_sus = (clocks.get("sm_mhz") or f_max * 1e3) / 1e3
        alg_bytes = batch.input_bytes() + 9 * batch.n_pairs  # 5 B/read base + 1 B/hap base in, 8 B + 1 B per pair out
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": max(world, args.gpus), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (+f64 rerun)", "data": "synthetic",
            "config": {"workload": desc, "pairs_per_step_per_gpu": batch.n_pairs, "cells_per_step_per_gpu": batch.cells,
                       "fp64_rerun_pairs": fp64_pairs, "l2": l2_note, "host_pack_threads_per_rank": pack_threads,
                       "parallelism": f"{max(world, args.gpus)} x independent region shards, no collective",
                       "timing": "CUDA events on the library's launching stream around the kernels of each step, summed; max over ranks"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(st["h2d_bytes"] // args.steps),
                    "d2h_bytes_per_step": int(st["d2h_bytes"] // args.steps), "ms_per_step": e2e_tot / args.steps * 1e3,
                    "ms_min_median_max_rank0": [float(e2e_ms.min()), float(np.median(e2e_ms)), float(e2e_ms.max())],
                    "call": "fcs_pairhmm_compute(handle, regions, n_regions): pack from caller pointers -> pinned staging -> H2D -> kernels -> D2H -> scatter"},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": {"bound": "fp32_fma", "achieved": main_gcups, "peak": peak_gcups, "unit": UNIT, "frac": main_gcups / peak_gcups,
                         "frac_at_sustained_clock": main_gcups / (n_sm * 128 * f_sus / 8.0), "kernel": "phmm_f32a_tier2 on config 2 (all-uniform form, G=4 R=38) / phmm_f32*_tier* (FP32 wavefront, run_task<float,G,R,FORM>)",
                         "per_unit": "8 FMA-pipe instructions (4 FFMA + 4 FMUL, 12 FLOP) per DP cell", "n_sm": n_sm, "f_max_ghz": f_max,
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(args.workload), "traffic_note": "dram__bytes_read+write of one FP32-kernel launch, "
                         "ncu --set full capture profiles/r01_c2_fp32a_tier2_full.md (config 2 only; the all-uniform kernel reads bases + base quals + haplotypes once, "
                         "the leftover launch the rest: no re-reads)",
                         "algorithmic_bytes": alg_bytes, "hbm": {"achieved_gbs": alg_bytes * args.steps / (t_max * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                                                  "peak_source": peaks_src, "note": "HBM is non-binding for this path"}},
            "clocks": clocks, "wall_s_timed_region": wall,
        }
        if not args.no_cpu_baseline and world == 1 and args.gpus == 1:
            line["cpu_baseline"] = cpu_baseline(batch)
        print(json.dumps(line), flush=True)
    for r in res:
        r.close()
    hmm.done()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
