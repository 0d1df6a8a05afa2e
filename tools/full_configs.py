"""BASELINE configs 3 and 4 at FULL size through the library, end to end with host buffers.

  python tools/full_configs.py --config c3 [--regions 250000] [--devices N] [--callers 4] [--parity-every 25]
  python tools/full_configs.py --config c4 [--regions 5000]   ...

config 3: "30x-WGS-shaped synthetic chr20 batch stream: reads 100-250bp, haplotypes 100-600bp, ~50M pairs" = 250 000 regions,
issued as 125 calls of 2000 regions (synth.config3_wgs chunks 0..124).  config 4: "Mutect2-shaped tumor/normal 100x synthetic
active regions" = 5000 regions, issued as 250 calls of 20 regions.  The calls go through ONE handle over `--devices` devices from
`--callers` threads (GATK's native PairHMM threads, /root/reference/src/workers/HTCWorker.cpp:85; the fan-out this replaces:
/root/reference/src/worker-htc.cpp:113-145).  Reported: GCUPS over the whole stream (wall clock, first call to last return),
clocks sampled through the run, and parity: every K-th call is re-scored by the CPU oracle (fallback decisions, tolerance,
per-read best haplotype: bench.parity_check), every call's result is compared bit for bit between a second pass and the first.
The synthetic stream is generated up front in worker processes (numpy generators are single-threaded): that is not timed.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3", choices=["c3", "c4"])
    ap.add_argument("--regions", type=int, default=0, help="regions of the stream (default: the full config: 250000 / 5000)")
    ap.add_argument("--devices", type=int, default=1)
    ap.add_argument("--callers", type=int, default=4)
    ap.add_argument("--parity-every", type=int, default=25)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import _pkg

    _pkg.load()
    from falcon_genome_b200 import PairHMM, RegionArray

    per_call = 2000 if a.config == "c3" else 20
    full = 250000 if a.config == "c3" else 5000
    n_calls = max(1, (a.regions or full) // per_call)
    procs = max(1, min(bench.host_threads() - 2, 24))
    t0 = time.perf_counter()
    batches = bench.make_workloads_parallel([(a.config, k) for k in range(n_calls)], procs)
    gen_s = time.perf_counter() - t0
    ras = [RegionArray(b) for b in batches]
    cells = int(sum(b.cells for b in batches))
    pairs = int(sum(b.n_pairs for b in batches))
    print(f"[{a.config}] {n_calls} calls x {per_call} regions = {n_calls * per_call} regions, {pairs} pairs, {cells / 1e12:.3f} Tcells"
          f" (generated in {gen_s:.1f} s on {procs} processes)", flush=True)
    res = {"config": a.config, "regions": n_calls * per_call, "calls": n_calls, "regions_per_call": per_call, "pairs": pairs, "cells": cells,
           "devices": a.devices, "callers": a.callers, "host_threads": bench.host_threads(), "generation_s": gen_s}
    sampler = bench.ClockSampler(0, 0.02)
    sampler.start()
    with PairHMM(devices=list(range(a.devices))) as hm:
        res["device_count"] = hm.device_count

        def run_pass(calls):
            nxt = [0]
            lock = threading.Lock()
            errs = []

            def caller():
                try:
                    while True:
                        with lock:
                            k = nxt[0]
                            nxt[0] += 1
                        if k >= calls:
                            return
                        hm.compute_regions(batches[k], ras[k])
                except Exception as e:  # noqa: BLE001
                    errs.append(repr(e))

            ths = [threading.Thread(target=caller) for _ in range(a.callers)]
            t = time.perf_counter()
            for th in ths:
                th.start()
            for th in ths:
                th.join()
            dt = time.perf_counter() - t
            if errs:
                raise RuntimeError(errs[0])
            return t, dt

        run_pass(min(n_calls, 2 * a.callers))  # warm-up: slot buffers grow to the merged batches' chunks
        hm.reset_stats()
        t_start, dt = run_pass(n_calls)
        st = hm.stats()
        first = [(ra.out.copy(), ra.used.copy()) for ra in ras]
        res.update({"seconds": dt, "value": cells / dt / 1e9, "unit": "GCUPS", "ms_per_call": dt / n_calls * 1e3, "chunks": int(st["chunks"]),
                    "fp64_pairs": int(st["fp64_pairs"]), "h2d_bytes": int(st["h2d_bytes"]), "d2h_bytes": int(st["d2h_bytes"]),
                    "host_ms_per_call": {k: float(st[k]) / n_calls for k in ("host_plan_ms", "host_pack_ms", "host_wait_ms", "host_scatter_ms")}})
        res["clocks"] = sampler.window(t_start, t_start + dt)
        res["clocks"]["reasons"] = sorted(sampler.reasons)
        res["clocks"]["sm_max_mhz"] = sampler.max_mhz
        # second pass: every result bit for bit equal to the first one (chunking and merging differ from run to run)
        _, dt2 = run_pass(n_calls)
        same = all(np.array_equal(ra.out, f[0]) and np.array_equal(ra.used, f[1]) for ra, f in zip(ras, first))
        res["second_pass"] = {"seconds": dt2, "value": cells / dt2 / 1e9, "bitwise_equal_to_first": bool(same)}
    sampler.stop_flag.set()
    # oracle parity on every K-th call
    checked, ok = [], True
    for k in range(0, n_calls, max(1, a.parity_every)):
        p = bench.parity_check(batches[k], first[k][0], first[k][1])
        checked.append({"call": k, "pairs": p["pairs"], "fallback_mismatches": p["fallback_mismatches"], "argmax_mismatches": p["argmax_mismatches"],
                        "max_abs_dlog10_vs_double_oracle": p["max_abs_dlog10_vs_double_oracle"], "fp64_pairs": p["fp64_pairs"], "ok": p["ok"]})
        ok = ok and p["ok"]
        print(f"  parity call {k}: {p['pairs']} pairs, fallback mismatches {p['fallback_mismatches']}, argmax mismatches {p['argmax_mismatches']},"
              f" max |dlog10 L| vs double {p['max_abs_dlog10_vs_double_oracle']:.2e} -> {'ok' if p['ok'] else 'FAIL'}", flush=True)
    res["parity"] = {"calls_checked": len(checked), "pairs_checked": int(sum(c["pairs"] for c in checked)), "ok": bool(ok and same), "checks": checked,
                     "tolerance": 1e-4, "oracle": "oracle/pairhmm_cpu_simd.c (float twin) + oracle/pairhmm_oracle.c (double); parity unpinned, see oracle header"}
    line = json.dumps(res)
    print(line)
    if a.out:
        with open(a.out, "w") as f:
            f.write(line + "\n")
    if not res["parity"]["ok"]:
        raise SystemExit("parity failed")


if __name__ == "__main__":
    main()
