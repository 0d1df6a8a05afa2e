"""Developer micro-benchmark: resident-batch kernel timing of a synthetic config (GCUPS and
fraction of the FP32 FMA roofline).  Not the contract bench (that is bench.py)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

_pkg.load()
from falcon_genome_b200 import PairHMM, synth  # noqa: E402


def make(cfg, scale):
    if cfg == "c2":
        return synth.config2_uniform(n_regions=int(100 * scale))
    if cfg == "c2b":
        return synth.config2_uniform(n_regions=int(100 * scale), random_quals=True)
    if cfg == "c1":
        return synth.config1_golden(n_regions=int(400 * scale))
    if cfg == "c3":
        return synth.config3_wgs(n_regions=int(2000 * scale))
    if cfg == "c4":
        return synth.config4_mutect2(n_regions=max(1, int(20 * scale)))
    if cfg == "c5":
        return synth.config5_underflow(n_regions=max(1, int(200 * scale)))
    raise SystemExit(cfg)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="c2")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--e2e", action="store_true")
    a = ap.parse_args()
    t0 = time.time()
    b = make(a.cfg, a.scale)
    print(f"[{a.cfg}] {b.n_regions} regions {b.n_reads} reads {b.n_haps} haps {b.n_pairs} pairs {b.cells/1e9:.3f} Gcells (gen {time.time()-t0:.1f}s)"
          f" HS_COLS={os.environ.get('FCS_PHMM_HS_COLS','dflt')}", flush=True)
    with PairHMM(devices=[0]) as h:
        rb = h.resident(b)
        for _ in range(3):
            rb.run_timed()
        ts = [rb.run_timed() for _ in range(a.iters)]
        tot = np.array([t[0] for t in ts])
        main_ms = np.array([t[1] for t in ts])
        out, used = rb.download()
        peak = 148 * 128 * 1.965 / 8  # GCUPS at max clock
        g = b.cells / (np.median(tot) * 1e-3) / 1e9
        print(f"  kernels: median {np.median(tot):.3f} ms (min {tot.min():.3f}), main {np.median(main_ms):.3f} ms, launches {rb.launches}"
              f"  -> {g:.0f} GCUPS = {100*g/peak:.1f}% of FP32 roofline @1.965GHz; fp64 pairs {int(used.sum())}/{len(used)}", flush=True)
        rb.close()
        if a.e2e:
            from falcon_genome_b200 import RegionArray
            ra = RegionArray(b)
            for _ in range(3):
                h.compute_regions(b, ra)
            h.reset_stats()
            ts = []
            for _ in range(5):
                t = time.perf_counter()
                h.compute_regions(b, ra)
                ts.append(time.perf_counter() - t)
            t = float(np.median(ts))
            st = h.stats()
            n = 5.0
            print(f"  e2e compute(): median {t*1e3:.2f} ms -> {b.cells/t/1e9:.0f} GCUPS; per call: plan {st['host_plan_ms']/n:.2f} pack {st['host_pack_ms']/n:.2f} "
                  f"wait {st['host_wait_ms']/n:.2f} scatter {st['host_scatter_ms']/n:.2f} kernels {st['kernel_ms']/n:.2f} ms, chunks {st['chunks']/n:.0f}", flush=True)


if __name__ == "__main__":
    main()
