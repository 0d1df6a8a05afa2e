"""Per-call latency of fcs_pairhmm_compute for GATK-sized calls (one active region per call, as the GKL
JNI contract has it) and for small multi-region calls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import _pkg; _pkg.load()
from falcon_genome_b200 import PairHMM, RegionArray, synth

full = synth.config1_golden(n_regions=64, seed=5)
with PairHMM(devices=[0]) as h:
    for nreg in (1, 4, 16, 64):
        subs = [full.select(range(i, i + nreg)) for i in range(0, 64 - nreg + 1, nreg)][:16]
        ras = [RegionArray(s) for s in subs]
        for s, ra in zip(subs, ras):
            h.compute_regions(s, ra)
        ts = []
        for _ in range(5):
            for s, ra in zip(subs, ras):
                t = time.perf_counter(); h.compute_regions(s, ra); ts.append(time.perf_counter() - t)
        pairs = np.mean([s.n_pairs for s in subs]); cells = np.mean([s.cells for s in subs])
        print(f"{nreg:3d} region(s)/call: {pairs:8.0f} pairs {cells/1e6:8.1f} Mcells  median {np.median(ts)*1e6:8.1f} us  p90 {np.percentile(ts,90)*1e6:8.1f} us -> {cells/np.median(ts)/1e9:7.1f} GCUPS")

# concurrent callers (GATK's native PairHMM threads): one-region calls from T threads on one handle
import threading
with PairHMM(devices=[0]) as h:
    subs = [full.select([i]) for i in range(64)]
    ras = [RegionArray(s) for s in subs]
    for s, ra in zip(subs, ras):
        h.compute_regions(s, ra)
    for T in (1, 2, 4, 8, 16):
        def work(tid):
            for rep in range(20):
                for k in range(tid, 64, T):
                    h.compute_regions(subs[k], ras[k])
        th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        dt = time.perf_counter() - t0
        calls = 20 * 64
        print(f"{T:2d} caller thread(s), one region per call: {calls/dt:8.0f} calls/s  ({dt/calls*1e6*T:7.1f} us per call per thread)  chunks {h.stats()['chunks']}")
        h.reset_stats()
