"""Developer probe of the in-process dispatcher: a stream of distinct config-3 chunk calls through ONE handle,
per-call wall times with 1 and several caller threads (see bench.py dispatcher_leg)."""
import argparse
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
    ap.add_argument("--devices", type=int, default=1)
    ap.add_argument("--batches", type=int, default=8)
    ap.add_argument("--calls", type=int, default=16)
    ap.add_argument("--callers", default="1,4")
    ap.add_argument("--workload", default="c3")
    a = ap.parse_args()
    import _pkg

    _pkg.load()
    from falcon_genome_b200 import PairHMM, RegionArray

    batches = bench.make_workloads_parallel([(a.workload, k) for k in range(a.batches)], max(1, min(bench.host_threads() - 2, a.batches)))
    ras = [RegionArray(b) for b in batches]
    with PairHMM(devices=list(range(a.devices))) as hm:
        for k in range(len(batches)):
            hm.compute_regions(batches[k], ras[k])
        for callers in [int(x) for x in a.callers.split(",")]:
            callers = min(callers, len(batches))
            hm.reset_stats()
            nxt = [0]
            lock = threading.Lock()
            per = []

            def caller():
                while True:
                    with lock:
                        k = nxt[0]
                        nxt[0] += 1
                    if k >= a.calls:
                        return
                    t0 = time.perf_counter()
                    hm.compute_regions(batches[k % len(batches)], ras[k % len(batches)])
                    per.append((time.perf_counter() - t0) * 1e3)

            ths = [threading.Thread(target=caller) for _ in range(callers)]
            t0 = time.perf_counter()
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            dt = time.perf_counter() - t0
            cells = sum(batches[k % len(batches)].cells for k in range(a.calls))
            st = hm.stats()
            print(f"devices {a.devices} callers {callers}: {cells / dt / 1e9:.0f} GCUPS, {dt / a.calls * 1e3:.2f} ms per call (wall / calls), per-call latency "
                  f"median {np.median(per):.2f} max {max(per):.2f} ms, chunks {st['chunks']}, host plan {st['host_plan_ms'] / a.calls:.2f} pack {st['host_pack_ms'] / a.calls:.2f} "
                  f"wait {st['host_wait_ms'] / a.calls:.2f} scatter {st['host_scatter_ms'] / a.calls:.2f} ms per call", flush=True)


if __name__ == "__main__":
    main()
