"""Benchmark of the daemon path (fcs-pairhmm-nam).

  python tools/nam_bench.py                          # one client, config 2, shared memory vs byte stream
  python tools/nam_bench.py --clients 32 --regions-per-call 1,8,64 [--seconds 3]

The second form is the reference's process model: up to 32 client processes (gatk.htc.nprocs,
/root/reference/src/config.cpp:56-82; fan-out /root/reference/src/worker-htc.cpp:113-145), each a stand-in for one
GATK JVM, issue small calls (one or a few active regions of the config-1 stand-in, as computeLikelihoodsNative does
per region) against ONE daemon that owns the GPU(s) (lifecycle /root/reference/src/BackgroundExecutor.cpp:13-84).
Reported: aggregate GCUPS, calls/s, per-call latency p50 / p99 over all clients, and how many device batches the
daemon formed (flat combining merges calls that arrive from different connections while a batch is on the device:
chunks << calls).  Every client checks its results against the in-process library result of the same regions.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

_pkg.load()
from falcon_genome_b200 import synth  # noqa: E402
from falcon_genome_b200.remote import NamDaemon, RemotePairHMM  # noqa: E402


def one_client_transports(a):
    b = synth.config2_uniform(n_regions=a.regions)
    sock = os.path.join(tempfile.mkdtemp(), "nam.sock")
    with NamDaemon(sock, devices=1):
        ref = None
        for shm in ("1", "0"):
            os.environ["FCS_PHMM_REMOTE_SHM"] = shm
            with RemotePairHMM(sock) as c:
                for _ in range(3):
                    out, used = c.compute_flat(b)
                ts = []
                for _ in range(a.iters):
                    t0 = time.perf_counter()
                    out, used = c.compute_flat(b)
                    ts.append(time.perf_counter() - t0)
                assert c.uses_shm == (shm == "1")
            if ref is None:
                ref = out
            assert np.array_equal(ref, out)
            t = float(np.median(ts))
            print(f"{'shared memory' if shm == '1' else 'byte stream  '}: median {t * 1e3:.2f} ms per call of {b.n_pairs} pairs "
                  f"({b.input_bytes() / 1e6:.1f} MB in) -> {b.cells / t / 1e9:.0f} GCUPS through the daemon", flush=True)


def _client(rank, sock, k, seconds, start_evt, q):
    """One stand-in JVM: its own slice of config-1 regions, calls of k regions each, closed loop."""
    b = synth.config1_golden(n_regions=max(64, 4 * k), seed=9000 + rank)
    groups = [b.select(list(range(g, min(g + k, b.n_regions)))) for g in range(0, b.n_regions - k + 1, k)]
    lat, cells, calls = [], 0, 0
    with RemotePairHMM(sock) as c:
        outs = [c.compute_flat(g)[0] for g in groups[:2]]  # connect, attach the segment, warm up
        start_evt.wait()
        t_end = time.perf_counter() + seconds
        i = 0
        first = {}
        while time.perf_counter() < t_end:
            g = groups[i % len(groups)]
            t0 = time.perf_counter()
            out, _ = c.compute_flat(g)
            lat.append(time.perf_counter() - t0)
            cells += g.cells
            calls += 1
            if i < len(groups):
                first[i] = out
            i += 1
    q.put({"rank": rank, "lat": lat, "cells": cells, "calls": calls, "first": {j: o.tolist() for j, o in list(first.items())[:3]},
           "seed": 9000 + rank, "k": k})


def many_clients(a):
    from falcon_genome_b200 import PairHMM

    sock = os.path.join(tempfile.mkdtemp(), "nam.sock")
    results = []
    ctx = mp.get_context("spawn")
    with NamDaemon(sock, devices=a.devices) as d:
        for k in [int(x) for x in a.regions_per_call.split(",")]:
            q = ctx.Queue()
            start = ctx.Event()
            procs = [ctx.Process(target=_client, args=(r, sock, k, a.seconds, start, q)) for r in range(a.clients)]
            for p in procs:
                p.start()
            time.sleep(a.settle)  # every client has connected and warmed up
            t0 = time.perf_counter()
            start.set()
            outs = [q.get() for _ in procs]
            wall = time.perf_counter() - t0
            for p in procs:
                p.join()
            lat = np.concatenate([np.asarray(o["lat"]) for o in outs]) * 1e6
            cells = sum(o["cells"] for o in outs)
            calls = sum(o["calls"] for o in outs)
            # parity of a sample: client 0's first calls against the in-process library
            with PairHMM(devices=[0]) as h:
                o0 = [o for o in outs if o["rank"] == 0][0]
                b = synth.config1_golden(n_regions=max(64, 4 * k), seed=o0["seed"])
                for j, vals in o0["first"].items():
                    g = b.select(list(range(int(j) * k, int(j) * k + k)))
                    ref, _ = h.compute_flat(g)
                    assert np.array_equal(ref, np.asarray(vals)), "daemon result differs from the in-process library"
            r = {"clients": a.clients, "regions_per_call": k, "seconds": a.seconds, "calls": int(calls), "calls_per_s": calls / a.seconds,
                 "aggregate_gcups": cells / a.seconds / 1e9, "latency_us": {"p50": float(np.percentile(lat, 50)), "p90": float(np.percentile(lat, 90)),
                                                                              "p99": float(np.percentile(lat, 99)), "max": float(lat.max())},
                 "wall_s": wall}
            results.append(r)
            print(json.dumps(r), flush=True)
    print("daemon: " + (d.proc.stdout.read() or "").strip())
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--regions", type=int, default=100)
    ap.add_argument("--iters", type=int, default=9)
    ap.add_argument("--clients", type=int, default=0)
    ap.add_argument("--regions-per-call", default="1,8,64")
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--settle", type=float, default=6.0)
    ap.add_argument("--devices", type=int, default=1)
    a = ap.parse_args()
    if a.clients > 0:
        many_clients(a)
    else:
        one_client_transports(a)


if __name__ == "__main__":
    main()
