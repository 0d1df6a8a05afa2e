"""Developer benchmark of the daemon path: one client process scoring a synthetic batch through
fcs-pairhmm-nam, shared-memory transport against the byte-stream protocol (FCS_PHMM_REMOTE_SHM=0)."""
import argparse
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--regions", type=int, default=100)
    ap.add_argument("--iters", type=int, default=9)
    a = ap.parse_args()
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


if __name__ == "__main__":
    main()
