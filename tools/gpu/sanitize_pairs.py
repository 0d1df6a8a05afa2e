"""compute-sanitizer target: a small call that runs the haplotype-pair kernels (odd and even haplotype counts, unequal
pair lengths, N, every read-blob layout), the scalar leftovers, the FP64 reruns and the chunked path, checked against the oracle.
  compute-sanitizer --tool memcheck python tools/gpu/sanitize_pairs.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

_pkg.load()
from falcon_genome_b200 import FlatBatch, PairHMM, Region, plan_check, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

rng = np.random.default_rng(5)
b1 = synth.config1_golden(n_regions=60, seed=21)
b3 = synth.config3_wgs(n_regions=60, seed=22)
b5 = synth.config5_underflow(n_regions=6)
for b in (b1, b3, b5):
    print(b.name, b.n_pairs, "pairs; pair tasks", plan_check(b)["n_tasks_hap_pairs"], flush=True)
with PairHMM(devices=[0], keep_raw_f32=True) as h:
    for b in (b1, b3, b5):
        out, used, raw = h.compute_flat(b, want_raw=True)
        ref, uref, rref, _ = O.batch_simd(b, 0, False)
        assert np.array_equal(used, uref) and np.array_equal(raw.view(np.uint32), rref.view(np.uint32)), b.name
        out2, used2 = h.compute_regions(b)
        assert np.array_equal(out, out2) and np.array_equal(used, used2)
print("sanitize target ok", flush=True)
