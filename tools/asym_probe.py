import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
import _pkg; _pkg.load()
from falcon_genome_b200 import PairHMM, synth
for (nreg, nr, nh) in ((1, 5920, 100), (1, 5920, 20), (1, 4736, 100), (8, 740, 100)):
    b = synth.config2_uniform(n_regions=nreg, reads_per_region=nr, haps_per_region=nh)
    with PairHMM(devices=[0]) as h:
        rb = h.resident(b)
        for _ in range(2): rb.run_timed()
        ts = [rb.run_timed() for _ in range(5)]
        m = np.median([t[1] for t in ts])
        print(f"regions {nreg} reads {nr} haps {nh}: {b.cells/1e9:.1f} Gcells main {m:.3f} ms -> {b.cells/m/1e6:.0f} GCUPS ({100*b.cells/m/1e6/4653:.1f}%) launches {rb.launches} HS_COLS={os.environ.get('FCS_PHMM_HS_COLS')}", flush=True)
        rb.close()
