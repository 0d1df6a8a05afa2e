"""Generates tests/golden/* from the CPU oracle (oracle/pairhmm_oracle.c).

The reference repository holds no likelihood fixtures (parity unpinned), so
the golden set is (a) hand-derivable known answers in GKL's text testcase format and (b) a
small "config 1" stand-in batch scored by the oracle.  Re-run:  python tools/make_golden.py
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

_pkg.load()
from falcon_genome_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def q33(vals):
    return "".join(chr(v + 33) for v in vals)


def kat_lines():
    """hap read q i d c expected_log10  (quals ASCII+33, the GKL/GATK testcase text convention [upstream]).
    Expected values are closed forms, not oracle output."""
    L = []
    e30 = 10 ** -3.0
    L.append(("A", "A", [30], [45], [45], [10], math.log10((1 - e30) * 0.9)))
    L.append(("C", "A", [30], [45], [45], [10], math.log10(e30 / 3 * 0.9)))
    L.append(("AAAAAAA", "A", [30], [45], [45], [10], math.log10((1 - e30) * 0.9)))
    L.append(("N", "A", [30], [45], [45], [10], math.log10((1 - e30) * 0.9)))
    L.append(("G", "N", [30], [45], [45], [10], math.log10((1 - e30) * 0.9)))
    e20 = 10 ** -2.0
    L.append(("T", "T", [20], [30], [30], [20], math.log10((1 - e20) * (1 - 10 ** -2.0))))
    # quals are masked with & 127: 158 = 30 + 128 behaves as 30
    return L


def main():
    os.makedirs(GOLD, exist_ok=True)
    with open(os.path.join(GOLD, "kat_closed_form.txt"), "w") as f:
        f.write("# hap read q i d c expected_log10 ; quals ASCII+33 ; closed forms of SURVEY.md A.5 #1-#4\n")
        for hap, read, q, i, d, c, exp in kat_lines():
            f.write(f"{hap} {read} {q33(q)} {q33(i)} {q33(d)} {q33(c)} {exp:.15f}\n")
    # config-1 stand-in sample, oracle-scored
    b = synth.config1_golden(n_regions=5, seed=1001)
    out, used, raw, dbl = O.batch_scalar(b)
    np.savez_compressed(
        os.path.join(GOLD, "c1_sample.npz"),
        read_bases=b.read_bases, read_q=b.read_q, read_i=b.read_i, read_d=b.read_d, read_c=b.read_c, rd_off=b.rd_off, rd_len=b.rd_len,
        hap_bases=b.hap_bases, hp_off=b.hp_off, hp_len=b.hp_len, reg_read0=b.reg_read0, reg_nreads=b.reg_nreads, reg_hap0=b.reg_hap0,
        reg_nhaps=b.reg_nhaps, reg_out0=b.reg_out0, out_log10=out, used_fp64=used, raw_f32_bits=raw.view(np.uint32), log10_double=dbl)
    # underflow sample (forces the FP64 path)
    b5 = synth.config5_underflow(n_regions=1, reads_per_region=4, haps_per_region=3, read_len=120, hap_len=300, seed=5005)
    out, used, raw, dbl = O.batch_scalar(b5)
    np.savez_compressed(
        os.path.join(GOLD, "c5_sample.npz"),
        read_bases=b5.read_bases, read_q=b5.read_q, read_i=b5.read_i, read_d=b5.read_d, read_c=b5.read_c, rd_off=b5.rd_off, rd_len=b5.rd_len,
        hap_bases=b5.hap_bases, hp_off=b5.hp_off, hp_len=b5.hp_len, reg_read0=b5.reg_read0, reg_nreads=b5.reg_nreads, reg_hap0=b5.reg_hap0,
        reg_nhaps=b5.reg_nhaps, reg_out0=b5.reg_out0, out_log10=out, used_fp64=used, raw_f32_bits=raw.view(np.uint32), log10_double=dbl)
    # config-2 sample: constant, equal insertion/deletion qualities -> the all-uniform kernels (big enough that the
    # batcher leaves the latency policy and the tail window: 5120 pairs)
    b2 = synth.config2_uniform(n_regions=8, reads_per_region=64, haps_per_region=10, read_len=100, hap_len=120, seed=2002)
    out, used, raw, dbl = O.batch_scalar(b2)
    np.savez_compressed(
        os.path.join(GOLD, "c2_sample.npz"),
        read_bases=b2.read_bases, read_q=b2.read_q, read_i=b2.read_i, read_d=b2.read_d, read_c=b2.read_c, rd_off=b2.rd_off, rd_len=b2.rd_len,
        hap_bases=b2.hap_bases, hp_off=b2.hp_off, hp_len=b2.hp_len, reg_read0=b2.reg_read0, reg_nreads=b2.reg_nreads, reg_hap0=b2.reg_hap0,
        reg_nhaps=b2.reg_nhaps, reg_out0=b2.reg_out0, out_log10=out, used_fp64=used, raw_f32_bits=raw.view(np.uint32), log10_double=dbl)
    print("golden written:", sorted(os.listdir(GOLD)), "c1 pairs", b.n_pairs, "fp64", int(used.sum()))


if __name__ == "__main__":
    main()
