import os

import numpy as np

from falcon_genome_b200 import FlatBatch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLD, name))
    b = FlatBatch(z["read_bases"], z["read_q"], z["read_i"], z["read_d"], z["read_c"], z["rd_off"], z["rd_len"], z["hap_bases"],
                  z["hp_off"], z["hp_len"], z["reg_read0"], z["reg_nreads"], z["reg_hap0"], z["reg_nhaps"], z["reg_out0"], name=name)
    return b, z


def parse_kat(path=os.path.join(GOLD, "kat_closed_form.txt")):
    """GKL text testcases: hap read q i d c expected (quals ASCII+33)."""
    cases = []
    for ln in open(path):
        if ln.startswith("#") or not ln.strip():
            continue
        hap, read, q, i, d, c, exp = ln.split()
        dec = lambda s: bytes(ord(ch) - 33 for ch in s)  # noqa: E731
        cases.append(((read.encode(), dec(q), dec(i), dec(d), dec(c)), hap.encode(), float(exp)))
    return cases
