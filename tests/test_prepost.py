"""f2: GATK-side pre/post-processing (host functions of the library) against an independent
restatement in plain Python of the published GATK4 algorithm [upstream, SURVEY A.6]."""
import math

import numpy as np

from falcon_genome_b200.prepost import PCR_CONSERVATIVE, PCR_HOSTILE, PCR_NONE, finalize_region, prepare_read


def py_count(unit, seq, leading):
    n, u, reps = len(seq), len(unit), 0
    if leading:
        s = 0
        while s + u <= n and seq[s:s + u] == unit:
            reps += 1
            s += u
    else:
        e = n
        while e - u >= 0 and seq[e - u:e] == unit:
            reps += 1
            e -= u
    return reps


def py_repeat_len(b, offset):
    max_bw, best_bw = 0, b[offset:offset + 1]
    for s in range(1, 9):
        if offset + 1 - s < 0:
            break
        unit = b[offset - s + 1:offset + 1]
        max_bw = py_count(unit, b[:offset + 1], False)
        if max_bw > 1:
            best_bw = unit
            break
    max_rl = max_bw
    if offset < len(b) - 1:
        best_fw, max_fw = b[offset + 1:offset + 2], 0
        for s in range(1, 9):
            if offset + s + 1 > len(b):
                break
            unit = b[offset + 1:offset + s + 1]
            max_fw = py_count(unit, b[offset + 1:], True)
            if max_fw > 1:
                best_fw = unit
                break
        if best_fw == best_bw:
            max_rl = max_bw + max_fw
        else:
            max_rl = max_fw + py_count(best_fw, b[:offset + 1], False)
    return min(max_rl, 20)


def py_prepare(bases, quals, mapq, model):
    n = len(bases)
    q = [min(x, mapq) if mapq >= 0 else x for x in quals]
    q = [6 if x < 18 else x for x in q]
    i, d = [45] * n, [45] * n
    if model:
        rate = {1: 1.0, 2: 2.0, 3: 3.0}[model]
        cache = [max(10, int(40.0 - math.exp(r / (rate * math.pi)) + 1.0 + 0.5)) for r in range(21)]
        for k in range(1, n):
            rl = py_repeat_len(bases, k - 1)
            i[k - 1] = min(i[k - 1], cache[rl])
            d[k - 1] = min(d[k - 1], cache[rl])
    return bytes(q), bytes(i), bytes(d), bytes([10] * n)


def test_prepare_read_matches_restatement():
    rng = np.random.default_rng(3)
    seqs = [b"ACGTACGTAC", b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA", b"ACACACACACACGTTTTTTTTTGCA", b"GATTACA", b"A", b"CAGCAGCAGCAGCAGCAGTTTACGCGCGCGCGAT"]
    seqs += [bytes(rng.choice(list(b"ACGT"), int(rng.integers(2, 120))).astype(np.uint8)) for _ in range(30)]
    for b in seqs:
        quals = bytes(rng.integers(2, 42, len(b)).astype(np.uint8))
        for mapq in (-1, 60, 25):
            for model in (PCR_NONE, PCR_HOSTILE, PCR_CONSERVATIVE):
                got = prepare_read(b, quals, mapq, pcr_model=model)
                assert got[1:] == py_prepare(b, quals, mapq, model), (b, mapq, model)
    # homopolymer of length >= 20 under the conservative model: 40 - exp(20/(3*pi)) + 1 = 32.6 -> 33
    got = prepare_read(b"A" * 30, bytes([30] * 30), pcr_model=PCR_CONSERVATIVE)
    assert set(got[2][:-1]) == {33} and got[2][-1] == 45 and set(got[4]) == {10}
    assert prepare_read(b"ACGT", bytes([5, 17, 18, 40]), mapq=30)[1] == bytes([6, 6, 18, 30])


def test_finalize_region_caps_and_flags():
    m = np.array([[-1.0, -3.0, -9.0], [-12.0, -20.0, -13.0], [-7.9, -8.5, -30.0]])
    out, flags = finalize_region(m, [150, 150, 40])
    assert np.allclose(out[0], [-1.0, -3.0, -5.5]) and np.allclose(out[1], [-12.0, -16.5, -13.0]) and np.allclose(out[2], [-7.9, -8.5, -12.4])
    # poorly modelled: best < min(2, ceil(len*0.02)) * -4  -> 150 bp: -8, 40 bp: -4
    assert flags.tolist() == [0, 1, 1]
    assert np.array_equal(m[0], [-1.0, -3.0, -9.0])  # input untouched
