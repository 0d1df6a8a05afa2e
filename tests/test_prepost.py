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


# ---- a second restatement, written from the description of the algorithm rather than from the C++ -------------------
# Different control structure on purpose (anchored regular expressions for the tandem-repeat search, whole-array numpy
# operations for the quality rules, GATK's own order: PCR model first, then capMinimumReadQualities), so that a
# misreading shared by the C++ and the loop-for-loop mirror above would show up here.

def _re_trailing(unit: bytes, text: bytes) -> int:
    import re

    m = re.search(b"(?:" + re.escape(unit) + b")+$", text, re.S)
    return len(m.group(0)) // len(unit) if m else 0


def _re_leading(unit: bytes, text: bytes) -> int:
    import re

    m = re.match(b"(?:" + re.escape(unit) + b")+", text, re.S)
    return len(m.group(0)) // len(unit) if m else 0


def re_repeat_length(read: bytes, offset: int) -> int:
    """findTandemRepeatUnits(read, offset).getRight(): the repeat unit ENDING at `offset` is the shortest suffix of
    read[:offset+1] (1..8 bases) that occurs more than once in a row there; the unit STARTING at offset+1 likewise
    forward.  Without such a unit the single base at that position stands in and the count of the longest unit tried
    is kept.  Equal units: the two counts add up; otherwise the forward unit is also counted backwards from `offset`."""
    left, right = read[:offset + 1], read[offset + 1:]
    bw_unit, bw = read[offset:offset + 1], 0
    for k in range(1, min(8, len(left)) + 1):
        bw = _re_trailing(left[-k:], left)
        if bw > 1:
            bw_unit = left[-k:]
            break
    total = bw
    if right:
        fw_unit, fw = right[:1], 0
        for k in range(1, min(8, len(right)) + 1):
            fw = _re_leading(right[:k], right)
            if fw > 1:
                fw_unit = right[:k]
                break
        total = bw + fw if fw_unit == bw_unit else fw + _re_trailing(fw_unit, left)
    return min(total, 20)


def np_prepare(bases: bytes, quals: bytes, mapq: int, model: int, bam_ins=None, bam_del=None):
    n = len(bases)
    ins = np.frombuffer(bam_ins, np.uint8).astype(np.int64) if bam_ins is not None else np.full(n, 45)
    dele = np.frombuffer(bam_del, np.uint8).astype(np.int64) if bam_del is not None else np.full(n, 45)
    if model and n > 1:  # applyPCRErrorModel: positions 0 .. n-2
        rate = {1: 1.0, 2: 2.0, 3: 3.0}[model]
        adj = np.array([max(10, int(np.floor(40.0 - math.exp(r / (rate * math.pi)) + 1.0 + 0.5))) for r in range(21)])
        rl = np.array([re_repeat_length(bases, k) for k in range(n - 1)])
        ins[:-1] = np.minimum(ins[:-1], adj[rl])
        dele[:-1] = np.minimum(dele[:-1], adj[rl])
    q = np.frombuffer(quals, np.uint8).astype(np.int64)
    if mapq >= 0:
        q = np.minimum(q, mapq)
    q = np.where(q < 18, 6, q)             # capMinimumReadQualities: base qualities below the threshold -> MIN_USABLE_Q_SCORE
    ins = np.where(ins < 6, 6, ins)         # ... insertion / deletion qualities floored at MIN_USABLE_Q_SCORE
    dele = np.where(dele < 6, 6, dele)
    return tuple(bytes(x.astype(np.uint8)) for x in (q, ins, dele, np.full(n, 10)))


def test_prepare_read_matches_second_restatement():
    rng = np.random.default_rng(11)
    seqs = [b"CAAG", b"AA", b"TTCTTCCCC", b"ACACACACACACGTTTTTTTTTGCA", b"GATTACAGATTACAGATTACA", b"A" * 40, b"ACGTTGCAACGTTGCAACGTTGCAT"]
    # low-complexity random reads: short tandem repeats of every unit length are common
    for _ in range(40):
        unit = bytes(rng.choice(list(b"ACGT"), int(rng.integers(1, 10))).astype(np.uint8))
        s = unit * int(rng.integers(1, 7)) + bytes(rng.choice(list(b"ACGT"), int(rng.integers(0, 12))).astype(np.uint8))
        seqs.append(s[: int(rng.integers(2, len(s) + 1))] + bytes(rng.choice(list(b"AC"), int(rng.integers(0, 9))).astype(np.uint8)))
    for b in seqs:
        n = len(b)
        quals = bytes(rng.integers(2, 42, n).astype(np.uint8))
        bi = bytes(rng.integers(0, 50, n).astype(np.uint8))   # BAM BI / BD tags, some below MIN_USABLE_Q_SCORE
        bd = bytes(rng.integers(0, 50, n).astype(np.uint8))
        for mapq in (-1, 60, 13):
            for model in (PCR_NONE, PCR_HOSTILE, 2, PCR_CONSERVATIVE):
                assert prepare_read(b, quals, mapq, pcr_model=model)[1:] == np_prepare(b, quals, mapq, model), (b, mapq, model)
                assert prepare_read(b, quals, mapq, bam_ins=bi, bam_del=bd, pcr_model=model)[1:] == np_prepare(b, quals, mapq, model, bi, bd), (b, mapq, model)
    # BI / BD below 6 are floored (GATK capMinimumReadQualities), after the PCR model
    got = prepare_read(b"ACGTACGT", bytes([30] * 8), bam_ins=bytes([0, 3, 5, 6, 7, 45, 2, 60]), bam_del=bytes([9, 1, 45, 5, 6, 4, 33, 0]), pcr_model=PCR_NONE)
    assert got[2] == bytes([6, 6, 6, 6, 7, 45, 6, 60]) and got[3] == bytes([9, 6, 45, 6, 6, 6, 33, 6])
    # an isolated dinucleotide under the HOSTILE model: repeat length 2 -> 40 - exp(2/pi) + 1 = 39.1 -> 39 at the first A
    got = prepare_read(b"CAAG", bytes([30] * 4), pcr_model=PCR_HOSTILE)
    assert got[2][1] == 39 and re_repeat_length(b"CAAG", 1) == 2
