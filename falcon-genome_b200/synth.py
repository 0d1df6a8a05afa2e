"""Deterministic synthetic PairHMM workloads — the five BASELINE.json configs
(shapes, value distributions and seeds).

There is no GATK in this image, so "config 1" (testcases captured from `fcs-genome htc`,
/root/reference/src/worker-htc.cpp:19-181) is replaced by an active-region simulator whose
regions look like what HaplotypeCaller hands to PairHMM: a handful of haplotypes that differ
by a few SNPs/indels, and reads sampled from them with sequencing errors.
PRNG: numpy Generator(PCG64(seed)).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from .batch import FlatBatch

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


class _Builder:
    """Accumulates regions straight into flat planes (no per-read Python objects)."""

    def __init__(self):
        self.rb, self.rq, self.ri, self.rd, self.rc, self.rlen = [], [], [], [], [], []
        self.hb, self.hlen = [], []
        self.reg_read0, self.reg_nreads, self.reg_hap0, self.reg_nhaps, self.reg_out0 = [], [], [], [], []
        self.nread = 0
        self.nhap = 0
        self.out0 = 0

    def add_region(self, reads, haps):
        """reads: list of (bases, q, i, d, c) uint8 arrays; haps: list of uint8 arrays."""
        self.reg_read0.append(self.nread)
        self.reg_nreads.append(len(reads))
        self.reg_hap0.append(self.nhap)
        self.reg_nhaps.append(len(haps))
        self.reg_out0.append(self.out0)
        self.out0 += len(reads) * len(haps)
        for (b, q, i, d, c) in reads:
            self.rb.append(b); self.rq.append(q); self.ri.append(i); self.rd.append(d); self.rc.append(c)
            self.rlen.append(len(b))
        self.nread += len(reads)
        for h in haps:
            self.hb.append(h)
            self.hlen.append(len(h))
        self.nhap += len(haps)

    def build(self, name: str, meta: Optional[dict] = None) -> FlatBatch:
        def cat(xs):
            return np.ascontiguousarray(np.concatenate(xs).astype(np.uint8)) if xs else np.zeros(0, np.uint8)

        rlen = np.asarray(self.rlen, dtype=np.int32)
        hlen = np.asarray(self.hlen, dtype=np.int32)
        rd_off = (np.cumsum(rlen.astype(np.int64)) - rlen).astype(np.int64)
        hp_off = (np.cumsum(hlen.astype(np.int64)) - hlen).astype(np.int64)
        return FlatBatch(
            cat(self.rb), cat(self.rq), cat(self.ri), cat(self.rd), cat(self.rc), rd_off, rlen,
            cat(self.hb), hp_off, hlen,
            np.asarray(self.reg_read0, np.int32), np.asarray(self.reg_nreads, np.int32),
            np.asarray(self.reg_hap0, np.int32), np.asarray(self.reg_nhaps, np.int32),
            np.asarray(self.reg_out0, np.int64), name=name, meta=meta or {},
        )


# ----------------------------------------------------------------------------------------
def _mutate_hap(rng, backbone: np.ndarray, n_snp: int, n_indel: int, indel_max: int, target_len: Optional[int] = None) -> np.ndarray:
    h = backbone.copy()
    for _ in range(n_snp):
        p = int(rng.integers(0, len(h)))
        h[p] = _ACGT[(int(np.searchsorted(_ACGT, h[p])) + int(rng.integers(1, 4))) % 4]
    for _ in range(n_indel):
        L = int(rng.integers(1, indel_max + 1))
        p = int(rng.integers(1, max(2, len(h) - L - 1)))
        if rng.random() < 0.5 and len(h) > L + 20:
            h = np.concatenate([h[:p], h[p + L:]])
        else:
            h = np.concatenate([h[:p], _ACGT[rng.integers(0, 4, L)], h[p:]])
    if target_len is not None:
        if len(h) > target_len:
            h = h[:target_len]
        elif len(h) < target_len:
            h = np.concatenate([h, _ACGT[rng.integers(0, 4, target_len - len(h))]])
    return np.ascontiguousarray(h)


def _homopolymer_runlen(b: np.ndarray) -> np.ndarray:
    """Length of the homopolymer run each base belongs to."""
    n = len(b)
    if n == 0:
        return np.zeros(0, np.int64)
    change = np.flatnonzero(b[1:] != b[:-1]) + 1
    starts = np.concatenate([[0], change])
    ends = np.concatenate([change, [n]])
    return np.repeat(ends - starts, ends - starts)


def _gatk_like_quals(rng, bases: np.ndarray, mean_q: int, lowq: bool = False):
    """Base quals: per-read mean with +-8 jitter clipped to [2,41], then GATK's
    pre-processing (q < 18 -> 6) [upstream, SURVEY A.6]; ins/del = 45 except inside
    homopolymer runs >= 4 where q = max(10, 45 - 3*run); gcp = 10."""
    n = len(bases)
    if lowq:
        q = rng.integers(2, 16, n)
        i = rng.integers(10, 21, n)
        d = rng.integers(10, 21, n)
    else:
        q = np.clip(mean_q + rng.integers(-8, 9, n), 2, 41)
        q = np.where(q < 18, 6, q)
        run = _homopolymer_runlen(bases)
        indel = np.where(run >= 4, np.maximum(10, 45 - 3 * run), 45)
        i = indel
        d = indel.copy()
    c = np.full(n, 10)
    return q.astype(np.uint8), i.astype(np.uint8), d.astype(np.uint8), c.astype(np.uint8)


def _sample_read(rng, hap: np.ndarray, length: int, sub_rate_from_q: bool, mean_q: int, n_rate: float,
                 lowq: bool = False, sub_rate: Optional[float] = None):
    if len(hap) >= length:
        s = int(rng.integers(0, len(hap) - length + 1))
        b = hap[s:s + length].copy()
    else:  # read longer than the haplotype: the overhang is random sequence (soft-clip like)
        b = np.concatenate([hap, _ACGT[rng.integers(0, 4, length - len(hap))]])
    q, i, d, c = _gatk_like_quals(rng, b, mean_q, lowq)
    if sub_rate is not None:
        err = rng.random(length) < sub_rate
    elif sub_rate_from_q:
        err = rng.random(length) < np.power(10.0, -q.astype(np.float64) / 10.0)
    else:
        err = np.zeros(length, bool)
    if err.any():
        idx = np.flatnonzero(err)
        cur = np.searchsorted(_ACGT, b[idx])
        b[idx] = _ACGT[(cur + rng.integers(1, 4, len(idx))) % 4]
    if n_rate > 0:
        b[rng.random(length) < n_rate] = ord("N")
    return np.ascontiguousarray(b), q, i, d, c


# ----------------------------------------------------------------------------------------
def config1_golden(n_regions: int = 400, seed: int = 1001) -> FlatBatch:
    """C1 stand-in for "htc on a 1 Mb slice": H in [2,12] haplotypes of length U[120,350]
    derived from one backbone by 0-3 SNPs and 0-2 indels (1-10 bp); R in [10,80] reads,
    Lr = 150 (10 % clipped to U[60,149]); 1 % N bases."""
    rng = _rng(seed)
    B = _Builder()
    for _ in range(n_regions):
        L0 = int(rng.integers(120, 351))
        backbone = _ACGT[rng.integers(0, 4, L0)]
        H = int(rng.integers(2, 13))
        haps = [np.ascontiguousarray(backbone)]
        for _h in range(H - 1):
            haps.append(_mutate_hap(rng, backbone, int(rng.integers(0, 4)), int(rng.integers(0, 3)), 10))
        R = int(rng.integers(10, 81))
        reads = []
        for _r in range(R):
            Lr = 150 if rng.random() >= 0.10 else int(rng.integers(60, 150))
            src = haps[int(rng.integers(0, H))]
            mean_q = int(rng.choice([20, 30, 37]))
            reads.append(_sample_read(rng, src, Lr, True, mean_q, 0.01))
        B.add_region(reads, haps)
    return B.build("C1-golden-standin", {"seed": seed, "config": 1})


def config2_uniform(n_regions: int = 100, reads_per_region: int = 100, haps_per_region: int = 10, read_len: int = 150,
                    hap_len: int = 300, seed: int = 2002, random_quals: bool = False) -> FlatBatch:
    """C2: 100 regions x 100 reads x 10 haps = 100 000 pairs, 150 bp x 300 bp exactly,
    "uniform quals" = constant q 30, i = d = 45, c = 10 (C2b: q ~ U[6,41]); reads are
    haplotype substrings with 1 % substitutions."""
    rng = _rng(seed)
    B = _Builder()
    for _ in range(n_regions):
        backbone = _ACGT[rng.integers(0, 4, hap_len)]
        haps = [np.ascontiguousarray(backbone)]
        for _h in range(haps_per_region - 1):
            haps.append(_mutate_hap(rng, backbone, int(rng.integers(1, 4)), int(rng.integers(0, 2)), 10, target_len=hap_len))
        reads = []
        for _r in range(reads_per_region):
            src = haps[int(rng.integers(0, haps_per_region))]
            s = int(rng.integers(0, hap_len - read_len + 1))
            b = src[s:s + read_len].copy()
            err = np.flatnonzero(rng.random(read_len) < 0.01)
            if len(err):
                b[err] = _ACGT[(np.searchsorted(_ACGT, b[err]) + rng.integers(1, 4, len(err))) % 4]
            q = rng.integers(6, 42, read_len).astype(np.uint8) if random_quals else np.full(read_len, 30, np.uint8)
            reads.append((np.ascontiguousarray(b), q, np.full(read_len, 45, np.uint8), np.full(read_len, 45, np.uint8),
                          np.full(read_len, 10, np.uint8)))
        B.add_region(reads, haps)
    return B.build("C2b-random-quals" if random_quals else "C2-uniform", {"seed": seed, "config": 2})


def config3_wgs(n_regions: int = 2000, seed: int = 3003, chunk: int = 0) -> FlatBatch:
    """C3: 30x-WGS-shaped stream.  Per region R ~ clip(NegBin(mean 40),10,120),
    H ~ clip(Geom(mean 5),2,16), read length U[100,250] per region (10 % of reads clipped
    shorter), haplotype length U[100,600] per region (haps of a region within 20 bp).
    The full config is 250 000 regions (~5e7 pairs); `n_regions` takes a chunk of it,
    seed = 3003 + chunk index."""
    rng = _rng(seed + chunk)
    B = _Builder()
    for _ in range(n_regions):
        R = int(np.clip(rng.negative_binomial(4, 4.0 / (4.0 + 40.0)), 10, 120))
        H = int(np.clip(rng.geometric(1.0 / 5.0), 2, 16))
        Lr = int(rng.integers(100, 251))
        Lh = int(rng.integers(100, 601))
        backbone = _ACGT[rng.integers(0, 4, Lh)]
        haps = [np.ascontiguousarray(backbone)]
        for _h in range(H - 1):
            haps.append(_mutate_hap(rng, backbone, int(rng.integers(0, 4)), int(rng.integers(0, 3)), 10))
        reads = []
        for _r in range(R):
            L = Lr if rng.random() >= 0.10 else int(rng.integers(max(30, Lr // 3), Lr))
            src = haps[int(rng.integers(0, H))]
            reads.append(_sample_read(rng, src, L, True, int(rng.choice([20, 30, 37])), 0.01))
        B.add_region(reads, haps)
    return B.build("C3-wgs", {"seed": seed + chunk, "config": 3})


def config4_mutect2(n_regions: int = 50, seed: int = 4004) -> FlatBatch:
    """C4: Mutect2-shaped tumor/normal 100x regions: R ~ U[200,2000], H ~ U[8,32],
    Lr = 150, Lh ~ U[200,400]; 2 % of reads come from a rare haplotype.  Full config: 5 000
    regions."""
    rng = _rng(seed)
    B = _Builder()
    for _ in range(n_regions):
        R = int(rng.integers(200, 2001))
        H = int(rng.integers(8, 33))
        Lh = int(rng.integers(200, 401))
        backbone = _ACGT[rng.integers(0, 4, Lh)]
        haps = [np.ascontiguousarray(backbone)]
        for _h in range(H - 1):
            haps.append(_mutate_hap(rng, backbone, int(rng.integers(0, 4)), int(rng.integers(0, 3)), 10))
        rare = H - 1
        reads = []
        for _r in range(R):
            hidx = rare if rng.random() < 0.02 else int(rng.integers(0, max(1, H - 1)))
            reads.append(_sample_read(rng, haps[hidx], 150, True, int(rng.choice([20, 30, 37])), 0.01))
        B.add_region(reads, haps)
    return B.build("C4-mutect2", {"seed": seed, "config": 4})


def config5_underflow(n_regions: int = 200, reads_per_region: int = 10, haps_per_region: int = 10, read_len: int = 250,
                      hap_len: int = 1000, seed: int = 5005) -> FlatBatch:
    """C5: underflow stress: 250 bp reads x 1 kb haplotypes carrying 10-30 indels of 5-50 bp
    relative to the read's source, q ~ U[2,15], ins/del ~ U[10,20], 15 % substitutions, so
    most FP32 sums fall under 1e-28 and the pair is recomputed in double."""
    rng = _rng(seed)
    B = _Builder()
    for _ in range(n_regions):
        source = _ACGT[rng.integers(0, 4, hap_len)]
        haps = []
        for _h in range(haps_per_region):
            haps.append(_mutate_hap(rng, source, 0, int(rng.integers(10, 31)), 50, target_len=hap_len))
            # indels of 5..50: _mutate_hap draws 1..50; bias away from tiny ones is not needed for the stress
        reads = []
        for _r in range(reads_per_region):
            reads.append(_sample_read(rng, source, read_len, False, 0, 0.0, lowq=True, sub_rate=0.15))
        B.add_region(reads, haps)
    return B.build("C5-underflow", {"seed": seed, "config": 5})


def tiny_mixed(seed: int = 7, n_regions: int = 6) -> FlatBatch:
    """Small ragged batch for smoke tests: mixed read/hap lengths, N bases, low quals."""
    rng = _rng(seed)
    B = _Builder()
    for g in range(n_regions):
        Lh = int(rng.integers(20, 120))
        backbone = _ACGT[rng.integers(0, 4, Lh)]
        H = int(rng.integers(1, 5))
        haps = [np.ascontiguousarray(backbone)] + [_mutate_hap(rng, backbone, 2, 1, 5) for _ in range(H - 1)]
        reads = []
        for _r in range(int(rng.integers(1, 9))):
            L = int(rng.integers(5, 100))
            reads.append(_sample_read(rng, haps[int(rng.integers(0, H))], L, True, int(rng.choice([10, 20, 30])), 0.02,
                                      lowq=(g % 3 == 2)))
        B.add_region(reads, haps)
    return B.build("tiny-mixed", {"seed": seed})
