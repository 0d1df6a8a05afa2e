"""Flat (structure-of-arrays) PairHMM batches — the host-side mirror of
``fcs_phmm_flat_batch`` in include/fcs_pairhmm.h.

A batch is a list of active regions; a region is R reads x H haplotypes, exactly the
unit GATK hands to ``computeLikelihoodsNative`` ([upstream]).  Reads carry
five byte planes (bases, base/ins/del/gcp quals), haplotypes one.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np


@dataclass
class Region:
    """One active region in array-of-objects form (what the GKL-style call takes)."""

    reads: List[Tuple[bytes, bytes, bytes, bytes, bytes]]  # (bases, q, i, d, c), equal lengths
    haps: List[bytes]


@dataclass
class FlatBatch:
    read_bases: np.ndarray  # uint8
    read_q: np.ndarray
    read_i: np.ndarray
    read_d: np.ndarray
    read_c: np.ndarray
    rd_off: np.ndarray  # int64 [n_reads]
    rd_len: np.ndarray  # int32 [n_reads]
    hap_bases: np.ndarray  # uint8
    hp_off: np.ndarray  # int64 [n_haps]
    hp_len: np.ndarray  # int32 [n_haps]
    reg_read0: np.ndarray  # int32 [n_regions]
    reg_nreads: np.ndarray
    reg_hap0: np.ndarray
    reg_nhaps: np.ndarray
    reg_out0: np.ndarray  # int64 [n_regions]
    name: str = ""
    meta: dict = field(default_factory=dict)

    # ---- sizes -----------------------------------------------------------------
    @property
    def n_regions(self) -> int:
        return int(self.reg_read0.shape[0])

    @property
    def n_reads(self) -> int:
        return int(self.rd_len.shape[0])

    @property
    def n_haps(self) -> int:
        return int(self.hp_len.shape[0])

    @property
    def n_pairs(self) -> int:
        return int((self.reg_nreads.astype(np.int64) * self.reg_nhaps.astype(np.int64)).sum())

    def region_cells(self) -> np.ndarray:
        """DP cells per region: (sum of read lengths) x (sum of haplotype lengths)."""
        rl = np.concatenate([[0], np.cumsum(self.rd_len.astype(np.int64))])
        hl = np.concatenate([[0], np.cumsum(self.hp_len.astype(np.int64))])
        r0 = self.reg_read0.astype(np.int64)
        h0 = self.reg_hap0.astype(np.int64)
        sr = rl[r0 + self.reg_nreads] - rl[r0]
        sh = hl[h0 + self.reg_nhaps] - hl[h0]
        return sr * sh

    @property
    def cells(self) -> int:
        return int(self.region_cells().sum())

    def input_bytes(self) -> int:
        """Algorithmic input bytes: 5 per read base, 1 per haplotype base."""
        return int(5 * self.rd_len.astype(np.int64).sum() + self.hp_len.astype(np.int64).sum())

    # ---- construction ----------------------------------------------------------
    @staticmethod
    def from_regions(regions: Sequence[Region], name: str = "") -> "FlatBatch":
        rb, rq, ri, rdq, rc, rlen = [], [], [], [], [], []
        hb, hlen = [], []
        reg_read0, reg_nreads, reg_hap0, reg_nhaps, reg_out0 = [], [], [], [], []
        nread = nhap = 0
        out0 = 0
        for reg in regions:
            reg_read0.append(nread)
            reg_nreads.append(len(reg.reads))
            reg_hap0.append(nhap)
            reg_nhaps.append(len(reg.haps))
            reg_out0.append(out0)
            out0 += len(reg.reads) * len(reg.haps)
            for (b, q, i, d, c) in reg.reads:
                n = len(b)
                if not (len(q) == len(i) == len(d) == len(c) == n):
                    raise ValueError("read planes differ in length")
                rb.append(np.frombuffer(bytes(b), dtype=np.uint8))
                rq.append(np.frombuffer(bytes(q), dtype=np.uint8))
                ri.append(np.frombuffer(bytes(i), dtype=np.uint8))
                rdq.append(np.frombuffer(bytes(d), dtype=np.uint8))
                rc.append(np.frombuffer(bytes(c), dtype=np.uint8))
                rlen.append(n)
                nread += 1
            for h in reg.haps:
                hb.append(np.frombuffer(bytes(h), dtype=np.uint8))
                hlen.append(len(h))
                nhap += 1

        def cat(xs):
            return np.ascontiguousarray(np.concatenate(xs)) if xs else np.zeros(0, dtype=np.uint8)

        rlen_a = np.asarray(rlen, dtype=np.int32)
        hlen_a = np.asarray(hlen, dtype=np.int32)
        rd_off = np.concatenate([[0], np.cumsum(rlen_a.astype(np.int64))])[:-1].astype(np.int64) if len(rlen) else np.zeros(0, np.int64)
        hp_off = np.concatenate([[0], np.cumsum(hlen_a.astype(np.int64))])[:-1].astype(np.int64) if len(hlen) else np.zeros(0, np.int64)
        return FlatBatch(
            cat(rb), cat(rq), cat(ri), cat(rdq), cat(rc), rd_off, rlen_a, cat(hb), hp_off, hlen_a,
            np.asarray(reg_read0, dtype=np.int32), np.asarray(reg_nreads, dtype=np.int32),
            np.asarray(reg_hap0, dtype=np.int32), np.asarray(reg_nhaps, dtype=np.int32),
            np.asarray(reg_out0, dtype=np.int64), name=name,
        )

    def region(self, g: int) -> Region:
        reads = []
        for r in range(int(self.reg_read0[g]), int(self.reg_read0[g] + self.reg_nreads[g])):
            o, n = int(self.rd_off[r]), int(self.rd_len[r])
            reads.append(tuple(bytes(p[o:o + n]) for p in (self.read_bases, self.read_q, self.read_i, self.read_d, self.read_c)))
        haps = []
        for h in range(int(self.reg_hap0[g]), int(self.reg_hap0[g] + self.reg_nhaps[g])):
            o, n = int(self.hp_off[h]), int(self.hp_len[h])
            haps.append(bytes(self.hap_bases[o:o + n]))
        return Region(reads, haps)

    def select(self, region_ids: Sequence[int], name: str = "") -> "FlatBatch":
        """Sub-batch holding the given regions (data planes are shared, tables rebuilt).

        reg_out0 is recomputed so the sub-batch's output array is dense."""
        ids = np.asarray(region_ids, dtype=np.int64)
        nr = self.reg_nreads[ids]
        nh = self.reg_nhaps[ids]
        out0 = np.concatenate([[0], np.cumsum(nr.astype(np.int64) * nh.astype(np.int64))])[:-1].astype(np.int64) if len(ids) else np.zeros(0, np.int64)
        return FlatBatch(
            self.read_bases, self.read_q, self.read_i, self.read_d, self.read_c, self.rd_off, self.rd_len,
            self.hap_bases, self.hp_off, self.hp_len,
            np.ascontiguousarray(self.reg_read0[ids]), np.ascontiguousarray(nr), np.ascontiguousarray(self.reg_hap0[ids]),
            np.ascontiguousarray(nh), out0, name=name or self.name, meta=dict(self.meta),
        )


def partition_regions(cells: np.ndarray, world: int) -> List[np.ndarray]:
    """Longest-processing-time-first split of regions over `world` ranks by DP cells —
    the same rule the library's in-process multi-GPU dispatcher uses (phmm_engine.cu
    Engine::compute; SURVEY.md §8(e): independent regions, no exchange).  Deterministic:
    ties go to the lower region index, then to the lower rank."""
    cells = np.asarray(cells, dtype=np.int64)
    order = sorted(range(len(cells)), key=lambda g: (-int(cells[g]), g))
    load = [0] * world
    parts: List[List[int]] = [[] for _ in range(world)]
    for g in order:
        best = min(range(world), key=lambda d: (load[d], d))
        parts[best].append(g)
        load[best] += int(cells[g])
    return [np.asarray(sorted(p), dtype=np.int64) for p in parts]
