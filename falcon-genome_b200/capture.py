"""Testcase capture / replay files.

* FCSPHMM1 binary captures — written by the library itself when capture is on
  (``PairHMM.set_capture`` / env ``FCS_PHMM_CAPTURE``) or by :func:`save_capture`; read back through the
  library's loader (``fcs_pairhmm_capture_load``), so Python and C++ share one parser.
* GKL / GATK text testcases [upstream convention]: one pair per line,
  ``hap read base_quals ins_quals del_quals gcp [expected_log10]`` with quals as ASCII+33.
"""
from __future__ import annotations

import ctypes as C
import struct
from typing import List, Optional, Tuple

import numpy as np

from . import _lib
from .batch import FlatBatch, Region

MAGIC = b"FCSPHMM1"
BLOCK_TAG = 0x4B4C4252


def save_capture(b: FlatBatch, path: str, append: bool = False) -> None:
    """Write a batch in the library's capture format (one block)."""
    with open(path, "ab" if append else "wb") as f:
        if not append or f.tell() == 0:
            f.write(MAGIC)
        f.write(struct.pack("<II", BLOCK_TAG, b.n_regions))
        for g in range(b.n_regions):
            nr, nh = int(b.reg_nreads[g]), int(b.reg_nhaps[g])
            f.write(struct.pack("<II", nr, nh))
            for r in range(int(b.reg_read0[g]), int(b.reg_read0[g]) + nr):
                o, n = int(b.rd_off[r]), int(b.rd_len[r])
                f.write(struct.pack("<I", n))
                for p in (b.read_bases, b.read_q, b.read_i, b.read_d, b.read_c):
                    f.write(p[o:o + n].tobytes())
            for h in range(int(b.reg_hap0[g]), int(b.reg_hap0[g]) + nh):
                o, n = int(b.hp_off[h]), int(b.hp_len[h])
                f.write(struct.pack("<I", n))
                f.write(b.hap_bases[o:o + n].tobytes())


def load_capture(path: str) -> FlatBatch:
    """Read a capture file through libfcs_pairhmm's own loader (host only, no GPU needed)."""
    lib = _lib.load()
    fs = _lib.FlatStruct()
    owner = C.c_void_p()
    rc = lib.fcs_pairhmm_capture_load(path.encode(), C.byref(fs), C.byref(owner))
    if rc != _lib.OK:
        from .pairhmm import PairHMMError

        raise PairHMMError(rc, (lib.fcs_pairhmm_last_error(None) or b"").decode())
    try:
        def arr(ptr, n, dt):
            if n == 0:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True)

        nreads, nhaps, nreg = int(fs.n_reads), int(fs.n_haps), int(fs.n_regions)
        rd_len = arr(fs.rd_len, nreads, np.int32)
        hp_len = arr(fs.hp_len, nhaps, np.int32)
        nrb, nhb = int(rd_len.astype(np.int64).sum()), int(hp_len.astype(np.int64).sum())
        return FlatBatch(
            arr(fs.read_bases, nrb, np.uint8), arr(fs.read_q, nrb, np.uint8), arr(fs.read_i, nrb, np.uint8),
            arr(fs.read_d, nrb, np.uint8), arr(fs.read_c, nrb, np.uint8), arr(fs.rd_off, nreads, np.int64), rd_len,
            arr(fs.hap_bases, nhb, np.uint8), arr(fs.hp_off, nhaps, np.int64), hp_len,
            arr(fs.reg_read0, nreg, np.int32), arr(fs.reg_nreads, nreg, np.int32), arr(fs.reg_hap0, nreg, np.int32),
            arr(fs.reg_nhaps, nreg, np.int32), arr(fs.reg_out0, nreg, np.int64), name=path)
    finally:
        lib.fcs_pairhmm_capture_free(owner)


# ---- GKL text testcases ---------------------------------------------------------------------
def _dec(s: str) -> bytes:
    return bytes(ord(ch) - 33 for ch in s)


def _enc(b: bytes) -> str:
    return "".join(chr(min(v, 93) + 33) for v in b)


def read_gkl_text(path: str) -> Tuple[FlatBatch, np.ndarray]:
    """Parse GKL-style text testcases.  Consecutive lines with the same read (bases and quals) are one
    region (one read x its haplotypes), as GATK's debug dump writes them.  Returns (batch, expected)
    with NaN where a line has no expected value."""
    regions: List[Region] = []
    expected: List[float] = []
    last_read: Optional[tuple] = None
    for ln in open(path):
        ln = ln.strip()
        if not ln or ln.startswith("#"):
            continue
        t = ln.split()
        if len(t) < 6:
            raise ValueError(f"bad testcase line: {ln[:60]}")
        hap, read = t[0].encode(), (t[1].encode(), _dec(t[2]), _dec(t[3]), _dec(t[4]), _dec(t[5]))
        if last_read == read:
            regions[-1].haps.append(hap)
        else:
            regions.append(Region([read], [hap]))
            last_read = read
        expected.append(float(t[6]) if len(t) > 6 else float("nan"))
    return FlatBatch.from_regions(regions), np.asarray(expected, dtype=np.float64)


def write_gkl_text(b: FlatBatch, path: str, results: Optional[np.ndarray] = None) -> None:
    """One line per (read, hap) pair, read-major inside a region."""
    with open(path, "w") as f:
        for g in range(b.n_regions):
            reg = b.region(g)
            o0 = int(b.reg_out0[g])
            for r, (bs, q, i, d, c) in enumerate(reg.reads):
                for h, hap in enumerate(reg.haps):
                    line = f"{hap.decode()} {bs.decode()} {_enc(q)} {_enc(i)} {_enc(d)} {_enc(c)}"
                    if results is not None:
                        line += f" {results[o0 + r * len(reg.haps) + h]:.10f}"
                    f.write(line + "\n")
