"""GATK-side steps either side of the PairHMM call — thin binding over
``fcs_pairhmm_prepare_read`` / ``fcs_pairhmm_finalize_region`` (host-only C functions of the library)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib

PCR_NONE, PCR_HOSTILE, PCR_AGGRESSIVE, PCR_CONSERVATIVE = 0, 1, 2, 3


def prepare_read(bases: bytes, raw_quals: bytes, mapq: int = -1, bam_ins: Optional[bytes] = None, bam_del: Optional[bytes] = None,
                 base_q_threshold: int = 18, min_usable_q: int = 6, default_indel_q: int = 45, gcp: int = 10,
                 pcr_model: int = PCR_CONSERVATIVE) -> Tuple[bytes, bytes, bytes, bytes, bytes]:
    """raw read -> (bases, base_q, ins_q, del_q, gcp) as the kernel takes them."""
    lib = _lib.load()
    n = len(bases)
    b = np.frombuffer(bytes(bases), np.uint8)
    q = np.frombuffer(bytes(raw_quals), np.uint8)
    oi = [np.zeros(n, np.uint8) for _ in range(4)]
    pp = _lib.PrepParams(base_q_threshold, min_usable_q, default_indel_q, gcp, pcr_model)
    bi = np.frombuffer(bytes(bam_ins), np.uint8) if bam_ins is not None else None
    bd = np.frombuffer(bytes(bam_del), np.uint8) if bam_del is not None else None
    rc = lib.fcs_pairhmm_prepare_read(_lib.as_u8p(b), _lib.as_u8p(q), n, mapq, _lib.as_u8p(bi) if bi is not None else None,
                                      _lib.as_u8p(bd) if bd is not None else None, C.byref(pp), *[_lib.as_u8p(x) for x in oi])
    if rc != _lib.OK:
        raise RuntimeError((lib.fcs_pairhmm_last_error(None) or b"").decode())
    return (bytes(bases), oi[0].tobytes(), oi[1].tobytes(), oi[2].tobytes(), oi[3].tobytes())


def finalize_region(log10: np.ndarray, read_len, log10_global_mismapping_rate: float = -4.5, expected_error_rate: float = 0.02):
    """Caps every read's row at best + mismapping rate (in place on a copy) and returns (matrix, poorly_modelled flags)."""
    lib = _lib.load()
    m = np.ascontiguousarray(log10, dtype=np.float64).copy()
    nr, nh = m.shape
    rl = np.ascontiguousarray(read_len, dtype=np.int32)
    flags = np.zeros(nr, np.uint8)
    rc = lib.fcs_pairhmm_finalize_region(m.ctypes.data_as(_lib.f64p), nr, nh, rl.ctypes.data_as(_lib.i32p), log10_global_mismapping_rate,
                                         expected_error_rate, _lib.as_u8p(flags))
    if rc != _lib.OK:
        raise RuntimeError((lib.fcs_pairhmm_last_error(None) or b"").decode())
    return m, flags
