"""ctypes wrapper of oracle/liboracle_pairhmm.so — the CPU checker.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by falcon-genome_b200/.
PARITY UNPINNED: see the header of pairhmm_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle_pairhmm.so")

u8p = C.POINTER(C.c_uint8)
_lib = None


def build():
    env = dict(os.environ)
    env.pop("CC", None)
    r = subprocess.run(["make", "-C", _HERE], env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the oracle failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    pair_args = [u8p, u8p, u8p, u8p, u8p, C.c_int, u8p, C.c_int]
    lib.phmm_oracle_sum_double.restype = C.c_double
    lib.phmm_oracle_sum_double.argtypes = pair_args
    lib.phmm_oracle_sum_float.restype = C.c_float
    lib.phmm_oracle_sum_float.argtypes = pair_args
    lib.phmm_oracle_log10_double.restype = C.c_double
    lib.phmm_oracle_log10_double.argtypes = pair_args
    lib.phmm_oracle_bruteforce_log10.restype = C.c_double
    lib.phmm_oracle_bruteforce_log10.argtypes = pair_args
    lib.phmm_oracle_float_sum_to_log10.restype = C.c_double
    lib.phmm_oracle_float_sum_to_log10.argtypes = [C.c_float]
    lib.phmm_oracle_pair.restype = C.c_double
    lib.phmm_oracle_pair.argtypes = pair_args + [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float)]
    for n, t in (("phmm_oracle_ph2pr_d", C.c_double), ("phmm_oracle_ph2pr_f", C.c_float)):
        getattr(lib, n).restype = t
        getattr(lib, n).argtypes = [C.c_int]
    for n, t in (("phmm_oracle_mm_d", C.c_double), ("phmm_oracle_mm_f", C.c_float)):
        getattr(lib, n).restype = t
        getattr(lib, n).argtypes = [C.c_int, C.c_int]
    i64p, i32p = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    lib.phmm_cpu_batch.restype = C.c_int64
    lib.phmm_cpu_batch.argtypes = [u8p, u8p, u8p, u8p, u8p, i64p, i32p, u8p, i64p, i32p, i32p, i32p, i32p, i32p, i64p,
                                   C.c_int, C.POINTER(C.c_double), u8p, C.POINTER(C.c_float), C.c_int, C.c_int]
    lib.phmm_cpu_max_threads.restype = C.c_int
    lib.phmm_variant_batch.restype = C.c_int64
    lib.phmm_variant_batch.argtypes = [C.c_int, u8p, u8p, u8p, u8p, u8p, i64p, i32p, u8p, i64p, i32p, i32p, i32p, i32p, i32p, i64p,
                                       C.c_int, C.POINTER(C.c_double), u8p, C.POINTER(C.c_float), C.c_int]
    lib.phmm_variant_ph2pr_diffs.restype = C.c_int
    lib.phmm_variant_ph2pr_powf.restype = C.c_float
    lib.phmm_variant_ph2pr_powf.argtypes = [C.c_int]
    lib.phmm_double_batch.restype = None
    lib.phmm_double_batch.argtypes = [u8p, u8p, u8p, u8p, u8p, i64p, i32p, u8p, i64p, i32p, i32p, i32p, i32p, i32p, i64p,
                                      C.c_int, C.POINTER(C.c_double), C.c_int]
    lib.phmm_oracle_init()
    _lib = lib
    return lib


def _p(a):
    return a.ctypes.data_as(u8p)


def _arr(x):
    return np.frombuffer(bytes(x), dtype=np.uint8).copy() if not isinstance(x, np.ndarray) else np.ascontiguousarray(x, dtype=np.uint8)


def pair(read, hap, force_double=False):
    """read = (bases, q, i, d, c); returns (log10L, used_double, raw_float_sum) per SURVEY A.4."""
    lib = load()
    b, q, i, d, c = (_arr(x) for x in read)
    h = _arr(hap)
    ud = C.c_int(0)
    rf = C.c_float(0)
    v = lib.phmm_oracle_pair(_p(b), _p(q), _p(i), _p(d), _p(c), len(b), _p(h), len(h), int(force_double), C.byref(ud), C.byref(rf))
    return float(v), int(ud.value), float(rf.value)


def log10_double(read, hap):
    lib = load()
    b, q, i, d, c = (_arr(x) for x in read)
    h = _arr(hap)
    return float(lib.phmm_oracle_log10_double(_p(b), _p(q), _p(i), _p(d), _p(c), len(b), _p(h), len(h)))


def sum_float(read, hap):
    lib = load()
    b, q, i, d, c = (_arr(x) for x in read)
    h = _arr(hap)
    return np.float32(lib.phmm_oracle_sum_float(_p(b), _p(q), _p(i), _p(d), _p(c), len(b), _p(h), len(h)))


def bruteforce_log10(read, hap):
    lib = load()
    b, q, i, d, c = (_arr(x) for x in read)
    h = _arr(hap)
    return float(lib.phmm_oracle_bruteforce_log10(_p(b), _p(q), _p(i), _p(d), _p(c), len(b), _p(h), len(h)))


def batch_scalar(b, force_double=False):
    """Scalar oracle over a FlatBatch: (out log10, used_double, raw float sums, log10 of the double path)."""
    lib = load()
    n = b.n_pairs
    out = np.zeros(n, np.float64)
    used = np.zeros(n, np.uint8)
    raw = np.zeros(n, np.float32)
    dbl = np.zeros(n, np.float64)
    for g in range(b.n_regions):
        nh = int(b.reg_nhaps[g])
        for r in range(int(b.reg_nreads[g])):
            ri = int(b.reg_read0[g]) + r
            ro, rl = int(b.rd_off[ri]), int(b.rd_len[ri])
            planes = [p[ro:ro + rl] for p in (b.read_bases, b.read_q, b.read_i, b.read_d, b.read_c)]
            planes = [np.ascontiguousarray(p) for p in planes]
            for h in range(nh):
                hi = int(b.reg_hap0[g]) + h
                ho, hl = int(b.hp_off[hi]), int(b.hp_len[hi])
                hp = np.ascontiguousarray(b.hap_bases[ho:ho + hl])
                ud = C.c_int(0)
                rf = C.c_float(0)
                o = int(b.reg_out0[g]) + r * nh + h
                out[o] = lib.phmm_oracle_pair(_p(planes[0]), _p(planes[1]), _p(planes[2]), _p(planes[3]), _p(planes[4]), rl,
                                              _p(hp), hl, int(force_double), C.byref(ud), C.byref(rf))
                used[o] = ud.value
                raw[o] = rf.value
                dbl[o] = out[o] if ud.value else lib.phmm_oracle_log10_double(
                    _p(planes[0]), _p(planes[1]), _p(planes[2]), _p(planes[3]), _p(planes[4]), rl, _p(hp), hl)
    return out, used, raw, dbl


def batch_simd(b, nthreads=0, ftz=False):
    """The AVX/OpenMP baseline over a FlatBatch: (out, used_double, raw float sums, n_double)."""
    lib = load()
    n = b.n_pairs
    out = np.zeros(n, np.float64)
    used = np.zeros(n, np.uint8)
    raw = np.zeros(n, np.float32)
    i64p, i32p = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    nd = lib.phmm_cpu_batch(
        _p(b.read_bases), _p(b.read_q), _p(b.read_i), _p(b.read_d), _p(b.read_c), b.rd_off.ctypes.data_as(i64p),
        b.rd_len.ctypes.data_as(i32p), _p(b.hap_bases), b.hp_off.ctypes.data_as(i64p), b.hp_len.ctypes.data_as(i32p),
        b.reg_read0.ctypes.data_as(i32p), b.reg_nreads.ctypes.data_as(i32p), b.reg_hap0.ctypes.data_as(i32p),
        b.reg_nhaps.ctypes.data_as(i32p), b.reg_out0.ctypes.data_as(i64p), b.n_regions,
        out.ctypes.data_as(C.POINTER(C.c_double)), _p(used), raw.ctypes.data_as(C.POINTER(C.c_float)), int(nthreads), int(bool(ftz)))
    return out, used, raw, int(nd)


# arithmetic variants of the float path (oracle/pairhmm_variants.c): what a real GKL binary may compute differently
VAR_NOFMA, VAR_FTZ, VAR_POWF, VAR_LOG10F, VAR_SPLITSUM = 1, 2, 4, 8, 16
VAR_GKL_STRICT_AVX = VAR_NOFMA | VAR_FTZ | VAR_POWF | VAR_LOG10F | VAR_SPLITSUM  # GKL's AVX build, strictest reading
VAR_GKL_STRICT_AVX512 = VAR_FTZ | VAR_POWF | VAR_LOG10F | VAR_SPLITSUM             # FMA-capable build


def batch_variant(b, flags, nthreads=0):
    """Scalar float-first / double-fallback scoring of a FlatBatch under arithmetic variant `flags`
    (0 = the pinned contract, bit-identical to batch_scalar / batch_simd): (out, used_double, raw float sums, n_double)."""
    lib = load()
    n = b.n_pairs
    out = np.zeros(n, np.float64)
    used = np.zeros(n, np.uint8)
    raw = np.zeros(n, np.float32)
    i64p, i32p = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    nd = lib.phmm_variant_batch(
        int(flags), _p(b.read_bases), _p(b.read_q), _p(b.read_i), _p(b.read_d), _p(b.read_c), b.rd_off.ctypes.data_as(i64p),
        b.rd_len.ctypes.data_as(i32p), _p(b.hap_bases), b.hp_off.ctypes.data_as(i64p), b.hp_len.ctypes.data_as(i32p),
        b.reg_read0.ctypes.data_as(i32p), b.reg_nreads.ctypes.data_as(i32p), b.reg_hap0.ctypes.data_as(i32p),
        b.reg_nhaps.ctypes.data_as(i32p), b.reg_out0.ctypes.data_as(i64p), b.n_regions,
        out.ctypes.data_as(C.POINTER(C.c_double)), _p(used), raw.ctypes.data_as(C.POINTER(C.c_float)), int(nthreads))
    return out, used, raw, int(nd)


def batch_double(b, nthreads=0):
    """Every pair of a FlatBatch in double precision (scalar C, OpenMP over reads): log10 L per pair."""
    lib = load()
    out = np.zeros(b.n_pairs, np.float64)
    i64p, i32p = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    lib.phmm_double_batch(
        _p(b.read_bases), _p(b.read_q), _p(b.read_i), _p(b.read_d), _p(b.read_c), b.rd_off.ctypes.data_as(i64p),
        b.rd_len.ctypes.data_as(i32p), _p(b.hap_bases), b.hp_off.ctypes.data_as(i64p), b.hp_len.ctypes.data_as(i32p),
        b.reg_read0.ctypes.data_as(i32p), b.reg_nreads.ctypes.data_as(i32p), b.reg_hap0.ctypes.data_as(i32p),
        b.reg_nhaps.ctypes.data_as(i32p), b.reg_out0.ctypes.data_as(i64p), b.n_regions, out.ctypes.data_as(C.POINTER(C.c_double)), int(nthreads))
    return out


def max_threads():
    return int(load().phmm_cpu_max_threads())
