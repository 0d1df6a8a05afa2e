"""Host-side mirror of the reference's PairHMM operator surface.

In the reference path the JVM that /root/reference/src/workers/HTCWorker.cpp:48-113 launches
drives the native library through three calls [upstream GATK VectorLoglessPairHMM / Intel GKL
IntelPairHmm]: ``initNative(readClass, hapClass, use_double, max_threads)``,
``computeLikelihoodsNative(reads[], haps[], double[] out)`` once per active region, and
``doneNative()``.  :class:`PairHMM` keeps those names and argument meanings
(``initialize`` / ``compute_likelihoods`` / ``done``) and adds the batched forms a GPU wants
(many regions per call, flat batches, device-resident batches).  Everything goes through the
C ABI of libfcs_pairhmm.so; nothing here computes.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .batch import FlatBatch, Region


class PairHMMError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libfcs_pairhmm: {_lib.ERROR_NAMES.get(code, code)}: {msg}")
        self.code = code


def _flat_struct(b: FlatBatch) -> _lib.FlatStruct:
    s = _lib.FlatStruct()
    s.read_bases = _lib.as_u8p(b.read_bases); s.read_q = _lib.as_u8p(b.read_q); s.read_i = _lib.as_u8p(b.read_i)
    s.read_d = _lib.as_u8p(b.read_d); s.read_c = _lib.as_u8p(b.read_c)
    s.rd_off = b.rd_off.ctypes.data_as(_lib.i64p); s.rd_len = b.rd_len.ctypes.data_as(_lib.i32p); s.n_reads = b.n_reads
    s.hap_bases = _lib.as_u8p(b.hap_bases)
    s.hp_off = b.hp_off.ctypes.data_as(_lib.i64p); s.hp_len = b.hp_len.ctypes.data_as(_lib.i32p); s.n_haps = b.n_haps
    s.reg_read0 = b.reg_read0.ctypes.data_as(_lib.i32p); s.reg_nreads = b.reg_nreads.ctypes.data_as(_lib.i32p)
    s.reg_hap0 = b.reg_hap0.ctypes.data_as(_lib.i32p); s.reg_nhaps = b.reg_nhaps.ctypes.data_as(_lib.i32p)
    s.reg_out0 = b.reg_out0.ctypes.data_as(_lib.i64p); s.n_regions = b.n_regions
    return s


class RegionArray:
    """An array of ``fcs_phmm_region`` structs over a FlatBatch's memory — the exact argument
    of ``fcs_pairhmm_compute`` (pointer-per-read form, as a JNI shim sees JVM arrays).
    Built once; holds the output arrays the library scatters into."""

    def __init__(self, b: FlatBatch, want_flags: bool = True):
        self.batch = b
        self.out = np.zeros(b.n_pairs, dtype=np.float64)
        self.used = np.zeros(b.n_pairs, dtype=np.uint8) if want_flags else None
        self.reads = (_lib.Read * max(1, b.n_reads))()
        self.haps = (_lib.Hap * max(1, b.n_haps))()
        self.regions = (_lib.RegionStruct * max(1, b.n_regions))()
        planes = [p.ctypes.data for p in (b.read_bases, b.read_q, b.read_i, b.read_d, b.read_c)]
        ra = C.addressof(self.reads)
        # fill through numpy views of the struct arrays (fast for 1e5 reads)
        rview = np.frombuffer(self.reads, dtype=np.dtype([("b", "<u8"), ("q", "<u8"), ("i", "<u8"), ("d", "<u8"), ("c", "<u8"), ("len", "<i4"), ("pad", "<i4")]))
        if b.n_reads:
            off = b.rd_off.astype(np.uint64)
            rview["b"][: b.n_reads] = planes[0] + off
            rview["q"][: b.n_reads] = planes[1] + off
            rview["i"][: b.n_reads] = planes[2] + off
            rview["d"][: b.n_reads] = planes[3] + off
            rview["c"][: b.n_reads] = planes[4] + off
            rview["len"][: b.n_reads] = b.rd_len
        hview = np.frombuffer(self.haps, dtype=np.dtype([("b", "<u8"), ("len", "<i4"), ("pad", "<i4")]))
        if b.n_haps:
            hview["b"][: b.n_haps] = b.hap_bases.ctypes.data + b.hp_off.astype(np.uint64)
            hview["len"][: b.n_haps] = b.hp_len
        gview = np.frombuffer(self.regions, dtype=np.dtype([("reads", "<u8"), ("n_reads", "<i4"), ("p0", "<i4"), ("haps", "<u8"), ("n_haps", "<i4"), ("p1", "<i4"), ("out", "<u8"), ("used", "<u8")]))
        if b.n_regions:
            gview["reads"][: b.n_regions] = ra + b.reg_read0.astype(np.uint64) * C.sizeof(_lib.Read)
            gview["n_reads"][: b.n_regions] = b.reg_nreads
            gview["haps"][: b.n_regions] = C.addressof(self.haps) + b.reg_hap0.astype(np.uint64) * C.sizeof(_lib.Hap)
            gview["n_haps"][: b.n_regions] = b.reg_nhaps
            gview["out"][: b.n_regions] = self.out.ctypes.data + b.reg_out0.astype(np.uint64) * 8
            gview["used"][: b.n_regions] = (self.used.ctypes.data + b.reg_out0.astype(np.uint64)) if want_flags else 0
        self.n = b.n_regions


class ResidentBatch:
    """A batch packed and uploaded once (``fcs_pairhmm_batch_*``): kernel-only runs."""

    def __init__(self, hmm: "PairHMM", b: FlatBatch, device_index: int = 0):
        self._hmm = hmm
        self._lib = hmm._lib
        self.batch = b
        self._h = C.c_void_p()
        fs = _flat_struct(b)
        hmm._check(self._lib.fcs_pairhmm_batch_create(hmm._h, C.byref(fs), device_index, C.byref(self._h)))
        self.pairs = int(self._lib.fcs_pairhmm_batch_pairs(self._h))
        self.cells = int(self._lib.fcs_pairhmm_batch_cells(self._h))
        self.launches = int(self._lib.fcs_pairhmm_batch_launches(self._h))

    def run(self):
        self._hmm._check(self._lib.fcs_pairhmm_batch_run(self._hmm._h, self._h))

    def run_timed(self) -> Tuple[float, float]:
        """Returns (total kernel ms, FP32 main-kernel ms) from CUDA events on the launching stream."""
        t = C.c_float()
        m = C.c_float()
        self._hmm._check(self._lib.fcs_pairhmm_batch_run_timed(self._hmm._h, self._h, C.byref(t), C.byref(m)))
        return float(t.value), float(m.value)

    def sync(self):
        self._hmm._check(self._lib.fcs_pairhmm_batch_sync(self._hmm._h, self._h))

    def download(self, want_raw: bool = False):
        n = self.batch.n_pairs
        out = np.zeros(n, np.float64)
        used = np.zeros(n, np.uint8)
        raw = np.zeros(n, np.float32) if want_raw else None
        self._hmm._check(self._lib.fcs_pairhmm_batch_download(
            self._hmm._h, self._h, out.ctypes.data_as(_lib.f64p), used.ctypes.data_as(_lib.u8p),
            raw.ctypes.data_as(_lib.f32p) if want_raw else None))
        return (out, used, raw) if want_raw else (out, used)

    def close(self):
        if self._h:
            self._lib.fcs_pairhmm_batch_destroy(self._hmm._h, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PairHMM:
    """GKL-style PairHMM object backed by the B200 library.

    initialize(use_double, max_threads)  ~ initNative
    compute_likelihoods(reads, haps)     ~ computeLikelihoodsNative (one region)
    done()                               ~ doneNative
    """

    def __init__(self, use_double: bool = False, max_threads: int = 0, devices: Optional[Sequence[int]] = None,
                 keep_raw_f32: bool = False, slots_per_device: int = 0, max_chunk_cells: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.initialize(use_double, max_threads, devices, keep_raw_f32, slots_per_device, max_chunk_cells)

    # -- lifecycle -------------------------------------------------------------------
    def initialize(self, use_double=False, max_threads=0, devices=None, keep_raw_f32=False, slots_per_device=0, max_chunk_cells=0):
        if self._h:
            self.done()
        cfg = _lib.Config()
        cfg.struct_size = C.sizeof(_lib.Config)
        self._dev_arr = None
        if devices is not None:
            self._dev_arr = (C.c_int32 * len(devices))(*devices)
            cfg.n_devices = len(devices)
            cfg.devices = self._dev_arr
        cfg.use_double = int(bool(use_double))
        cfg.max_threads = int(max_threads)
        cfg.slots_per_device = int(slots_per_device)
        cfg.max_chunk_cells = int(max_chunk_cells)
        cfg.keep_raw_f32 = int(bool(keep_raw_f32))
        self.keep_raw_f32 = bool(keep_raw_f32)
        rc = self._lib.fcs_pairhmm_create(C.byref(cfg), C.byref(self._h))
        if rc != _lib.OK:
            self._h = C.c_void_p()
            raise PairHMMError(rc, (self._lib.fcs_pairhmm_last_error(None) or b"").decode())

    def done(self):
        if self._h:
            self._lib.fcs_pairhmm_destroy(self._h)
            self._h = C.c_void_p()

    close = done

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.done()

    def __del__(self):
        try:
            self.done()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != _lib.OK:
            raise PairHMMError(rc, (self._lib.fcs_pairhmm_last_error(self._h) or b"").decode())

    @property
    def device_count(self) -> int:
        return int(self._lib.fcs_pairhmm_device_count(self._h))

    # -- compute ---------------------------------------------------------------------
    def compute_likelihoods(self, reads: Sequence[Tuple[bytes, bytes, bytes, bytes, bytes]], haps: Sequence[bytes],
                            return_flags: bool = False):
        """One region, GKL argument meaning: returns the reads x haps log10 likelihood matrix."""
        b = FlatBatch.from_regions([Region(list(reads), list(haps))])
        out, used = self.compute_regions(b)
        m = out.reshape(len(reads), len(haps))
        return (m, used.reshape(len(reads), len(haps))) if return_flags else m

    def compute_regions(self, b, region_array: Optional[RegionArray] = None):
        """Many regions through ``fcs_pairhmm_compute`` (pointer-per-read form).  `b` is a
        FlatBatch or a list of Region."""
        if not isinstance(b, FlatBatch):
            b = FlatBatch.from_regions(b)
        ra = region_array or RegionArray(b)
        self._check(self._lib.fcs_pairhmm_compute(self._h, ra.regions, ra.n))
        return ra.out, ra.used

    def compute_flat(self, b: FlatBatch, want_raw: bool = False):
        """Many regions through ``fcs_pairhmm_compute_flat``."""
        fs = _flat_struct(b)
        n = b.n_pairs
        out = np.zeros(n, np.float64)
        used = np.zeros(n, np.uint8)
        raw = np.zeros(n, np.float32) if want_raw else None
        self._check(self._lib.fcs_pairhmm_compute_flat(
            self._h, C.byref(fs), out.ctypes.data_as(_lib.f64p), used.ctypes.data_as(_lib.u8p),
            raw.ctypes.data_as(_lib.f32p) if want_raw else None))
        return (out, used, raw) if want_raw else (out, used)

    def set_finalize(self, enabled: bool = True, log10_global_mismapping_rate: float = -4.5, expected_error_rate: float = 0.02):
        """Fuse GATK's per-read cap (best + mismapping rate) and the poorly-modelled-read test into the device pipeline."""
        fp = _lib.FinalizeParams(int(bool(enabled)), 0, log10_global_mismapping_rate, expected_error_rate)
        self._check(self._lib.fcs_pairhmm_set_finalize(self._h, C.byref(fp)))

    def compute_flat_finalized(self, b: FlatBatch):
        """``fcs_pairhmm_compute_flat_finalized``: (capped log10 matrix, used_fp64, poorly-modelled flag per read)."""
        fs = _flat_struct(b)
        out = np.zeros(b.n_pairs, np.float64)
        used = np.zeros(b.n_pairs, np.uint8)
        poorly = np.zeros(b.n_reads, np.uint8)
        self._check(self._lib.fcs_pairhmm_compute_flat_finalized(self._h, C.byref(fs), out.ctypes.data_as(_lib.f64p), used.ctypes.data_as(_lib.u8p),
                                                                 poorly.ctypes.data_as(_lib.u8p)))
        return out, used, poorly

    def submit(self, region_array: RegionArray) -> int:
        t = C.c_int64()
        self._check(self._lib.fcs_pairhmm_submit(self._h, region_array.regions, region_array.n, C.byref(t)))
        return int(t.value)

    def wait(self, ticket: int):
        self._check(self._lib.fcs_pairhmm_wait(self._h, ticket))

    def resident(self, b: FlatBatch, device_index: int = 0) -> ResidentBatch:
        return ResidentBatch(self, b, device_index)

    def set_capture(self, path: Optional[str]):
        """Append every region scored from now on to `path` (FCSPHMM1 format); None stops."""
        self._check(self._lib.fcs_pairhmm_set_capture(self._h, path.encode() if path else None))

    # -- introspection ------------------------------------------------------------------
    def stats(self) -> dict:
        s = _lib.Stats()
        self._check(self._lib.fcs_pairhmm_get_stats(self._h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in _lib.Stats._fields_}

    def reset_stats(self):
        self._check(self._lib.fcs_pairhmm_reset_stats(self._h))


def plan_check(b: FlatBatch, sm_count: int = 148) -> dict:
    """Run the batcher on the host (no GPU): verifies pair coverage, returns its decisions and timings."""
    lib = _lib.load()
    fs = _flat_struct(b)
    info = _lib.PlanInfo()
    rc = lib.fcs_pairhmm_plan_check(C.byref(fs), sm_count, C.byref(info))
    if rc != _lib.OK:
        raise PairHMMError(rc, (lib.fcs_pairhmm_last_error(None) or b"").decode())
    return {f: getattr(info, f) for f, _ in _lib.PlanInfo._fields_}


def kernel_class(read_len: int, fp64: bool = False) -> Tuple[int, int]:
    lib = _lib.load()
    g = C.c_int32()
    r = C.c_int32()
    rc = lib.fcs_pairhmm_kernel_class(read_len, int(fp64), C.byref(g), C.byref(r))
    if rc != _lib.OK:
        raise PairHMMError(rc, (lib.fcs_pairhmm_last_error(None) or b"").decode())
    return int(g.value), int(r.value)
