"""Client of the fcs-pairhmm-nam daemon and helpers to run the daemon the way the
reference runs its accelerator manager: started in the background before the fan-out, stopped with
SIGALRM afterwards (/root/reference/src/BackgroundExecutor.cpp:13-84)."""
from __future__ import annotations

import ctypes as C
import os
import signal
import subprocess
import time
from typing import Optional

import numpy as np

from . import _lib
from .batch import FlatBatch
from .pairhmm import PairHMMError, _flat_struct

NAM_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fcs-pairhmm-nam")


class RemotePairHMM:
    """compute_flat() executed by the daemon that owns the GPUs."""

    def __init__(self, socket_path: str):
        self._lib = _lib.load_client()
        self._h = C.c_void_p()
        rc = self._lib.fcs_pairhmm_remote_open(socket_path.encode(), C.byref(self._h))
        if rc != _lib.OK:
            raise PairHMMError(rc, (self._lib.fcs_pairhmm_remote_last_error(None) or b"").decode())

    def compute_flat(self, b: FlatBatch):
        fs = _flat_struct(b)
        out = np.zeros(b.n_pairs, np.float64)
        used = np.zeros(b.n_pairs, np.uint8)
        rc = self._lib.fcs_pairhmm_remote_compute_flat(self._h, C.byref(fs), out.ctypes.data_as(_lib.f64p), used.ctypes.data_as(_lib.u8p))
        if rc != _lib.OK:
            raise PairHMMError(rc, (self._lib.fcs_pairhmm_remote_last_error(self._h) or b"").decode())
        return out, used

    def compute_in_segment(self, b: FlatBatch):
        """The zero-copy path (``fcs_pairhmm_remote_reserve`` / ``compute_reserved``): the batch is written straight
        into the connection's shared segment through the views the client hands out (here with numpy; a JNI shim would
        GetByteArrayRegion into them), the results are read in place.  Needs a dense FlatBatch (regions in order)."""
        v = _lib.RemoteViews()
        nr, nh, ng = b.n_reads, b.n_haps, b.n_regions
        rbytes, hbytes = int(b.rd_len.astype(np.int64).sum()), int(b.hp_len.astype(np.int64).sum())
        rc = self._lib.fcs_pairhmm_remote_reserve(self._h, ng, nr, nh, rbytes, hbytes, b.n_pairs, C.byref(v))
        if rc != _lib.OK:
            raise PairHMMError(rc, (self._lib.fcs_pairhmm_remote_last_error(self._h) or b"").decode())

        def view(ptr, n, dtype):
            return np.ctypeslib.as_array(ptr, shape=(max(n, 1),))[:n] if n else np.zeros(0, dtype)

        # densely re-indexed planes (reads / haplotypes in region order), written in place
        rd_off = (np.cumsum(b.rd_len.astype(np.int64)) - b.rd_len).astype(np.int64)
        hp_off = (np.cumsum(b.hp_len.astype(np.int64)) - b.hp_len).astype(np.int64)
        for name, src in (("read_bases", b.read_bases), ("read_q", b.read_q), ("read_i", b.read_i), ("read_d", b.read_d), ("read_c", b.read_c)):
            dst = view(getattr(v, name), rbytes, np.uint8)
            for k in range(nr):
                dst[rd_off[k]:rd_off[k] + b.rd_len[k]] = src[b.rd_off[k]:b.rd_off[k] + b.rd_len[k]]
        dst = view(v.hap_bases, hbytes, np.uint8)
        for k in range(nh):
            dst[hp_off[k]:hp_off[k] + b.hp_len[k]] = b.hap_bases[b.hp_off[k]:b.hp_off[k] + b.hp_len[k]]
        view(v.rd_off, nr, np.int64)[:] = rd_off
        view(v.rd_len, nr, np.int32)[:] = b.rd_len
        view(v.hp_off, nh, np.int64)[:] = hp_off
        view(v.hp_len, nh, np.int32)[:] = b.hp_len
        view(v.reg_read0, ng, np.int32)[:] = b.reg_read0
        view(v.reg_nreads, ng, np.int32)[:] = b.reg_nreads
        view(v.reg_hap0, ng, np.int32)[:] = b.reg_hap0
        view(v.reg_nhaps, ng, np.int32)[:] = b.reg_nhaps
        rc = self._lib.fcs_pairhmm_remote_compute_reserved(self._h)
        if rc != _lib.OK:
            raise PairHMMError(rc, (self._lib.fcs_pairhmm_remote_last_error(self._h) or b"").decode())
        n = b.n_pairs
        return view(v.out_log10, n, np.float64).copy(), view(v.out_used_fp64, n, np.uint8).copy()

    @property
    def uses_shm(self) -> bool:
        """True while requests go through the shared-memory segment, False on the byte-stream protocol."""
        return bool(self._lib.fcs_pairhmm_remote_uses_shm(self._h))

    def close(self):
        if self._h:
            self._lib.fcs_pairhmm_remote_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class NamDaemon:
    """Background-executor style lifecycle: start, wait for "ready", stop with SIGALRM."""

    def __init__(self, socket_path: str, devices: int = 0, timeout_s: float = 60.0):
        self.socket_path = socket_path
        cmd = [NAM_PATH, socket_path] + (["--devices", str(devices)] if devices else [])
        self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        line = self.proc.stdout.readline()
        t0 = time.time()
        while "ready" not in line:
            if self.proc.poll() is not None or time.time() - t0 > timeout_s:
                err = self.proc.stderr.read()
                raise RuntimeError(f"fcs-pairhmm-nam did not start (exit {self.proc.poll()}): {err.strip()}")
            line = self.proc.stdout.readline()

    def stop(self) -> int:
        if self.proc.poll() is None:
            self.proc.send_signal(signal.SIGALRM)
        try:
            return self.proc.wait(timeout=30)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            return self.proc.wait()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.stop()
