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
