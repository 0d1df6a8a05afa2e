"""ctypes binding of libfcs_pairhmm.so (the C ABI in include/fcs_pairhmm.h).

The library is built in-tree (falcon-genome_b200/libfcs_pairhmm.so) by
``python -m falcon_genome_b200.build`` / ``__graft_entry__.build()``.  There is no Python or
CPU implementation behind this module: if the shared library is missing, loading fails
loudly; if no B200 is present, ``fcs_pairhmm_create`` fails with FCS_PHMM_ENODEV.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FCS_PHMM_LIB") or os.path.join(_HERE, "libfcs_pairhmm.so")  # env override: developer A/B builds

OK, EINVAL, ENODEV, ECUDA, ENOMEM, EUNSUPPORTED, ETICKET = 0, -1, -2, -3, -4, -5, -6
ERROR_NAMES = {EINVAL: "EINVAL", ENODEV: "ENODEV", ECUDA: "ECUDA", ENOMEM: "ENOMEM", EUNSUPPORTED: "EUNSUPPORTED", ETICKET: "ETICKET"}

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
f64p = C.POINTER(C.c_double)
f32p = C.POINTER(C.c_float)


class Read(C.Structure):
    _fields_ = [("bases", u8p), ("base_q", u8p), ("ins_q", u8p), ("del_q", u8p), ("gcp", u8p), ("len", C.c_int32)]


class Hap(C.Structure):
    _fields_ = [("bases", u8p), ("len", C.c_int32)]


class RegionStruct(C.Structure):
    _fields_ = [("reads", C.POINTER(Read)), ("n_reads", C.c_int32), ("haps", C.POINTER(Hap)), ("n_haps", C.c_int32),
                ("out_log10", f64p), ("out_used_fp64", u8p)]


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n_devices", C.c_int32), ("devices", i32p), ("use_double", C.c_int32),
                ("max_threads", C.c_int32), ("slots_per_device", C.c_int32), ("max_chunk_cells", C.c_int64),
                ("keep_raw_f32", C.c_int32), ("reserved", C.c_int32)]


class PrepParams(C.Structure):
    _fields_ = [("base_q_threshold", C.c_int32), ("min_usable_q", C.c_int32), ("default_indel_q", C.c_int32), ("gcp", C.c_int32),
                ("pcr_model", C.c_int32)]


class FinalizeParams(C.Structure):
    _fields_ = [("enabled", C.c_int32), ("reserved", C.c_int32), ("log10_global_mismapping_rate", C.c_double),
                ("expected_error_rate_per_base", C.c_double)]


class PlanInfo(C.Structure):
    _fields_ = [("n_pairs", C.c_int64), ("n_tasks", C.c_int64), ("n_generic_pairs", C.c_int64), ("in_bytes", C.c_int64),
                ("max_smem_bytes", C.c_int64), ("n_launches_f32", C.c_int32), ("n_launches_f64", C.c_int32), ("n_sym", C.c_int32),
                ("latency_mode", C.c_int32), ("geometric_efficiency", C.c_double), ("plan_ms", C.c_double), ("pack_ms", C.c_double),
                ("n_tasks_general", C.c_int64), ("n_tasks_uniform_gcp", C.c_int64), ("n_tasks_all_uniform", C.c_int64), ("n_tasks_hap_pairs", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("pairs", C.c_uint64), ("cells", C.c_uint64), ("fp64_pairs", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("chunks", C.c_uint64), ("kernel_ms", C.c_double),
                ("main_kernel_ms", C.c_double), ("host_plan_ms", C.c_double), ("host_pack_ms", C.c_double),
                ("host_wait_ms", C.c_double), ("host_scatter_ms", C.c_double)]


class FlatStruct(C.Structure):
    _fields_ = [("read_bases", u8p), ("read_q", u8p), ("read_i", u8p), ("read_d", u8p), ("read_c", u8p),
                ("rd_off", i64p), ("rd_len", i32p), ("n_reads", C.c_int64),
                ("hap_bases", u8p), ("hp_off", i64p), ("hp_len", i32p), ("n_haps", C.c_int64),
                ("reg_read0", i32p), ("reg_nreads", i32p), ("reg_hap0", i32p), ("reg_nhaps", i32p), ("reg_out0", i64p),
                ("n_regions", C.c_int64)]


# every symbol include/fcs_pairhmm.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "fcs_pairhmm_abi_version": (C.c_int, []),
    "fcs_pairhmm_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "fcs_pairhmm_destroy": (None, [C.c_void_p]),
    "fcs_pairhmm_last_error": (C.c_char_p, [C.c_void_p]),
    "fcs_pairhmm_device_count": (C.c_int, [C.c_void_p]),
    "fcs_pairhmm_compute": (C.c_int, [C.c_void_p, C.POINTER(RegionStruct), C.c_int32]),
    "fcs_pairhmm_compute_flat": (C.c_int, [C.c_void_p, C.POINTER(FlatStruct), f64p, u8p, f32p]),
    "fcs_pairhmm_submit": (C.c_int, [C.c_void_p, C.POINTER(RegionStruct), C.c_int32, C.POINTER(C.c_int64)]),
    "fcs_pairhmm_wait": (C.c_int, [C.c_void_p, C.c_int64]),
    "fcs_pairhmm_batch_create": (C.c_int, [C.c_void_p, C.POINTER(FlatStruct), C.c_int32, C.POINTER(C.c_void_p)]),
    "fcs_pairhmm_batch_run": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fcs_pairhmm_batch_run_timed": (C.c_int, [C.c_void_p, C.c_void_p, f32p, f32p]),
    "fcs_pairhmm_batch_sync": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fcs_pairhmm_batch_download": (C.c_int, [C.c_void_p, C.c_void_p, f64p, u8p, f32p]),
    "fcs_pairhmm_batch_pairs": (C.c_int64, [C.c_void_p]),
    "fcs_pairhmm_batch_cells": (C.c_int64, [C.c_void_p]),
    "fcs_pairhmm_batch_launches": (C.c_int32, [C.c_void_p]),
    "fcs_pairhmm_batch_destroy": (None, [C.c_void_p, C.c_void_p]),
    "fcs_pairhmm_set_capture": (C.c_int, [C.c_void_p, C.c_char_p]),
    "fcs_pairhmm_capture_load": (C.c_int, [C.c_char_p, C.POINTER(FlatStruct), C.POINTER(C.c_void_p)]),
    "fcs_pairhmm_capture_parse": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(FlatStruct), C.POINTER(C.c_void_p)]),
    "fcs_pairhmm_capture_free": (None, [C.c_void_p]),
    "fcs_pairhmm_prepare_read": (C.c_int, [u8p, u8p, C.c_int32, C.c_int32, u8p, u8p, C.POINTER(PrepParams), u8p, u8p, u8p, u8p]),
    "fcs_pairhmm_finalize_region": (C.c_int, [f64p, C.c_int32, C.c_int32, i32p, C.c_double, C.c_double, u8p]),
    "fcs_pairhmm_set_finalize": (C.c_int, [C.c_void_p, C.POINTER(FinalizeParams)]),
    "fcs_pairhmm_compute_flat_finalized": (C.c_int, [C.c_void_p, C.POINTER(FlatStruct), f64p, u8p, u8p]),
    "fcs_pairhmm_plan_check": (C.c_int, [C.POINTER(FlatStruct), C.c_int32, C.POINTER(PlanInfo)]),
    "fcs_pairhmm_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "fcs_pairhmm_reset_stats": (C.c_int, [C.c_void_p]),
    "fcs_pairhmm_lut_ph2pr_f32": (C.c_float, [C.c_int]),
    "fcs_pairhmm_lut_ph2pr_f64": (C.c_double, [C.c_int]),
    "fcs_pairhmm_lut_mm_f32": (C.c_float, [C.c_int, C.c_int]),
    "fcs_pairhmm_lut_mm_f64": (C.c_double, [C.c_int, C.c_int]),
    "fcs_pairhmm_kernel_class": (C.c_int, [C.c_int32, C.c_int32, i32p, i32p]),
}

# libfcs_pairhmm_client.so (no CUDA): client of the fcs-pairhmm-nam daemon
CLIENT_LIB_PATH = os.path.join(_HERE, "libfcs_pairhmm_client.so")
class RemoteViews(C.Structure):
    _fields_ = [("read_bases", u8p), ("read_q", u8p), ("read_i", u8p), ("read_d", u8p), ("read_c", u8p), ("rd_off", i64p), ("rd_len", i32p),
                ("hap_bases", u8p), ("hp_off", i64p), ("hp_len", i32p), ("reg_read0", i32p), ("reg_nreads", i32p), ("reg_hap0", i32p),
                ("reg_nhaps", i32p), ("out_log10", f64p), ("out_used_fp64", u8p)]


CLIENT_SIGNATURES = {
    "fcs_pairhmm_remote_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "fcs_pairhmm_remote_compute_flat": (C.c_int, [C.c_void_p, C.POINTER(FlatStruct), f64p, u8p]),
    "fcs_pairhmm_remote_last_error": (C.c_char_p, [C.c_void_p]),
    "fcs_pairhmm_remote_close": (None, [C.c_void_p]),
    "fcs_pairhmm_remote_uses_shm": (C.c_int, [C.c_void_p]),
    "fcs_pairhmm_remote_reserve": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(RemoteViews)]),
    "fcs_pairhmm_remote_compute_reserved": (C.c_int, [C.c_void_p]),
}

_lib = None
_client = None


def load_client() -> C.CDLL:
    global _client
    if _client is None:
        if not os.path.exists(CLIENT_LIB_PATH):
            raise OSError(f"{CLIENT_LIB_PATH} is missing: build it with `python __graft_entry__.py build`")
        lib = C.CDLL(CLIENT_LIB_PATH)
        for name, (res, args) in CLIENT_SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _client = lib
    return _client



def load() -> C.CDLL:
    """Load the in-tree shared library; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("FCS_PHMM_LIB", LIB_PATH)  # developer knob: A/B builds of the same library (tools/gpu/*.sh)
    if not os.path.exists(path):
        raise OSError(
            f"{path} is missing: build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a).  There is no CPU or pure-Python fallback for the PairHMM path.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the binding drift apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def as_u8p(a):
    return a.ctypes.data_as(u8p)
