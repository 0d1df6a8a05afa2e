"""CPU tests of the C-ABI boundary: the library loads, exports every symbol the header declares,
its lookup tables equal the oracle's bit for bit, and it fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

from falcon_genome_b200 import _lib, kernel_class

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "fcs_pairhmm.h")).read()
    return sorted(set(re.findall(r"FCS_PHMM_API[^;(]*?\b(fcs_pairhmm_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    client = _lib.load_client()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        owner = client if s.startswith("fcs_pairhmm_remote_") else lib  # the daemon's client library has no CUDA in it
        assert hasattr(owner, s), f"{s} declared in include/fcs_pairhmm.h but not exported"
    assert sorted(list(_lib.SIGNATURES) + list(_lib.CLIENT_SIGNATURES)) == syms, "ctypes binding and header drifted apart"
    assert lib.fcs_pairhmm_abi_version() == 1


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.Read) == 48 and C.sizeof(_lib.Hap) == 16 and C.sizeof(_lib.RegionStruct) == 48
    assert C.sizeof(_lib.FlatStruct) == 18 * 8
    assert C.sizeof(_lib.Config) == 48
    assert C.sizeof(_lib.Stats) == 13 * 8


def test_luts_bit_identical_to_oracle(oracle):
    lib = _lib.load()
    o = oracle.load()
    for q in range(128):
        assert lib.fcs_pairhmm_lut_ph2pr_f32(q) == o.phmm_oracle_ph2pr_f(q)
        assert lib.fcs_pairhmm_lut_ph2pr_f64(q) == o.phmm_oracle_ph2pr_d(q)
    for i in range(0, 128, 3):
        for d in range(0, 128, 5):
            assert lib.fcs_pairhmm_lut_mm_f32(i, d) == o.phmm_oracle_mm_f(i, d)
            assert lib.fcs_pairhmm_lut_mm_f64(i, d) == o.phmm_oracle_mm_d(i, d)
    assert lib.fcs_pairhmm_lut_mm_f64(45, 45) == lib.fcs_pairhmm_lut_mm_f64(45 + 128, 45)  # & 127
    assert abs(lib.fcs_pairhmm_lut_mm_f64(45, 45) - (1 - 2 * 10 ** -4.5)) < 1e-8


def test_kernel_class_covers_read_lengths():
    for L in range(1, 384):
        for f64 in (False, True):
            g, r = kernel_class(L, f64)
            assert g in (4, 8, 16, 32) and g * r >= L + 1
    assert kernel_class(150) == (8, 19)
    from falcon_genome_b200 import PairHMMError

    with pytest.raises(PairHMMError):
        kernel_class(5000)


def test_no_cpu_fallback_without_gpu():
    """On a box without a B200 the product must refuse to compute, not fall back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from falcon_genome_b200 import PairHMM, PairHMMError

    with pytest.raises(PairHMMError) as e:
        PairHMM()
    assert e.value.code == _lib.ENODEV and "no CPU fallback" in str(e.value)


def test_product_does_not_reference_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "falcon-genome_b200")
    for dp, _, files in os.walk(pkg):
        if "build" in dp.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in txt and "phmm_oracle_" not in txt, f
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert not re.search(r"#include\s+[\"<][^\">]*oracle", txt), f
