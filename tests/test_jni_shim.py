"""f1: the JNI shim cannot be built for real here (no JDK); type-check it against a stub jni.h and make
sure it exports exactly the three GKL symbol names GATK binds [upstream]."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "falcon-genome_b200", "jni", "fcs_pairhmm_jni.c")


def test_shim_type_checks_against_stub_jni_header():
    r = subprocess.run(["/usr/bin/gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "jni_stub"), "-I",
                        os.path.join(ROOT, "include"), SHIM], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_shim_exports_the_gkl_symbol_names():
    src = open(SHIM).read()
    names = re.findall(r"JNICALL\s+(Java_\w+)\s*\(", src)
    assert names == ["Java_com_intel_gkl_pairhmm_IntelPairHmm_initNative",
                     "Java_com_intel_gkl_pairhmm_IntelPairHmm_computeLikelihoodsNative",
                     "Java_com_intel_gkl_pairhmm_IntelPairHmm_doneNative"]
    for call in ("fcs_pairhmm_create", "fcs_pairhmm_compute", "fcs_pairhmm_destroy"):
        assert call in src


# ---- the shim's logic, driven through a fake JNIEnv (tests/jni_stub/fake_jvm.c) ------------------------

def _build_shim_and_fake_jvm(tmp_path):
    pkg = os.path.join(ROOT, "falcon-genome_b200")
    stub = os.path.join(ROOT, "tests", "jni_stub")
    shim_so = str(tmp_path / "libgkl_pairhmm.so")
    fake_so = str(tmp_path / "libfake_jvm.so")
    for cmd in (["/usr/bin/gcc", "-O2", "-fPIC", "-shared", "-Wall", "-Werror", "-I", stub, "-I", os.path.join(ROOT, "include"), SHIM,
                 "-L", pkg, "-l:libfcs_pairhmm.so", f"-Wl,-rpath,{pkg}", "-o", shim_so],
                ["/usr/bin/gcc", "-O2", "-fPIC", "-shared", "-Wall", "-Werror", "-fvisibility=hidden", "-I", stub,
                 os.path.join(stub, "fake_jvm.c"), "-ldl", "-o", fake_so]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    return shim_so, fake_so


def _run_region(fake_so, shim_so, b, g, repeats=1, use_double=0, max_threads=2, break_fields=0, info=None):
    """One region of a FlatBatch through initNative / computeLikelihoodsNative / doneNative.
    max_threads < 0: computeLikelihoodsNative WITHOUT initNative; break_fields: the holder classes lack a field."""
    import ctypes as C

    import numpy as np

    lib = C.CDLL(fake_so)
    lib.fake_jvm_set_mode(break_fields)
    u8p, i64p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    lib.fake_jvm_run.restype = C.c_int
    lib.fake_jvm_run.argtypes = [C.c_char_p, u8p, u8p, u8p, u8p, u8p, i64p, i32p, C.c_int32, u8p, i64p, i32p, C.c_int32, C.c_int32,
                                 C.c_int32, C.c_int32, C.POINTER(C.c_double), C.c_char_p, C.c_int32]
    r0, nr, h0, nh = int(b.reg_read0[g]), int(b.reg_nreads[g]), int(b.reg_hap0[g]), int(b.reg_nhaps[g])
    rd_off = np.ascontiguousarray(b.rd_off[r0:r0 + nr], np.int64)
    rd_len = np.ascontiguousarray(b.rd_len[r0:r0 + nr], np.int32)
    hp_off = np.ascontiguousarray(b.hp_off[h0:h0 + nh], np.int64)
    hp_len = np.ascontiguousarray(b.hp_len[h0:h0 + nh], np.int32)
    out = np.zeros(nr * nh, np.float64)
    err = C.create_string_buffer(600)
    planes = [np.ascontiguousarray(p, np.uint8) for p in (b.read_bases, b.read_q, b.read_i, b.read_d, b.read_c, b.hap_bases)]
    rc = lib.fake_jvm_run(shim_so.encode(), *[p.ctypes.data_as(u8p) for p in planes[:5]], rd_off.ctypes.data_as(i64p),
                          rd_len.ctypes.data_as(i32p), nr, planes[5].ctypes.data_as(u8p), hp_off.ctypes.data_as(i64p),
                          hp_len.ctypes.data_as(i32p), nh, use_double, max_threads, repeats, out.ctypes.data_as(C.POINTER(C.c_double)), err, 600)
    lib.fake_jvm_set_mode(0)
    if info is not None:
        info.update(max_live_refs=lib.fake_jvm_max_live_refs(), ref_capacity=lib.fake_jvm_ref_capacity(),
                    calls_with_pending=lib.fake_jvm_calls_with_pending_exception())
    return rc, out, err.value.decode()


def test_shim_raises_a_java_exception_without_a_gpu(tmp_path):
    """Host-only: the shim builds against the stub header, links the library and, with no B200, initNative
    surfaces the library's ENODEV text as a RuntimeException instead of falling back to anything."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import _pkg

    _pkg.load()
    from falcon_genome_b200 import synth

    shim_so, fake_so = _build_shim_and_fake_jvm(tmp_path)
    rc, _, err = _run_region(fake_so, shim_so, synth.tiny_mixed(seed=81, n_regions=1), 0)
    assert rc == -1 and err.startswith("java/lang/RuntimeException: ") and "no CPU fallback" in err, err


def test_shim_checks_field_ids_and_the_handle_before_touching_arrays(tmp_path):
    """Host-only robustness of the shim: (1) holder classes without the expected fields: the NoSuchFieldError GetFieldID
    leaves pending is checked at once (no further JNI call is made with it pending, no NULL jfieldID is ever used);
    (2) computeLikelihoodsNative without a successful initNative throws IllegalStateException before any array is
    read or pinned and leaks no local reference."""
    import _pkg

    _pkg.load()
    from falcon_genome_b200 import synth

    shim_so, fake_so = _build_shim_and_fake_jvm(tmp_path)
    b = synth.tiny_mixed(seed=81, n_regions=1)
    info = {}
    rc, _, err = _run_region(fake_so, shim_so, b, 0, break_fields=1, info=info)
    assert rc == -1 and err.startswith("java/lang/NoSuchFieldError"), err
    assert info["calls_with_pending"] == 0 and info["max_live_refs"] == 0, info
    rc, out, err = _run_region(fake_so, shim_so, b, 0, max_threads=-1, info=info)
    assert rc == -1 and err.startswith("java/lang/IllegalStateException: ") and "initNative" in err, err
    assert info["calls_with_pending"] == 0 and info["max_live_refs"] <= 1, info  # the exception class only
    assert (out == 1.0).all()  # the "Java" array was never written


@pytest.mark.gpu
def test_shim_deep_pileup_keeps_local_references_bounded(tmp_path, hmm, oracle):
    """A Mutect2-sized region (2000 reads x 12 haplotypes) through the fake JNIEnv: the shim never holds more local
    references than a native frame is guaranteed (16), leaks none, and the likelihoods the "JVM" receives pass the
    oracle's parity bars (float path and use_double path)."""
    import numpy as np

    from falcon_genome_b200 import FlatBatch, Region

    rng = np.random.default_rng(4242)
    hap = bytes(rng.choice(list(b"ACGT"), 330).astype(np.uint8))
    haps = []
    for _ in range(12):
        h = bytearray(hap)
        for k in rng.integers(0, len(h), 3):
            h[k] = int(rng.choice(list(b"ACGT")))
        haps.append(bytes(h))
    reads = []
    for _ in range(2000):
        L = 150 if rng.random() < 0.9 else int(rng.integers(60, 150))
        s0 = int(rng.integers(0, 330 - L))
        bs = bytearray(haps[int(rng.integers(0, 12))][s0:s0 + L])
        q = rng.integers(6, 42, L).astype(np.uint8)
        for k in np.nonzero(rng.random(L) < 10.0 ** (-q / 10.0))[0]:
            bs[k] = int(rng.choice(list(b"ACGT")))
        reads.append((bytes(bs), bytes(q), bytes([45] * L), bytes([45] * L), bytes([10] * L)))
    b = FlatBatch.from_regions([Region(reads, haps)])
    shim_so, fake_so = _build_shim_and_fake_jvm(tmp_path)
    info = {}
    rc, out, err = _run_region(fake_so, shim_so, b, 0, info=info)
    assert rc == 0, err
    assert info["max_live_refs"] <= 16 and info["max_live_refs"] <= info["ref_capacity"], info
    o_ref, u_ref, _, dbl = oracle.batch_scalar(b)
    assert np.abs(out - dbl).max() <= 1e-4 and np.abs(out - o_ref).max() <= 4e-6
    assert np.array_equal(out.reshape(2000, 12).argmax(1), o_ref.reshape(2000, 12).argmax(1))
    rc, outd, err = _run_region(fake_so, shim_so, b, 0, use_double=1)
    assert rc == 0, err
    assert np.abs(outd - dbl).max() <= 1e-9  # the FP64 kernels against the double-precision oracle


@pytest.mark.gpu
def test_shim_end_to_end_through_a_fake_jnienv(tmp_path, hmm, oracle):
    """GPU: VectorLoglessPairHMM's call sequence against the shim; the double[] the "JVM" holds afterwards equals the
    library's own result for the region, every pinned array was released (inputs with JNI_ABORT), nothing was thrown."""
    import numpy as np

    from falcon_genome_b200 import synth

    shim_so, fake_so = _build_shim_and_fake_jvm(tmp_path)
    b = synth.tiny_mixed(seed=82, n_regions=4)
    ref, _ = hmm.compute_flat(b)
    for g in range(b.n_regions):
        rc, out, err = _run_region(fake_so, shim_so, b, g, repeats=2)
        assert rc == 0, err
        o0 = int(b.reg_out0[g])
        assert np.array_equal(out, ref[o0:o0 + out.size])
        assert (out < 0).all()
    rc, out, err = _run_region(fake_so, shim_so, b, 0, use_double=1)  # use_double handle: every pair through the FP64 kernels
    assert rc == 0, err
    _, _, _, dbl = oracle.batch_scalar(b.select([0]))
    assert np.abs(out - dbl).max() <= 1e-9  # against the double-precision oracle, not against the library itself
