"""f1: the JNI shim cannot be built for real here (no JDK); type-check it against a stub jni.h and make
sure it exports exactly the three GKL symbol names GATK binds [upstream]."""
import os
import re
import subprocess

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
