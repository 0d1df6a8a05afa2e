"""f1, worker-side knob: integration/fcs-genome-pairhmm.patch applies to the reference's command builders and the
patched code emits the flags that make the GATK JVM load the B200 PairHMM shim.

House style of the reference's own worker tests (golden command strings, /root/reference/test/TestWorker.cpp:401):
the patched `HTCWorker::setup` / `Mutect2Worker::setup` bodies are compiled HERE into a small harness (config
lookups, BamInput and glog stubbed; the reference needs Boost/glog/a private deps server to build for real) and
the command strings they produce are asserted by regex, with the knob off (byte-identical to the unpatched code)
and on.  Nothing of the reference is stored in this repo: the harness is generated at test time from
/root/reference, which exists only in the build container (skipped on the GPU box).
"""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PATCH = os.path.join(ROOT, "integration", "fcs-genome-pairhmm.patch")
FILES = ["src/config.cpp", "src/workers/HTCWorker.cpp", "src/workers/Mutect2Worker.cpp"]

pytestmark = pytest.mark.skipif(not os.path.isdir(REF) or shutil.which("patch") is None, reason="needs /root/reference and patch(1)")

HARNESS = r"""
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
static std::map<std::string, std::string> g_cfg;
template <typename T> T get_config(const std::string& k, const std::string& dflt_key = "");
template <> std::string get_config<std::string>(const std::string& k, const std::string&) { return g_cfg.count(k) ? g_cfg[k] : std::string(); }
template <> int get_config<int>(const std::string& k, const std::string& d) { return std::stoi(g_cfg.count(k) ? g_cfg[k] : g_cfg[d]); }
template <> bool get_config<bool>(const std::string& k, const std::string&) { return g_cfg.count(k) && g_cfg[k] == "1"; }
struct NullLog { template <typename T> NullLog& operator<<(const T&) { return *this; } };
#define INFO 0
#define DLOG(x) NullLog()
struct BamInput {
  enum InputType { DEFAULT, NORMAL, TUMOR };
  std::string tag;
  std::string get_gatk_args(int contig, InputType = DEFAULT) { return " -I " + tag + "/part-" + std::to_string(contig) + ".bam "; }
};
struct HTCWorker {
  int contig_ = 3; bool produce_vcf_ = false; bool flag_gatk_ = false; std::string ref_path_ = "ref.fa";
  std::vector<std::string> intv_paths_{"p3.list"}; BamInput input_paths_{"in"}; std::string output_path_ = "out.g.vcf";
  std::map<std::string, std::vector<std::string>> extra_opts_; std::string cmd_;
  void setup();
};
struct Mutect2Worker {
  std::string ref_path_ = "ref.fa"; std::vector<std::string> intv_path_{"p3.list"}; BamInput normal_path_{"normal"}, tumor_path_{"tumor"};
  std::string output_path_ = "out.vcf"; std::vector<std::string> dbsnp_path_, cosmic_path_; std::string germline_path_, panels_of_normals_;
  std::string normal_name_ = "N", tumor_name_ = "T"; int contig_ = 3; bool flag_gatk_ = false;
  std::map<std::string, std::vector<std::string>> extra_opts_; std::string cmd_;
  void setup();
};
@HTC_SETUP@
@M2_SETUP@
int main(int argc, char** argv) {
  // argv: gatk4(0/1) lib_path extra_opt_key
  g_cfg = {{"java_path", "java -d64"}, {"gatk.memory", "8"}, {"gatk.nct", "4"}, {"gatk4_path", "GATK4.jar"}, {"gatk_path", "GATK3.jar"},
           {"use_gatk4", argv[1]}, {"gatk.pairhmm.lib_path", argv[2]}};
  HTCWorker h; Mutect2Worker m;
  if (argc > 3 && argv[3][0]) { h.extra_opts_[argv[3]] = {"X"}; m.extra_opts_[argv[3]] = {"X"}; }
  h.setup(); m.setup();
  std::cout << h.cmd_ << "\n" << m.cmd_ << "\n";
}
"""


def _setup_body(src, cls):
    m = re.search(r"void %s::setup\(\)\s*\{" % cls, src)
    assert m, cls
    depth, i = 0, m.end() - 1
    while True:
        depth += {"{": 1, "}": -1}.get(src[i], 0)
        i += 1
        if depth == 0:
            return src[m.start():i]


def _build(tmp_path, tree, name):
    code = HARNESS.replace("@HTC_SETUP@", _setup_body(open(os.path.join(tree, FILES[1])).read(), "HTCWorker"))
    code = code.replace("@M2_SETUP@", _setup_body(open(os.path.join(tree, FILES[2])).read(), "Mutect2Worker"))
    cpp, exe = str(tmp_path / f"{name}.cpp"), str(tmp_path / name)
    open(cpp, "w").write(code)
    r = subprocess.run(["/usr/bin/g++", "-std=c++11", "-O0", "-w", cpp, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def _cmds(exe, gatk4, lib, extra=""):
    out = subprocess.run([exe, "1" if gatk4 else "0", lib, extra], capture_output=True, text=True, check=True).stdout.splitlines()
    return out[0], out[1]


def test_patch_applies_and_patched_workers_emit_the_pairhmm_flags(tmp_path):
    orig, patched = str(tmp_path / "orig"), str(tmp_path / "patched")
    for tree in (orig, patched):
        for f in FILES:
            os.makedirs(os.path.dirname(os.path.join(tree, f)), exist_ok=True)
            shutil.copy(os.path.join(REF, f), os.path.join(tree, f))
    dry = subprocess.run(["patch", "-p1", "--dry-run", "-i", PATCH], cwd=patched, capture_output=True, text=True)
    assert dry.returncode == 0 and "FAILED" not in dry.stdout and "fuzz" not in dry.stdout, dry.stdout + dry.stderr
    assert subprocess.run(["patch", "-p1", "-i", PATCH], cwd=patched, capture_output=True, text=True).returncode == 0
    # the new key sits with the accelerator keys of src/config.cpp:353-354
    conf = open(os.path.join(patched, FILES[0])).read()
    assert re.search(r'arg_decl_string_w_def\("blaze\.conf_path".*\n\s*arg_decl_string_w_def\("gatk\.pairhmm\.lib_path",\s*""', conf)
    exe_o, exe_p = _build(tmp_path, orig, "orig_harness"), _build(tmp_path, patched, "patched_harness")
    lib = "/opt/fcs-pairhmm/lib"
    for gatk4 in (True, False):
        # knob off: the patched builders are byte-identical to the reference's
        assert _cmds(exe_p, gatk4, "") == _cmds(exe_o, gatk4, "")
        htc, m2 = _cmds(exe_p, gatk4, lib)
        for cmd, tool in ((htc, "HaplotypeCaller"), (m2, "Mutect2" if gatk4 else "MuTect2")):
            # JVM options come before -jar: GKL's NativeLibraryLoader then takes libgkl_pairhmm.so from java.library.path [upstream]
            assert re.search(r"^java -d64 -Xmx8g -DUSE_LIBRARY_PATH=true -Djava\.library\.path=%s -jar \S+ (-T )?%s " % (re.escape(lib), tool), cmd), cmd
            if gatk4:
                assert len(re.findall(r"--pair-hmm-implementation AVX_LOGLESS_CACHING ", cmd)) == 1, cmd
                assert len(re.findall(r"--native-pair-hmm-threads=4 ", cmd)) == 1, cmd  # HTCWorker.cpp:85; Mutect2 gets it with the knob
                assert "-pairHMM" not in cmd
            else:
                assert len(re.findall(r"-pairHMM VECTOR_LOGLESS_CACHING -nct 4 ", cmd)) == 1, cmd
                assert "--pair-hmm-implementation" not in cmd
        # everything else of the command line is untouched
        strip = lambda c: re.sub(r"-DUSE_LIBRARY_PATH=true -Djava\.library\.path=\S+ |--pair-hmm-implementation AVX_LOGLESS_CACHING |-pairHMM VECTOR_LOGLESS_CACHING ", "", c)  # noqa: E731
        o_htc, o_m2 = _cmds(exe_o, gatk4, lib)
        assert strip(htc) == o_htc
        assert re.sub(r"--native-pair-hmm-threads=4 ", "", strip(m2)) == o_m2
    # a user's own choice through -O/--extra-options wins (override rule of test/bats/cases/extra-opts-check.bats)
    htc, m2 = _cmds(exe_p, True, lib, "--pair-hmm-implementation")
    assert "AVX_LOGLESS_CACHING" not in htc and "AVX_LOGLESS_CACHING" not in m2 and "--pair-hmm-implementation X" in htc
    htc, m2 = _cmds(exe_p, False, lib, "-pairHMM")
    assert "VECTOR_LOGLESS_CACHING" not in htc and "VECTOR_LOGLESS_CACHING" not in m2
