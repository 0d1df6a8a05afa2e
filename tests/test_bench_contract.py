"""Host-only checks of bench.py's contract pieces: the reference arm's JSON line (same `config` object as our arm would
print for the same command line), the parity scorer (accepts the oracle's own results, rejects a flipped decision, a
moved likelihood and a changed best haplotype) and the tracked ncu traffic table."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_reference_arm_prints_the_contract_line_with_our_config():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0",
                        "--workload", "c1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT and line["higher_is_better"] is True
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "not GKL" in line["cpu_baseline"]["sample"]

    class A:
        gpus = 1

    b, desc = bench.make_workload("c1", 0)
    assert line["config"] == bench.shared_config(A, desc, b)  # identical keys and values in both arms


def test_parity_scorer_accepts_the_oracle_and_rejects_deviations(oracle):
    from falcon_genome_b200 import synth

    b = synth.config1_golden(n_regions=12, seed=5)
    out, used, raw, _ = oracle.batch_simd(b)
    p = bench.parity_check(b, out, used, raw, nthreads=2)
    assert p["ok"] and p["fallback_mismatches"] == 0 and p["raw_f32_bit_mismatches"] == 0 and p["argmax_mismatches"] == 0
    assert p["max_abs_dlog10_vs_double_oracle"] <= 1e-4
    u2 = used.copy(); u2[3] ^= 1
    assert not bench.parity_check(b, out, u2, raw, nthreads=2)["ok"]
    o2 = out.copy(); o2[7] += 2e-4
    assert not bench.parity_check(b, o2, used, raw, nthreads=2)["ok"]
    r2 = raw.copy(); r2.view(np.uint32)[11] ^= 1
    assert not bench.parity_check(b, out, used, r2, nthreads=2)["ok"]
    # best haplotype of the first read moved to another column (within tolerance elsewhere)
    nh = int(b.reg_nhaps[0])
    if nh > 1:
        o3 = out.copy()
        row = o3[:nh]
        j = int(np.argmin(row))
        row[j] = row.max() + 5e-5
        assert bench.parity_check(b, o3, used, raw, nthreads=2)["argmax_mismatches"] >= 1


def test_traffic_comes_from_the_tracked_profile_table():
    t, src = bench.ncu_traffic("c2")
    table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    assert t == table["c2"]["dram_bytes_read"] + table["c2"]["dram_bytes_write"] and src["kernel"].startswith("phmm_f32a_tier2")
    assert bench.ncu_traffic("no-such-workload") == (None, None)
    assert "NCU_DRAM_BYTES_PER_LAUNCH" not in open(os.path.join(ROOT, "bench.py")).read()
