"""Compact, human-readable summary of a bench.py JSON line (developer tool): python tools/bench_summary.py FILE..."""
import json
import sys


def main():
    for path in sys.argv[1:]:
        d = json.loads([ln for ln in open(path) if ln.startswith("{")][-1])
        if "unavailable" in d:
            print(path, d)
            continue
        e = d.get("e2e", {})
        print(f"{path}: n_gpus {d.get('n_gpus')} impl {d.get('impl', 'b200')} value {d['value']:.0f} {d['unit']} ms/step {d.get('ms_per_step', 0):.3f} e2e {e.get('value', 0):.0f}"
              f" (h2d {e.get('h2d_bytes_per_step')} d2h {e.get('d2h_bytes_per_step')}) launches {d.get('gpu_launches')} roofline {d.get('roofline', {}).get('frac')}"
              f" parity {d.get('parity', {}).get('ok')} clocks {d.get('clocks', {}).get('sm_mhz')} {d.get('clocks', {}).get('reasons')}")
        rows = e.get("per_rank_ms_per_call", {}).get("rows")
        if rows:
            cols = e["per_rank_ms_per_call"]["columns"]
            print("  per rank ms per call (" + ", ".join(cols) + "):", "; ".join("/".join(f"{x:.2f}" for x in r) for r in rows))
        for k, c in (d.get("configs") or {}).items():
            print(f"  {k}: {c['value']:.0f} GCUPS, FP32 phase {100 * c['roofline']['frac']:.1f} %, whole step {100 * c['roofline']['whole_step_frac']:.1f} %, e2e {c['e2e']['value']:.0f},"
                  f" fp64 {c['fp64_pairs']}/{c['pairs']} at {100 * c.get('fp64_kernel', {}).get('frac_of_fp64_pipe', 0):.0f} % of its pipe, parity {c['parity']['ok']}")
        disp = d.get("e2e_dispatcher") or {}
        for k in ("c3_stream", "c4"):
            v = disp.get(k)
            if v:
                h = v["host_ms_per_call"]
                print(f"  dispatcher {k}: devices {disp.get('devices')} host threads {disp.get('host_threads')} callers {v['callers']}: {v['value']:.0f} GCUPS, {v['ms_per_call']:.2f} ms/call,"
                      f" chunks {v['chunks']}, host plan {h['host_plan_ms']:.1f} pack {h['host_pack_ms']:.1f} wait {h['host_wait_ms']:.1f} scatter {h['host_scatter_ms']:.1f} ms/call,"
                      f" == 1 device {v.get('equals_one_device_bitwise')} ok {v.get('ok')}")
        cb = d.get("cpu_baseline")
        if cb:
            print(f"  cpu_baseline: {cb['value']:.1f} {cb['unit']} on {cb['cores']} threads ({cb['kind']})")


if __name__ == "__main__":
    main()
