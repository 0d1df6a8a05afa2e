"""Summarise an .ncu-rep (read here, no GPU needed) into the metrics DESIGN.md / bench.py cite.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--traffic-key c2 --source profiles/<name>.md] > profiles/<name>.md
With --traffic-key the first kernel's DRAM bytes are written to profiles/ncu_traffic.json under that key: bench.py
takes roofline.traffic from there (never from a constant in the code)."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return int(round(v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)))


def main():
    import argparse
    import json
    import os

    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--traffic-key")
    ap.add_argument("--source", default="")
    a = ap.parse_args()
    rep = a.rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu summary of `{rep}`\n")
    for n_row, r in enumerate(rows[2:]):
        d = dict(zip(hdr, r))
        if a.traffic_key and n_row == 0:
            root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
            path = os.path.join(root, "profiles", "ncu_traffic.json")
            try:
                table = json.load(open(path))
            except Exception:
                table = {}
            u = lambda k: units[hdr.index(k)]  # noqa: E731
            table[a.traffic_key] = {
                "kernel": d.get("Kernel Name"), "grid": d.get("Grid Size"),
                "dram_bytes_read": to_bytes(d["dram__bytes_read.sum"], u("dram__bytes_read.sum")),
                "dram_bytes_write": to_bytes(d["dram__bytes_write.sum"], u("dram__bytes_write.sum")),
                "gpu_time_us": float(d["gpu__time_duration.sum"].replace(",", "")) * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u("gpu__time_duration.sum"), 1),
                "registers_per_thread": int(d["launch__registers_per_thread"]),
                "fma_pipe_pct": float(d["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]),
                "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                "source": a.source, "capture": "ncu --set full --clock-control none, one launch (cold cache, serialised)"}
            json.dump(table, open(path, "w"), indent=1, sort_keys=True)
        print(f"## {d.get('Kernel Name', '?')}  grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d:
                print(f"| {k} | {d[k]} | {units[hdr.index(k)]} |")
        print("\nwarp stall reasons (warps per issue-active cycle):\n")
        print("| reason | ratio |\n|---|---|")
        st = []
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(d[k].replace(",", "")), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        for v, k in sorted(st, reverse=True):
            if v >= 0.01:
                print(f"| {k} | {v:.3f} |")
        print()


if __name__ == "__main__":
    main()
