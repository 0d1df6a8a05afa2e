"""Summarise an .ncu-rep (read here, no GPU needed) into the metrics DESIGN.md / bench.py cite.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md"""
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


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu summary of `{rep}`\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
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
