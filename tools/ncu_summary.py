#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers the roofline discussion needs.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--out profiles/x.txt]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        out.append(f"== {d.get('Kernel Name', ('?',))[0]}  (id {d.get('ID', ('?',))[0]})")
        for k in KEYS:
            if k in d:
                out.append(f"  {k:70s} {d[k][0]:>16s} {d[k][1]}")
        st = []
        for h in hdr:
            if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(d[h][0]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        out.append("  warps stalled per issue-active cycle: " + ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:9]))
    text = "\n".join(out)
    print(text)
    if "--out" in sys.argv:
        open(sys.argv[sys.argv.index("--out") + 1], "w").write(text + "\n")


if __name__ == "__main__":
    main()
