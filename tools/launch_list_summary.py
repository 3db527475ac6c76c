#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration.sum CSV) per kernel: launches, total, mean, share.

    python tools/launch_list_summary.py gpurun_out/launches.csv "command line" > profiles/x.txt
"""
import collections
import csv
import sys


def main():
    path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    ix = {h: i for i, h in enumerate(rows[0])}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]].split("(")[0][:70]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        agg.setdefault(name, []).append(v)
    mine = lambda k: "sti_" in k or "median" in k
    tot = sum(sum(v) for k, v in agg.items() if mine(k))
    print(f"# ncu launch list of `{cmd}` (gpu__time_duration, --clock-control none; cold-cache, serialised)")
    print("# kernel | launches | total us | mean us | share of this library's kernels")
    for k, v in agg.items():
        share = f"{100 * sum(v) / tot:5.1f}%" if mine(k) else "  (torch: synthetic input generation, outside the timed region)"
        print(f"{k:70s} {len(v):4d} {sum(v):12.1f} {sum(v) / len(v):10.1f} {share}")


if __name__ == "__main__":
    main()
