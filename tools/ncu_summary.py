"""Summarise ncu artefacts into text for profiles/: (a) a --page raw --csv export of a --set full report,
(b) a gpu__time_duration launch list.  Usage:
    ncu -i rep.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py full raw.csv
    python tools/ncu_summary.py launches launches.csv
"""
import csv
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.min", "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "lts__t_sector_hit_rate.pct"]


def full(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("kernel:", r[col["Kernel Name"]][:110])
        for k in KEYS:
            if k in col and r[col[k]] != "":
                print(f"  {k:82s} {r[col[k]]:>18s} {units[col[k]]}")
        print()


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows[1:]:
        name = r[col["Kernel Name"]].split("(")[0]
        ns = float(r[col["Metric Value"]].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':60s} {'launches':>8s} {'total us':>12s} {'mean us':>10s} {'share':>7s}")
    for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:60s} {cnt:8d} {ns / 1e3:12.1f} {ns / 1e3 / cnt:10.2f} {100 * ns / tot:6.1f}%")
    print(f"{'TOTAL':60s} {sum(a[0] for a in agg.values()):8d} {tot / 1e3:12.1f}")


if __name__ == "__main__":
    (full if sys.argv[1] == "full" else launches)(sys.argv[2])
