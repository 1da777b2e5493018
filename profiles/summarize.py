#!/usr/bin/env python
"""Turn ncu CSV exports (launch list / --page raw) into the markdown tables kept in profiles/.

  python profiles/summarize.py launches <launches.csv>
  python profiles/summarize.py raw <raw.csv>        (from: ncu -i x.ncu-rep --page raw --csv)
"""
import collections
import csv
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
       "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
       "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
       "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v = v / 1e3 if d["Metric Unit"] in ("nsecond", "ns") else v
        agg.setdefault(d["Kernel Name"].split("(")[0].replace("void ", "")[:60] + " grid=" + d.get("Grid Size", ""), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | mean µs | share |\n|---|---|---|---|")
    for k, v in agg.items():
        print("| `%s` | %d | %.1f | %.1f %% |" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    names = [r[idx["Kernel Name"]].split("(")[0].replace("void ", "")[:34] for r in rows[2:]]
    print("| metric | unit | " + " | ".join("`%s`" % n for n in names) + " |\n|---|---|" + "---|" * len(names))
    for w in RAW:
        if w in idx:
            print("| %s | %s | " % (w, units[idx[w]]) + " | ".join(r[idx[w]][:12] for r in rows[2:]) + " |")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
