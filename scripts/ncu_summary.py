#!/usr/bin/env python
"""Summarise an ncu report (key metrics per captured kernel) and/or a launch-list csv.
usage: ncu_summary.py --rep report.ncu-rep [--launches launches.csv]"""
import argparse, csv, subprocess, collections
ap = argparse.ArgumentParser(); ap.add_argument("--rep"); ap.add_argument("--launches"); a = ap.parse_args()
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
if a.launches:
    rows = list(csv.reader(open(a.launches)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    ki, vi = rows[h].index("Kernel Name"), rows[h].index("Metric Value")
    tot, cnt = collections.defaultdict(float), collections.defaultdict(int)
    for r in rows[h + 2:]:
        if len(r) > vi:
            n = r[ki].split("(")[0][:60]; tot[n] += float(r[vi].replace(",", "")); cnt[n] += 1
    s = sum(tot.values())
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print("| %s | %d | %.3f | %.1f%% |" % (k, cnt[k], v / 1e6, 100 * v / s))
if a.rep:
    txt = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        print("\n### %s" % d["Kernel Name"][:80])
        for k in KEYS:
            if k in d:
                print("%-78s %s %s" % (k, d[k], u[k]))
