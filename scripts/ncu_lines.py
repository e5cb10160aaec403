#!/usr/bin/env python
"""Rank CUDA source lines of an ncu report by executed instructions and stall samples.
usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
out = []; cur = None; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Name": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "": continue
    d = {}
    for k, v in zip(hdr, r):
        d.setdefault(k, v)
    try:
        out.append((cur, int(d["Line No"]), d["Source"].strip()[:84], int(d["Instructions Executed"] or 0),
                    int(d["# Samples"] or 0), int(d["stall_long_sb"] or 0), int(d["L2 Theoretical Sectors Global"] or 0)))
    except (ValueError, KeyError):
        continue
ti = sum(o[3] for o in out) or 1; ts = sum(o[4] for o in out) or 1
print("total warp instructions %d, samples %d" % (ti, ts))
print("--- by instructions");
for o in sorted(out, key=lambda x: -x[3])[:top]:
    print("%-16s %4d %5.1f%% inst %5.1f%% samp  %s" % (o[0], o[1], 100 * o[3] / ti, 100 * o[4] / ts, o[2]))
print("--- by stall samples (long_sb share)")
for o in sorted(out, key=lambda x: -x[4])[:top]:
    print("%-16s %4d %5.1f%% inst %5.1f%% samp  lsb %4.1f%% L2sect %10d  %s" % (o[0], o[1], 100 * o[3] / ti, 100 * o[4] / ts, 100 * o[5] / ts, o[6], o[2]))
