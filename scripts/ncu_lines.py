#!/usr/bin/env python
"""Rank CUDA source lines of an ncu report by executed instructions and stall samples.
usage: ncu_lines.py report.ncu-rep [top_n] [kernel-substring]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
want = sys.argv[3] if len(sys.argv) > 3 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
out = {}; cur = None; hdr = None; kern = None; active = False
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        active = want in r[1]; hdr = None
        if active and kern != r[1]:
            kern = r[1]; print("kernel:", kern[:100])
        continue
    if not active: continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "": continue
    d = {}
    for k, v in zip(hdr, r):
        d.setdefault(k, v)
    try:
        key = (cur, int(d["Line No"]))
        o = out.setdefault(key, [d["Source"].strip()[:84], 0, 0, 0, 0, 0])
        o[1] += int(d["Instructions Executed"] or 0); o[2] += int(d["# Samples"] or 0)
        o[3] += int(d["stall_long_sb"] or 0); o[4] += int(d["L2 Theoretical Sectors Global"] or 0)
        o[5] += int(d["Thread Instructions Executed"] or 0)
    except (ValueError, KeyError):
        continue
items = [(k[0], k[1]) + tuple(v) for k, v in out.items()]
ti = sum(o[3] for o in items) or 1; ts = sum(o[4] for o in items) or 1
print("total warp instructions %d, samples %d, avg threads/inst %.1f" % (ti, ts, sum(o[7] for o in items) / ti))
print("--- by instructions")
for o in sorted(items, key=lambda x: -x[3])[:top]:
    print("%-16s %4d %5.1f%% inst (%4.1f thr) %5.1f%% samp  %s" % (o[0], o[1], 100 * o[3] / ti, o[7] / max(1, o[3]), 100 * o[4] / ts, o[2]))
print("--- by stall samples (long_sb share)")
for o in sorted(items, key=lambda x: -x[4])[:top]:
    print("%-16s %4d %5.1f%% inst %5.1f%% samp  lsb %4.1f%% L2sect %10d  %s" % (o[0], o[1], 100 * o[3] / ti, 100 * o[4] / ts, 100 * o[5] / ts, o[6], o[2]))
