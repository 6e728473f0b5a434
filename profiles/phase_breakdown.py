#!/usr/bin/env python
"""Aggregate `ncu --page source` per file and line range.
usage: phase_breakdown.py report.ncu-rep file_substr:lo-hi:name [...]   (no ranges: per-file totals + top lines)"""
import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file, hdr, data = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or r[0] in ("Function Name",) or len(r) < len(hdr):
        continue
    try:
        ln = int(r[0]); inst = float(r[hdr.index("Instructions Executed")] or 0)
        samp = float(r[hdr.index("# Samples")] or 0)
    except ValueError:
        continue
    data.append((cur_file, ln, inst, samp, r[1].strip()))
tot_i = sum(d[2] for d in data); tot_s = sum(d[3] for d in data)
print("total warp instructions %.4g, samples %d" % (tot_i, tot_s))
files = {}
for f, ln, i, s, _ in data:
    a = files.setdefault(f, [0, 0]); a[0] += i; a[1] += s
for f, (i, s) in sorted(files.items(), key=lambda x: -x[1][0]):
    print("  %-70s inst %5.1f%% samp %5.1f%%" % (f[-70:], 100 * i / tot_i, 100 * s / tot_s))
for spec in sys.argv[2:]:
    fs, rng, name = spec.split(":")
    lo, hi = map(int, rng.split("-"))
    i = sum(d[2] for d in data if fs in d[0] and lo <= d[1] <= hi)
    s = sum(d[3] for d in data if fs in d[0] and lo <= d[1] <= hi)
    print("%-28s L%d-%d  inst %5.1f%%  samp %5.1f%%" % (name, lo, hi, 100 * i / tot_i, 100 * s / tot_s))
