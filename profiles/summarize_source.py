#!/usr/bin/env python
"""Summarise `ncu --page source --print-source cuda,sass --csv` output per CUDA source line.
usage: summarize_source.py report.ncu-rep [min_pct]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
import io
rows = list(csv.reader(io.StringIO(txt)))
hdr = None
lines = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r or r[0] in ("", "File Path", "Function Name"):
        continue
    if len(r) < len(hdr):
        continue
    try:
        float(r[hdr.index("Instructions Executed")] or 0)
    except ValueError:
        continue
    lines.append(r)
ci = {}
for k, n in enumerate(hdr):
    ci.setdefault(n, k)
def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot_i = sum(num(r[ci["Instructions Executed"]]) for r in lines)
tot_s = sum(num(r[ci["# Samples"]]) for r in lines)
print("total warp instructions %.4g, samples %d" % (tot_i, tot_s))
for r in lines:
    i = num(r[ci["Instructions Executed"]])
    s = num(r[ci["# Samples"]])
    if 100 * i / tot_i < min_pct and 100 * s / tot_s < min_pct:
        continue
    top = sorted(((num(r[ci[n]]), n) for n in stalls), reverse=True)[:2]
    tops = " ".join("%s=%d" % (n[6:], v) for v, n in top if v > 0)
    print("L%-4s inst %5.1f%% samp %5.1f%% thr/inst %4s | %-70s | %s" %
          (r[0], 100 * i / tot_i, 100 * s / tot_s, r[ci["Avg. Threads Executed"]],
           r[1].strip()[:70], tops))
