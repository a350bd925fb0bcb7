#!/usr/bin/env python
"""Samples / instructions per CUDA source line of an ncu report: python tools/ncu_lines.py <rep> [top]"""
import csv, io, subprocess, sys, collections, os
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
smp = collections.Counter(); ins = collections.Counter(); txt = {}
fname, ismp, iins = "?", None, None
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        fname = os.path.basename(r[1]); continue
    if r and r[0] == "Line No":
        ismp, iins = r.index("# Samples"), r.index("Instructions Executed"); continue
    if ismp is None or len(r) <= iins or not r[0].strip().isdigit():
        continue
    try:
        s = float(r[ismp]); n = float(r[iins])
    except ValueError:
        continue
    key = (fname, int(r[0]))
    smp[key] += s; ins[key] += n; txt[key] = r[1]
ts, ti = sum(smp.values()), sum(ins.values())
print("samples %d  warp instructions %.3g" % (ts, ti))
for k, s in smp.most_common(top):
    print("%-18s %5d %5.1f%% smp %5.1f%% ins  %s" % (k[0][:18], k[1], 100 * s / ts, 100 * ins[k] / ti, txt[k].strip()[:100]))
