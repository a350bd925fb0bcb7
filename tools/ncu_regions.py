#!/usr/bin/env python
"""Instruction/sample totals of an ncu report by opcode class: python tools/ncu_regions.py <rep>"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
H = rows[0]; ix = {h: i for i, h in enumerate(H)}
data = rows[1:]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot_i = sum(f(r, "Instructions Executed") for r in data); tot_s = sum(f(r, "# Samples") for r in data)
cls = collections.Counter(); smp = collections.Counter()
for r in data:
    src = r[ix["Source"]].split()
    op = src[0] if not src[0].startswith("@") else src[1]
    op = op.split(".")[0]
    cls[op] += f(r, "Instructions Executed"); smp[op] += f(r, "# Samples")
print("total warp-instr %.3g samples %d" % (tot_i, tot_s))
for op, v in cls.most_common(25):
    print("%-10s inst %6.2f%%  (%.3g)   samples %5.2f%%" % (op, 100 * v / tot_i, v, 100 * smp[op] / tot_s))
if len(sys.argv) > 2:  # dump everything with exec counts
    for i, r in enumerate(data):
        print("%5d %9.3g %6.2f  %s" % (i, f(r, "Instructions Executed"), 100 * f(r, "# Samples") / tot_s, r[ix["Source"]][:90]))
