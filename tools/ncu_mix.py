#!/usr/bin/env python
"""Instruction mix of an ncu report (source page): python tools/ncu_mix.py <rep> -- warp-instructions executed per opcode"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
H = rows[0]; ix = {h: i for i, h in enumerate(H)}
mix = collections.Counter(); smp = collections.Counter()
tot = 0
for r in rows[1:]:
    try:
        n = float(r[ix["Instructions Executed"]]); s = float(r[ix["# Samples"]])
    except Exception:
        continue
    src = r[ix["Source"]].split()
    op = src[1] if src and src[0].startswith("@") and len(src) > 1 else (src[0] if src else "?")
    op = op.split(".")[0]
    mix[op] += n; smp[op] += s; tot += n
st = sum(smp.values())
print("total warp instructions %.3g" % tot)
for op, n in mix.most_common(28):
    print("%-10s %10.3g  %5.1f%% of instructions  %5.1f%% of samples" % (op, n, 100 * n / tot, 100 * smp[op] / st))
