#!/usr/bin/env python
"""Hot SASS instructions of an ncu report (source page): python tools/ncu_hot.py <rep> [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
# first line = kernel name
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
H = rows[0]
ix = {h: i for i, h in enumerate(H)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
data = rows[1:]
tot = sum(f(r, "# Samples") for r in data)
print("total samples", tot)
stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(f(r, s) for r in data) for s in stalls}
print("by reason:", {k: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.005 * tot})
print("%5s %-8s %6s %9s  %-60s %s" % ("line", "addr", "smp%", "inst", "sass", "top stalls"))
order = sorted(range(len(data)), key=lambda i: -f(data[i], "# Samples"))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((f(r, s), s[6:]) for s in stalls), reverse=True)[:3]
    print("%5d %-8s %6.2f %9.3g  %-60s %s" % (i, r[ix["Address"]][-5:], 100 * f(r, "# Samples") / tot, f(r, "Instructions Executed"),
          r[ix["Source"]][:60], " ".join("%s:%.0f" % (n, v) for v, n in st if v > 0)))
