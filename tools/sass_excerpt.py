#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel, the counts of the mnemonics that identify the sm_100a features it uses
(bulk TMA, mbarrier, async copies, fp64 pipe, named barriers) and a short excerpt around the first DFMA block.

    python tools/sass_excerpt.py getfem_b200/csrc/_obj/recompute_tiles.o 'k_tilesILi3ELi3ELi10ELi1' > profiles/...txt
"""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", out)
KEYS = ["UBLKCP", "UTMACMDFLUSH", "SYNCS", "LDGSTS", "LDGDEPBAR", "DEPBAR", "BAR.SYNC", "BAR.ARV", "DMMA", "DFMA", "DMUL", "DADD", "LDS", "STS", "LDG",
        "STG", "SHFL", "ATOMS", "ATOMG", "RED", "NANOSLEEP", "LDC", "ULDC", "R2UR", "BRA", "CALL"]
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", b)
    cnt = collections.Counter()
    for i in ins:
        for k in KEYS:
            if i.startswith(k):
                cnt[k] += 1
    arch = re.search(r"EF_CUDA_SM(\d+)", out)
    print("kernel  :", subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip())
    print("arch    : sm_%s   instructions: %d" % (arch.group(1) if arch else "?", len(ins)))
    print("counts  :", "  ".join("%s %d" % (k, cnt[k]) for k in KEYS if cnt[k]))
    lines = [l for l in b.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    for tag in ("UBLKCP", "SYNCS", "LDGSTS", "BAR.SYNC", "BAR.ARV", "DMMA", "DFMA"):
        idx = next((k for k, l in enumerate(lines) if tag in l), None)
        if idx is not None:
            print("first %s:" % tag)
            for l in lines[max(0, idx - 1): idx + 3]:
                print("   ", re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l).strip()[:120])
    print()
