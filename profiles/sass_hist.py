#!/usr/bin/env python3
"""Opcode histogram of an `ncu --page source --csv` export: executed warp instructions and stall samples per SASS opcode.
usage: python profiles/sass_hist.py gpurun_out/ncu_<tag>_source.csv [top]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci, cs, cn = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
ex, sm = defaultdict(int), defaultdict(int)
tot = tots = 0
n_static = 0
for r in rows[hdr_i + 1:]:
    if len(r) <= ci:
        continue
    txt = r[cs].strip()
    parts = txt.split()
    if not parts:
        continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    op = op.rstrip(";")
    base = op.split(".")[0] if not op.startswith(("LDS", "STS", "LDG", "STG", "SHFL", "REDUX", "MUFU")) else ".".join(op.split(".")[:2])
    try:
        n, s = int(r[ci]), int(r[cn])
    except ValueError:
        continue
    ex[base] += n; sm[base] += s; tot += n; tots += s; n_static += 1
print(f"static SASS instructions {n_static}; executed warp instructions {tot}; stall samples {tots}")
print(f"{'opcode':16s} {'executed':>14s} {'%':>6s} {'samples':>9s} {'%':>6s}")
for k, v in sorted(ex.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{k:16s} {v:14d} {100.0 * v / tot:6.2f} {sm[k]:9d} {100.0 * sm[k] / max(tots, 1):6.2f}")
