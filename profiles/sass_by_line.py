#!/usr/bin/env python3
"""Join an `ncu --page source --csv` export (per-SASS-instruction executed counts and stall samples) with the line table of the
same build (`nvdisasm -g -c` of the cubin) and print executed warp instructions / stall samples per source line.
usage: python profiles/sass_by_line.py <ncu_source.csv> <nvdisasm.sass> <mangled-kernel-substring> [file-filter] [top]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, kern = sys.argv[1:4]
ffilter = sys.argv[4] if len(sys.argv) > 4 else ""
top = int(sys.argv[5]) if len(sys.argv) > 5 else 60
# ---- line table of the kernel's SASS, in instruction order
lines = open(sass).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":"))
cur = None
table = []
for l in lines[start + 1:]:
    if l.startswith("\t.section") or (l.startswith(".text.") and l.rstrip().endswith(":")):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        chain = [(m.group(1), int(m.group(2)))] + [(a, int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))]
        cur = chain
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        table.append((m.group(2).strip(), cur))
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci, cs, cn = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
body = [r for r in rows[hi + 1:] if len(r) > ci]
if len(body) != len(table):
    print(f"warning: {len(body)} profiled instructions vs {len(table)} in the disassembly", file=sys.stderr)
ex, sm, fp = defaultdict(int), defaultdict(int), defaultdict(int)
tot = tots = 0
for (txt, chain), r in zip(table, body):
    n, s = int(r[ci]), int(r[cn])
    key = None
    if chain:
        for f, ln in chain:                      # innermost frame inside the filtered file
            if ffilter in f:
                key = (f.split("/")[-1], ln)
                break
        if key is None:
            key = (chain[-1][0].split("/")[-1], chain[-1][1])
    ex[key] += n; sm[key] += s; tot += n; tots += s
    op = txt.split()[1] if txt.startswith("@") else txt.split()[0]
    if op.startswith(("DFMA", "DMUL", "DADD", "DSETP")):
        fp[key] += n
print(f"executed {tot}, samples {tots}")
for k, v in sorted(ex.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{str(k[0]):18s}:{k[1]:<5d} exec {v:12d} {100. * v / tot:5.2f}%  fp64 {100. * fp[k] / max(v, 1):5.1f}%  samples {100. * sm[k] / max(tots, 1):5.2f}%")
