#!/usr/bin/env python3
"""Coverage of the reference's own decks: runs the host front end + the engine's create-time checks on every shipped `.par`
file (dev container only: needs /root/reference and a working directory laid out by oracle/make_rundir.sh) and reports which
decks are accepted and why the others are refused.  Most decks do not name their species file (the reference takes it from
`-spc`), so every `.spc` of the deck's directory is tried and the first that parses wins.

Uses the CTA-emulator build of the library (tests/emu) because the dev container has no GPU; nothing is stepped.

usage: python tools/scan_decks.py [--rundir DIR] [--lib LIB] [--out report.md]"""
import argparse
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aiolos_b200 as A  # noqa: E402

REF = "/root/reference"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rundir", default="/tmp/aiolos_scan_run")
    ap.add_argument("--lib", default=os.path.join(ROOT, "tests", "emu", "_build", "libaiolos_emu.so"))
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if a.out:
        a.out = os.path.abspath(a.out)
    subprocess.run(["bash", os.path.join(ROOT, "oracle", "make_rundir.sh"), a.rundir], check=True, stdout=subprocess.DEVNULL)
    os.chdir(a.rundir)
    rows = []
    for p in sorted(glob.glob(os.path.join(REF, "*", "*.par"))):
        d = os.path.basename(os.path.dirname(p)) + "/"
        f = os.path.basename(p)
        if not os.path.isdir(d):
            continue
        spcs = [""] + sorted(os.path.basename(s) for s in glob.glob(os.path.join(os.path.dirname(p), "*.spc")))
        best = None
        for s in spcs:
            try:
                e = A.Ensemble.from_files(d, f, s, lib=a.lib)
                best = ("OK", s, e.num_cells, e.num_species, e.num_bands_in, e.num_bands_out)
                e.close()
                break
            except Exception as ex:      # noqa: BLE001  (every failure is a data point here)
                m = re.sub(r"aiolos_b200 error -?\d+: ", "", str(ex).replace("\n", " "))[:140]
                rank = 2 if "unsupported" in m else (0 if "cannot open species" in m else 1)
                if best is None or rank > best[0]:
                    best = (rank, s, m)
        rows.append((d + f, best))
    ok = [r for r in rows if r[1][0] == "OK"]
    why = collections.Counter()
    for r in rows:
        if r[1][0] != "OK":
            why[re.sub(r"parameterfile .*", "parameterfile", r[1][2])[:110]] += 1
    lines = [f"# Deck coverage: {len(ok)} of {len(rows)} shipped `.par` files accepted", "",
             f"largest accepted column: {max(r[1][2] for r in ok)} cells; most species: {max(r[1][3] for r in ok)}; "
             f"most bands: {max(r[1][4] for r in ok)}", "", "| refused | reason |", "|---|---|"]
    lines += [f"| {n} | {k} |" for k, n in why.most_common()]
    lines += ["", "Refused decks:", ""] + [f"* `{r[0]}` — {r[1][2]}" for r in rows if r[1][0] != "OK"]
    txt = "\n".join(lines) + "\n"
    print(txt)
    if a.out:
        open(a.out, "w").write(txt)


if __name__ == "__main__":
    main()
