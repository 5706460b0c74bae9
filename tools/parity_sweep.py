#!/usr/bin/env python3
"""Parity sweep over the reference's own decks (dev container only): every shipped `.par` file the front end accepts
(tools/scan_decks.py logic) is advanced K steps by the UNMODIFIED reference (oracle/_ref/aiolos_ref_dump) and by the engine
sources running on the CTA emulator (tests/emu; the container has no GPU), and the states are compared with the metric of
the parity tests.  Test infrastructure: the oracle is only ever the checker here.

usage: python tools/parity_sweep.py [--steps 30] [--jobs 3] [--out report.md] [--only substring]"""
import argparse
import glob
import os
import subprocess
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"
RUN = "/tmp/aiolos_sweep_run"
EMU = os.path.join(ROOT, "tests", "emu", "_build", "libaiolos_emu.so")
ORACLE = os.path.join(ROOT, "oracle", "_ref", "aiolos_ref_dump")


def one(args):
    try:
        return _one(args)
    except Exception as ex:      # noqa: BLE001  (a worker must never die: every failure is a row of the report)
        return (args[0] + args[1], "error", str(ex).replace("\n", " ")[:160])


def _one(args):
    d, f, k = args
    import aiolos_b200 as A
    from aiodump import read_dump
    from parity import compare_with_golden
    os.chdir(RUN)
    spcs = [""] + sorted(os.path.basename(s) for s in glob.glob(os.path.join(REF, d, "*.spc")))
    ens, spc, why = None, None, "no species file parses"
    for s in spcs:
        try:
            ens = A.Ensemble.from_files(d, f, s, lib=EMU)
            spc = s
            break
        except Exception as ex:      # noqa: BLE001
            m = str(ex).replace("\n", " ")
            if "unsupported" in m or "cannot open species" not in m:
                why = m[:120]
    if ens is None:
        return (d + f, "refused", why)
    out = f"/tmp/aiolos_sweep_run/_dump_{os.getpid()}"
    cmd = [ORACLE, "-dir", d, "-par", f, "-nsteps", str(k), "-out", out] + (["-spc", spc] if spc else [])
    try:
        r = subprocess.run(cmd, stdin=subprocess.DEVNULL, capture_output=True, text=True, timeout=600)
    except subprocess.TimeoutExpired:
        return (d + f, "oracle-timeout", "")
    if r.returncode != 0 or not os.path.exists(out + "_final.bin"):
        return (d + f, "oracle-failed", (r.stderr or r.stdout)[-100:].replace("\n", " "))
    ref = read_dump(out + "_final.bin")
    os.remove(out + "_final.bin")
    kk = int(ref["steps"][0])
    ens.advance(kk)
    st = ens.status(0)
    refbad = not all(map(lambda v: v == v, [float(ref["globalTime"][0])]))
    errs = compare_with_golden(ens, ref)
    worst = max((v for v in errs.values() if v == v), default=0.0)
    nan = any(v != v for v in errs.values())
    dterr = abs(st["dt"] - ref["dt"][0]) / abs(ref["dt"][0]) if ref["dt"][0] else 0.0
    tag = "ok" if (worst < 1e-8 and dterr < 1e-9 and not nan) else ("both-crashed" if (st["state"] != 0 and (nan or refbad or kk < k)) else "MISMATCH")
    return (d + f, tag, f"spc={spc or '-'} cells={ens.num_cells} S={ens.num_species} B={ens.num_bands_in}/{ens.num_bands_out} steps={kk} "
                        f"state={st['state']} worst={worst:.1e} dt_err={dterr:.1e}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--jobs", type=int, default=3)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default="")
    ap.add_argument("--resume", default=None, help="log of an earlier run (one result tuple per line): those decks are kept, not rerun")
    a = ap.parse_args()
    out = os.path.abspath(a.out) if a.out else None
    subprocess.run(["bash", os.path.join(ROOT, "oracle", "make_rundir.sh"), RUN], check=True, stdout=subprocess.DEVNULL)
    work = []
    for p in sorted(glob.glob(os.path.join(REF, "*", "*.par"))):
        d = os.path.basename(os.path.dirname(p)) + "/"
        if os.path.isdir(os.path.join(RUN, d)) and a.only in p:
            work.append((d, os.path.basename(p), a.steps))
    rows = []
    if a.resume and os.path.exists(a.resume):
        import ast
        for ln in open(a.resume):
            if ln.startswith("('"):
                try:
                    rows.append(ast.literal_eval(ln.strip()))
                except (ValueError, SyntaxError):
                    pass
        done = {r[0] for r in rows}
        work = [w for w in work if w[0] + w[1] not in done]
    with ProcessPoolExecutor(a.jobs) as ex:
        for r in ex.map(one, work):
            rows.append(r)
            print(r, flush=True)
    n = {t: sum(1 for r in rows if r[1] == t) for t in ("ok", "both-crashed", "MISMATCH", "refused", "oracle-failed", "oracle-timeout", "error")}
    lines = [f"# Parity sweep over the shipped decks, {a.steps} steps each (CTA emulator vs the unmodified reference)", "",
             f"{len(rows)} decks: " + ", ".join(f"{v} {k}" for k, v in n.items() if v), "",
             "| deck | result | detail |", "|---|---|---|"]
    lines += [f"| `{r[0]}` | {r[1]} | {r[2]} |" for r in rows]
    txt = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(txt)


if __name__ == "__main__":
    main()
