#!/usr/bin/env python3
"""SURVEY 8(d) "science check": run the whole PARI_UVSTAR sweep of configs[3] to the deck's t_max (1.01e6 s, ~8e4 time steps
per column) on the GPU and compare a handful of columns with their CPU twins (the unmodified reference, one process per
host core, each with the same PARI_UVSTAR).  Reports, per twin, the mass-loss rate at the outer boundary
Mdot = 4 pi r^2 rho v, the step counts and the largest relative state differences, plus whether Mdot is monotone in F_UV over
the ensemble.  Test/measurement infrastructure: executes oracle/_ref (never part of the product path).

usage: python tools/science_check.py [--ncols 4096] [--twins N(=host cores)] [--tmax T] [--out file.json]"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import aiolos_b200 as A  # noqa: E402
import bench  # noqa: E402  (uv_value, deck names, reference process runner)
from aiodump import read_dump  # noqa: E402


def run_twins(uvs, tmax, workdir):
    """bench.run_reference_processes with a t_max instead of a step count; returns the final dumps."""
    import shutil
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "aiolos_ref_dump")
    procs, dirs = [], []
    for i, uv in enumerate(uvs):
        d = os.path.join(workdir, f"p{i}")
        shutil.copytree(os.path.join(ROOT, "decks"), d)
        par = os.path.join(d, bench.DECK_DIR, bench.DECK_PAR)
        txt = open(par).read().splitlines()
        txt = [("PARI_UVSTAR             %.17g" % uv) if ln.startswith("PARI_UVSTAR") else ln for ln in txt]
        open(par, "w").write("\n".join(txt) + "\n")
        dirs.append(d)
    t0 = time.perf_counter()
    for d in dirs:
        cmd = [exe, "-dir", bench.DECK_DIR, "-par", bench.DECK_PAR, "-out", "twin"]
        if tmax is not None:
            cmd += ["-tmax", repr(tmax)]
        procs.append(subprocess.Popen(cmd, cwd=d, stdin=subprocess.DEVNULL, stdout=open(os.path.join(d, "log.txt"), "w"),
                                      stderr=subprocess.STDOUT))
    for p in procs:
        p.wait()
    wall = time.perf_counter() - t0
    return [read_dump(os.path.join(d, "twin_final.bin")) for d in dirs], wall


def mdot(x, rho, mom, j):
    return 4.0 * np.pi * x[j] ** 2 * mom[j]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ncols", type=int, default=4096)
    ap.add_argument("--twins", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--tmax", type=float, default=None, help="override the deck's t_max (both sides)")
    ap.add_argument("--tridiag", default=None, help="aio_set_option tridiag value (e.g. thomas)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--lib", default=None, help="alternative C-ABI library (tests/emu build, for a CPU dry run of this script)")
    a = ap.parse_args()

    n = a.ncols
    ov = [(k, "PARI_UVSTAR", "%.17g" % bench.uv_value(k, n)) for k in range(n)]
    if a.tmax is not None:
        ov.append((-1, "PARI_TIME_TMAX", repr(a.tmax)))
    cwd = os.getcwd()
    lib = os.path.abspath(a.lib) if a.lib else None
    os.chdir(os.path.join(ROOT, "decks"))      # opacity files are opened relative to the CWD (io.cpp:143)
    try:
        ens = A.Ensemble.from_files(bench.DECK_DIR, bench.DECK_PAR, "", ncols=n, overrides=ov, lib=lib)
    finally:
        os.chdir(cwd)
    if a.tridiag:
        ens.set_option("tridiag", a.tridiag)
    t0 = time.perf_counter()
    ens.execute()
    gpu_wall = time.perf_counter() - t0
    st = ens.status_all()
    steps = np.asarray(st["steps"]); states = np.asarray(st["state"])
    x = ens.problem_field("x_i12", 0)
    assert x.size == ens.num_cells + 2
    jout = ens.num_cells - 1           # a cell well inside the domain next to the outer boundary
    md = np.array([mdot(x, ens.get("u1", c, 0), ens.get("u2", c, 0), jout) for c in range(n)])
    ok = states == 0
    mono = bool(np.all(np.diff(md[ok]) >= -1e-3 * np.abs(md[ok][1:]))) if ok.sum() > 1 else None
    viol = np.nonzero(np.diff(md) < -1e-3 * np.abs(md[1:]))[0]       # columns k where Mdot(k+1) < Mdot(k) by more than 1e-3

    tw_cols = sorted({int(round(i * (n - 1) / max(a.twins - 1, 1))) for i in range(a.twins)})
    with tempfile.TemporaryDirectory() as wd:
        dumps, cpu_wall = run_twins([bench.uv_value(c, n) for c in tw_cols], a.tmax, wd)
    twins = []
    for c, ref in zip(tw_cols, dumps):
        r1, r2, r3 = ref["s0.u1"], ref["s0.u2"], ref["s0.u3"]
        cs = ref["s0.sound_speed"]
        u1, u2, u3 = ens.get("u1", c, 0), ens.get("u2", c, 0), ens.get("u3", c, 0)
        md_ref = mdot(x, r1, r2, jout)
        twins.append({"column": c, "uvstar": bench.uv_value(c, n), "steps_gpu": int(steps[c]), "steps_cpu": int(ref["steps"][0]),
                      "t_gpu": float(st["t"][c]), "t_cpu": float(ref["globalTime"][0]),
                      "mdot_gpu": float(md[c]), "mdot_cpu": float(md_ref), "mdot_rel": float(abs(md[c] - md_ref) / abs(md_ref)),
                      "rho_rel_max": float(np.max(np.abs(u1 - r1) / np.abs(r1))),
                      "mom_rel_max": float(np.max(np.abs(u2 - r2) / np.abs(r1 * cs))),
                      "E_rel_max": float(np.max(np.abs(u3 - r3) / np.abs(r3)))})
    out = {"ncols": n, "t_max": a.tmax, "gpu_wall_s": gpu_wall, "gpu_total_steps": int(steps.sum()),
           "gpu_updates_per_s_wall": float(steps.sum()) * ens.num_cells / gpu_wall, "columns_ok": int(ok.sum()),
           "mdot_monotone_in_uv": mono,
           "monotone_violations": int(viol.size), "last_violating_column": (int(viol.max()) if viol.size else None),
           "first_violating_column": (int(viol.min()) if viol.size else None),
           "mdot_every_64th_column": [float(v) for v in md[::64]], "steps_every_64th_column": [int(v) for v in steps[::64]], "mdot_min": float(md[ok].min()), "mdot_max": float(md[ok].max()),
           "cpu_wall_s": cpu_wall, "cpu_processes": len(tw_cols), "twins": twins,
           "worst_mdot_rel": max(t["mdot_rel"] for t in twins)}
    txt = json.dumps(out, indent=1)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
