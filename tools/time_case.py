#!/usr/bin/env python3
"""Device-time any golden case on the GPU: ncols identical columns, spin-up, then K timed steps (CUDA events).
usage: python tools/time_case.py <case> [--ncols N] [--spinup M] [--steps K] [--bytes B]   (prints one JSON line)"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aiolos_b200 as A  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("case")
ap.add_argument("--ncols", type=int, default=4096)
ap.add_argument("--spinup", type=int, default=600)
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--bytes", type=float, default=0.0, help="algorithmic bytes per cell-species-update (SURVEY 8d)")
ap.add_argument("--lib", default=None, help="path of an alternative build of the library")
ap.add_argument("--opt", action="append", default=[], help="engine option key=value (repeatable)")
a = ap.parse_args()
ens = A.Ensemble.from_dump(os.path.join(ROOT, "tests", "golden", f"{a.case}_init.bin"), ncols=a.ncols, lib=a.lib)
for kv in a.opt:
    k_, v_ = kv.split("=", 1)
    ens.set_option(k_, v_)
ens.advance(a.spinup)
s0 = ens.total_steps
ms = ens.advance(a.steps, timed=True)
cs = ens.total_steps - s0
upd = cs * ens.num_cells * ens.num_species
kt = ens.kernel_times(min(a.steps, 10))
out = {"case": a.case, "ncols": a.ncols, "num_cells": ens.num_cells, "num_species": ens.num_species, "steps": a.steps,
       "ms_per_step": ms / a.steps, "updates_per_s": upd / (ms * 1e-3), "kernel_ms": kt,
       "state": ens.status(0), "options": a.opt}
if a.bytes:
    out["hbm_frac_of_6536GBs"] = out["updates_per_s"] * a.bytes / 6536.4e9
print(json.dumps(out))
