import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import aiolos_b200 as A
from parity import run_case, compare_with_golden
for case, k in (("c1_planet_spherical", 100), ("c4_dynamic3_euv", 100)):
    for opts in ({}, {"steps_per_launch": "3"}, {"steps_per_launch": "4"}, {"steps_per_launch": "8"}, {"steps_per_launch": "100"}):
        ens, ref = run_case(A, None, case, k, options=opts)
        e = compare_with_golden(ens, ref)
        st = ens.status(0)
        print(case, k, opts, "steps", st["steps"], "t", st["t"], "ref t", ref["globalTime"][0], "maxerr %.3e" % max(e.values()), {a: "%.2e" % b for a, b in e.items() if b > 1e-9})
