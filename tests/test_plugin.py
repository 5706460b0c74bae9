"""Link-time problem plugins: the reference's `make PROBLEM=name` links problems/name.cpp, which defines the
c_Species::user_* hooks.  Here the same, unmodified file is compiled against the facade header aiolos_b200/plugin/aiolos.h
into the engine's translation unit; the boundary hooks run inside the hydro kernels.  Oracle: the reference linked with the
same problem file (oracle/_ref/aiolos_ref_dump_dusty_wind) on PBC_model/dusty_wind.par at a quarter of its resolution."""
import os
import sys

import numpy as np
import pytest

import aiolos_b200 as A
from aiodump import read_dump
from conftest import GOLDEN, ROOT
from parity import compare_with_golden

sys.path.insert(0, os.path.join(ROOT, "aiolos_b200", "plugin"))
from prepare_problem import transform  # noqa: E402

DECK_DIR = os.path.join(GOLDEN, "decks") + "/"
RUN_DIR = os.path.join(ROOT, "decks")          # holds inputdata/smalldust.opa (opened relative to the CWD)


def test_prepare_problem_transform():
    src = """#include "aiolos.h"
// helper
static double P_vap(double T) {
    return 6.72e14 * std::exp(-1.0/T) ;
}
inline double twice(double x) { return 2 * x; }
static double T_eq ;
static int counter = 3;
static const double L = 3.21e10;
void c_Species::user_boundary_left(std::vector<AOS>& u) {
    static double inside = 1;
    double f(double);
    u[0].u1 = P_vap(T_eq) * inside;
}
void c_Species::user_initial_conditions(){
    T_eq = base->T_star ;
}
"""
    out = transform(src, "x.cpp").splitlines()
    assert out[0] == '#line 1 "x.cpp"'
    body = out[1:]
    assert body[2] == "__host__ __device__ static double P_vap(double T) {"
    assert body[5] == "__host__ __device__ inline double twice(double x) { return 2 * x; }"
    assert body[6] == "static __managed__ double T_eq ;"
    assert body[7] == "static __managed__ int counter = 3;"
    assert body[8] == "static const double L = 3.21e10;"                      # constants are usable on the device as they are
    assert body[9] == "void c_Species::user_boundary_left(std::vector<AOS>& u) {"   # execution space comes from the facade
    assert body[10] == "    static double inside = 1;" and body[11] == "    double f(double);"   # nothing inside a body is touched
    assert body[14] == "void c_Species::user_initial_conditions(){"
    assert len(body) == len(src.splitlines())                                 # line for line


def test_stock_library_rejects_hook_decks(emu_lib, monkeypatch):
    monkeypatch.chdir(RUN_DIR)
    with pytest.raises(A.AiolosError) as e:
        A.Ensemble.from_files(DECK_DIR, "dusty_wind_1k.par", "silicate.spc", lib=emu_lib)
    assert e.value.code == -3 and "PROBLEM=" in str(e.value)


def test_plugin_constructor_and_first_steps(emu_plugin_lib, monkeypatch, capfd):
    """user_initial_conditions + the constructor's boundary calls on the host, then 20 steps with both hooks in the kernel."""
    monkeypatch.chdir(RUN_DIR)
    ens = A.Ensemble.from_files(DECK_DIR, "dusty_wind_1k.par", "silicate.spc", lib=emu_plugin_lib)
    assert "Dusty wind setup: T_eq=" in capfd.readouterr().out                # the plugin's own host-side print
    ref0 = read_dump(os.path.join(GOLDEN, "p1_dusty_wind_init.bin"))
    assert ens.num_cells == 1002 and ens.num_species == 2
    for name in ["phi", "Jrad_FLD", "dt"] + [f"s{s}.{f}" for s in range(2) for f in ("u1", "u2", "u3", "boundary_left", "boundary_right")]:
        a, b = ens.problem_field(name), ref0[name]
        assert a.shape == b.shape and np.all(np.abs(a - b) <= 2e-12 * np.abs(b)), name
    assert ref0["s0.boundary_left"][0] == 0 and ref0["s0.boundary_right"][0] == 0
    ref = read_dump(os.path.join(GOLDEN, "p1_dusty_wind_k20.bin"))
    for i in range(20):
        ens.advance(1)
        assert abs(ens.status(0)["dt"] - ref["dt_sequence"][i]) <= 1e-12 * ref["dt_sequence"][i]
    errs = compare_with_golden(ens, ref)
    assert max(errs.values()) < 1e-10, errs


@pytest.mark.parametrize("name", ["cross_over", "dusty_wind", "default"])
def test_shipped_problem_files_compile_for_the_device(name, tmp_path):
    """Every problem file the reference ships goes through prepare_problem.py and nvcc (sm_100a) against the facade header,
    with both boundary hooks called from device code -- the compile-time half of "plugins build unchanged"."""
    import shutil
    import subprocess
    from conftest import REF_ROOT
    src = os.path.join(REF_ROOT, "problems", name + ".cpp")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(src) or not os.path.exists(nvcc):
        pytest.skip("reference problem files or nvcc not present")
    prepared = tmp_path / "problem.cu"
    prepared.write_text(transform(open(src).read(), src))
    tu = tmp_path / "tu.cu"
    tu.write_text('#include "%s"\n'
                  "__global__ void k(c_Species* sp) { sp->user_boundary_left(sp->u); sp->user_boundary_right(sp->u); }\n" % prepared)
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "--expt-relaxed-constexpr",
                        "-I", os.path.join(ROOT, "aiolos_b200", "plugin"), "-c", str(tu), "-o", str(tmp_path / "tu.o")],
                       capture_output=True, text=True)
    if name == "default":
        # the default problem file's hooks only throw ("not implemented"): a throw cannot be compiled for the device, which is
        # the intended diagnostic for hooks that use host-only facilities
        assert r.returncode != 0 and "device code does not support exception handling" in r.stderr
    else:
        assert r.returncode == 0, r.stderr[-3000:]
