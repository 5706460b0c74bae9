"""INTEGRATION.md's main_b200.cpp -- the binding a maintainer of the reference would add -- is compiled from the document itself
against include/aiolos_b200.h, linked with the product library and run."""
import os
import re
import subprocess

import numpy as np
import pytest

import aiolos_b200 as A
from conftest import ROOT


@pytest.fixture(scope="module")
def example_exe(tmp_path_factory, product_lib):
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"```cpp\n(// main_b200\.cpp.*?)```", doc, flags=re.S)
    assert m, "main_b200.cpp listing not found in INTEGRATION.md"
    d = tmp_path_factory.mktemp("example")
    src, exe = d / "main_b200.cpp", d / "aiolos_b200_run"
    src.write_text(m.group(1))
    libdir = os.path.dirname(product_lib)
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), str(src), "-L", libdir, "-laiolos_b200",
                        "-Wl,-rpath," + libdir, "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return str(exe)


def test_example_links_and_fails_loudly_without_gpu(example_exe, product_lib):
    if A.load_library(product_lib).aio_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    r = subprocess.run([example_exe, "-dir", "runs_mdot_vs_euv_plot/", "-par", "dynamic3_euv1e3_C_1s2b_rerun.par", "-tmax", "1e-15"],
                       cwd=os.path.join(ROOT, "decks"), capture_output=True, text=True)
    assert r.returncode == 1 and "CUDA" in r.stderr, (r.returncode, r.stderr[-500:])


@pytest.mark.gpu
def test_example_runs_configs3(example_exe, product_lib, tmp_path):
    import shutil
    run = tmp_path / "decks"
    shutil.copytree(os.path.join(ROOT, "decks"), run)
    tmax = "3e-16"
    r = subprocess.run([example_exe, "-dir", "runs_mdot_vs_euv_plot/", "-par", "dynamic3_euv1e3_C_1s2b_rerun.par", "-ncols", "2", "-tmax", tmax],
                       cwd=run, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("column ")]
    assert len(lines) == 2 and all("state=0" in ln for ln in lines), r.stdout[-1000:]
    steps = [int(re.search(r"steps=(\d+)", ln).group(1)) for ln in lines]
    # the same run through the Python binding of the same ABI
    cwd = os.getcwd()
    os.chdir(run)
    try:
        ov = [(k, "PARI_UVSTAR", "%.17g" % (1e31 * 10.0 ** (-2.0 + 4.0 * k))) for k in range(2)]
        ens = A.Ensemble.from_files("runs_mdot_vs_euv_plot/", "dynamic3_euv1e3_C_1s2b_rerun.par", "", ncols=2, overrides=ov, lib=product_lib)
        ens.set_tmax(float(tmax))
        ens.execute()
        st = ens.status_all()
    finally:
        os.chdir(cwd)
    assert list(st["steps"]) == steps
    out = run / "runs_mdot_vs_euv_plot"
    assert any(f.startswith("output_") and f.endswith("_t-1.dat") for f in os.listdir(out)), os.listdir(out)
