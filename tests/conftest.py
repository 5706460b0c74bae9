import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "_build", "libaiolos_emu.so")
REF_ROOT = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def emu_lib():
    """CTA emulator build of the engine sources (tests/emu): test infrastructure, never the product path."""
    r = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.fail("tests/emu build failed:\n" + r.stdout[-2000:] + r.stderr[-4000:])
    return EMU_LIB


@pytest.fixture(scope="session")
def emu_plugin_lib():
    """Emulator build with the reference's own problems/dusty_wind.cpp compiled in (make PROBLEM=..., dev container only)."""
    src = os.path.join(REF_ROOT, "problems", "dusty_wind.cpp")
    if not os.path.exists(src):
        pytest.skip("reference problem files not present (GPU box)")
    r = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu"), "PROBLEM=" + src], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.fail("tests/emu plugin build failed:\n" + r.stdout[-2000:] + r.stderr[-4000:])
    return os.path.join(ROOT, "tests", "emu", "_build", "dusty_wind", "libaiolos_emu_dusty_wind.so")


@pytest.fixture(scope="session")
def product_lib():
    import aiolos_b200
    if not os.path.exists(aiolos_b200.lib_path):
        import __graft_entry__
        __graft_entry__.build()
    return aiolos_b200.lib_path


@pytest.fixture(scope="session")
def oracle_bin():
    p = os.path.join(ROOT, "oracle", "_ref", "aiolos_ref_dump")
    if not os.path.isdir(REF_ROOT):
        pytest.skip("reference tree not present (GPU box): oracle is checked against the reference in the dev container")
    if not os.path.exists(p):
        r = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], capture_output=True, text=True)
        if r.returncode != 0:
            pytest.fail("oracle build failed:\n" + r.stderr[-4000:])
    return p
