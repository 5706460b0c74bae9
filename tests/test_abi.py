"""The C-ABI library loads and exports every symbol include/aiolos_b200.h declares; without a GPU it fails loudly."""
import ctypes
import os
import re

import pytest

import aiolos_b200 as A
from conftest import GOLDEN, ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "aiolos_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aio_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported(product_lib):
    lib = ctypes.CDLL(product_lib)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/aiolos_b200.h but not exported"


def test_python_binding_covers_header():
    assert set(A._SYMBOLS) == set(_declared_symbols())


def test_emulator_exports_same_abi(emu_lib):
    lib = ctypes.CDLL(emu_lib)
    for n in _declared_symbols():
        assert hasattr(lib, n)
    assert b"emulator" in ctypes.cast(lib.aio_version, ctypes.CFUNCTYPE(ctypes.c_char_p))()


def test_no_cpu_fallback(product_lib):
    """On a machine without a CUDA device the product path must refuse to run (AIO_ERR_CUDA), not compute on the CPU."""
    lib = A.load_library(product_lib)
    if lib.aio_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(A.AiolosError) as ei:
        A.Ensemble.from_dump(os.path.join(GOLDEN, "c1_planet_spherical_init.bin"))
    assert ei.value.code == -4, ei.value


def test_bad_arguments(emu_lib):
    with pytest.raises(A.AiolosError) as ei:
        A.Ensemble.from_dump("/nonexistent/dump.bin", lib=emu_lib)
    assert ei.value.code == -2
    with pytest.raises(A.AiolosError):
        A.Ensemble.from_dump(os.path.join(GOLDEN, "c4_dynamic3_euv_init.bin"), lib=emu_lib, overrides=[(0, "PARI_NOT_A_KEY", "1")])
    e = A.Ensemble.from_dump(os.path.join(GOLDEN, "c1_planet_spherical_init.bin"), lib=emu_lib)
    with pytest.raises(A.AiolosError):
        e.get("u1", column=5)
    with pytest.raises(A.AiolosError):
        e.set_option("tridiag", "gauss")
