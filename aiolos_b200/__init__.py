"""aiolos_b200 -- host-side Python mirror of the reference's simulation object for the per-timestep solver core.

The reference exposes one object, ``c_Sim(par, spc, dir, ...)`` with ``execute()`` (init_and_bounds.cpp:26,
advection.cpp:18).  ``Ensemble`` is the same thing for a batch of independent columns resident on one B200: it is a
ctypes binding of the C ABI in ``include/aiolos_b200.h`` (``libaiolos_b200.so``, hand-written sm_100a kernels).
There is no CPU implementation behind this module: if the CUDA library is missing or no GPU is visible, construction
raises.  (``tests/emu`` builds a CTA emulator of the same sources for debugging on GPU-less machines; it is only ever
loaded when a test passes its path explicitly.)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, Optional, Sequence, Tuple

import numpy as np

__all__ = ["Ensemble", "AiolosError", "load_library", "FIELDS", "lib_path"]

_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.path.join(_HERE, "libaiolos_b200.so")

FIELDS = {"u1": 0, "u2": 1, "u3": 2, "eint_last": 3, "phi": 4, "phi_s": 5, "dS": 6, "temperature": 7, "pres": 8,
          "speed": 9, "J": 16}
COL_STATES = {0: "ok", 1: "negative_T", 2: "negative_J", 3: "nan", 4: "hook_out_of_window"}


def plugin_lib_path(name: str) -> str:
    """In-tree library built with a problem plugin: ``make -C aiolos_b200/csrc PROBLEM=<.../name.cpp>`` (the reference's
    ``make PROBLEM=name``, makefile:22,28-29).  Pass it as ``lib=`` to ``Ensemble.from_files``."""
    return os.path.join(_HERE, f"libaiolos_b200_{name}.so")


class AiolosError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"aiolos_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg

    def __reduce__(self):           # picklable (multiprocessing workers)
        return (AiolosError, (self.code, self.msg))


class _Override(C.Structure):
    _fields_ = [("column", C.c_int), ("key", C.c_char_p), ("value", C.c_char_p)]


_SYMBOLS = {
    "aio_device_count": (C.c_int, []),
    "aio_device_peaks": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "aio_selftest_block_thomas": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "aio_create": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(_Override), C.c_int, C.c_int,
                             C.POINTER(C.c_void_p)]),
    "aio_create_from_dump": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_Override), C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "aio_destroy": (None, [C.c_void_p]),
    "aio_dims": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int)] * 5),
    "aio_advance": (C.c_int, [C.c_void_p, C.c_longlong, C.POINTER(C.c_double)]),
    "aio_run": (C.c_int, [C.c_void_p, C.c_double, C.c_longlong, C.POINTER(C.c_longlong)]),
    "aio_set_tmax": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "aio_kernel_times": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p]),
    "aio_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "aio_status": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong),
                             C.POINTER(C.c_int)]),
    "aio_status_all": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "aio_get": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "aio_problem_field": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_int]),
    "aio_upload_state": (C.c_int, [C.c_void_p, C.c_void_p]),
    "aio_download_state": (C.c_int, [C.c_void_p, C.c_void_p]),
    "aio_upload_state_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "aio_commit_upload": (C.c_int, [C.c_void_p]),
    "aio_column_mass": (C.c_int, [C.c_void_p, C.c_void_p]),
    "aio_write_outputs": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "aio_write_monitor": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int]),
    "aio_kernel_launches": (C.c_longlong, [C.c_void_p]),
    "aio_kernel_name": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int]),
    "aio_total_steps": (C.c_longlong, [C.c_void_p]),
    "aio_last_error": (C.c_char_p, []),
    "aio_version": (C.c_char_p, []),
}

_lib_cache = {}


def load_library(path: Optional[str] = None) -> C.CDLL:
    """Load the C-ABI library and declare every symbol of include/aiolos_b200.h.  No fallback of any kind."""
    path = path or lib_path
    if path in _lib_cache:
        return _lib_cache[path]
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(aiolos_b200 has no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in _SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib_cache[path] = lib
    return lib


def _mk_overrides(overrides: Optional[Iterable[Tuple[int, str, str]]]):
    items = list(overrides or [])
    arr = (_Override * max(len(items), 1))()
    keep = []
    for i, (col, key, val) in enumerate(items):
        k = str(key).encode(); v = (val if isinstance(val, str) else repr(float(val))).encode()
        keep += [k, v]
        arr[i].column = int(col); arr[i].key = k; arr[i].value = v
    return arr, len(items), keep


def device_peaks(device: int = 0, lib: Optional[str] = None, copy: bool = True):
    """Measured FP64 FMA rate (TFLOP/s) and copy bandwidth (GB/s) of one GPU: the roofline denominators bench.py uses."""
    L = load_library(lib)
    f, b = C.c_double(0.0), C.c_double(0.0)
    rc = L.aio_device_peaks(device, C.byref(f), C.byref(b) if copy else None)
    if rc != 0:
        raise AiolosError(rc, L.aio_last_error().decode())
    return {"fp64_tflops": f.value, "copy_gbs": b.value if copy else None}


def block_thomas(coef: np.ndarray, bands_out: int, num_species: int, lib: Optional[str] = None) -> np.ndarray:
    """Run the product's block-Thomas sweep (full FLD solver) on caller-supplied block rows; see aio_selftest_block_thomas.
    coef: [ncols][cells][NC]; returns x [ncols][cells][n]."""
    L = load_library(lib)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    ncols, cells, _ = coef.shape
    x = np.zeros((ncols, cells, bands_out + num_species))
    rc = L.aio_selftest_block_thomas(ncols, cells, bands_out, num_species, coef.ctypes.data, x.ctypes.data)
    if rc != 0:
        raise AiolosError(rc, L.aio_last_error().decode())
    return x


class Ensemble:
    """A batch of independent 1-D columns on one GPU, advanced by the sm_100a kernels."""

    def __init__(self, handle, lib):
        self._h = handle
        self._lib = lib
        n = [C.c_int() for _ in range(5)]
        self._ck(lib.aio_dims(handle, *[C.byref(x) for x in n]))
        self.ncols, self.num_cells, self.num_species, self.num_bands_in, self.num_bands_out = [x.value for x in n]

    # -- construction ------------------------------------------------------------------------------------------------------
    @classmethod
    def from_files(cls, directory: str, par: str, spc: str = "", ncols: int = 1,
                   overrides: Optional[Iterable[Tuple[int, str, str]]] = None, device: int = 0, lib: Optional[str] = None):
        """Same three strings as the reference's ``c_Sim`` constructor (``-dir``, ``-par``, ``-spc`` of main.cpp:48-99)."""
        L = load_library(lib)
        arr, n, _keep = _mk_overrides(overrides)
        h = C.c_void_p()
        rc = L.aio_create(directory.encode(), par.encode(), spc.encode(), ncols, arr, n, device, C.byref(h))
        if rc != 0:
            raise AiolosError(rc, L.aio_last_error().decode())
        return cls(h, L)

    @classmethod
    def from_dump(cls, path: str, ncols: int = 1, overrides: Optional[Iterable[Tuple[int, str, str]]] = None,
                  device: int = 0, lib: Optional[str] = None):
        """Start from the state the reference's constructor produced (an oracle init dump, tests/golden/)."""
        L = load_library(lib)
        arr, n, _keep = _mk_overrides(overrides)
        h = C.c_void_p()
        rc = L.aio_create_from_dump(path.encode(), ncols, arr, n, device, C.byref(h))
        if rc != 0:
            raise AiolosError(rc, L.aio_last_error().decode())
        return cls(h, L)

    def close(self):
        if self._h:
            self._lib.aio_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise AiolosError(rc, self._lib.aio_last_error().decode())

    # -- time loop -----------------------------------------------------------------------------------------------------------
    def advance(self, steps: int, timed: bool = False) -> float:
        """Advance every live column by up to ``steps`` steps; returns device milliseconds if ``timed``."""
        ms = C.c_double(0.0)
        self._ck(self._lib.aio_advance(self._h, int(steps), C.byref(ms) if timed else None))
        return ms.value

    def execute(self, t_stop: float = float("inf"), max_steps: int = 0) -> int:
        """``c_Sim::execute()``: run every column to its t_max (or ``t_stop``)."""
        n = C.c_longlong(0)
        self._ck(self._lib.aio_run(self._h, float(t_stop), int(max_steps), C.byref(n)))
        return n.value

    def set_tmax(self, t_max: float, column: int = -1):
        self._ck(self._lib.aio_set_tmax(self._h, column, float(t_max)))

    def kernel_times(self, steps: int):
        """Average device ms per step of the hydro / friction / radiation kernels over ``steps`` ordinary steps."""
        out = np.zeros(3)
        self._ck(self._lib.aio_kernel_times(self._h, int(steps), out.ctypes.data))
        return {"hydro": out[0], "friction": out[1], "radiation": out[2]}

    def kernel_name(self, slot) -> str:
        """Kernel(s) behind one slot of kernel_times ("hydro"/"step" = 0, "friction" = 1, "radiation" = 2)."""
        i = {"hydro": 0, "step": 0, "friction": 1, "radiation": 2}.get(slot, slot)
        buf = C.create_string_buffer(256)
        self._ck(self._lib.aio_kernel_name(self._h, int(i), buf, 256))
        return buf.value.decode()

    def set_option(self, key: str, value: str):
        self._ck(self._lib.aio_set_option(self._h, key.encode(), value.encode()))

    # -- inspection ----------------------------------------------------------------------------------------------------------
    def status(self, column: int = 0):
        t, dt, st, state = C.c_double(), C.c_double(), C.c_longlong(), C.c_int()
        self._ck(self._lib.aio_status(self._h, column, C.byref(t), C.byref(dt), C.byref(st), C.byref(state)))
        return {"t": t.value, "dt": dt.value, "steps": st.value, "state": state.value}

    def status_all(self):
        n = self.ncols
        t = np.empty(n); dt = np.empty(n); st = np.empty(n, dtype=np.int64); state = np.empty(n, dtype=np.int32)
        self._ck(self._lib.aio_status_all(self._h, t.ctypes.data, dt.ctypes.data, st.ctypes.data, state.ctypes.data))
        return {"t": t, "dt": dt, "steps": st, "state": state}

    def get(self, field, column: int = 0, species: int = 0, band: int = 0) -> np.ndarray:
        fid = FIELDS[field] if isinstance(field, str) else int(field)
        if fid == FIELDS["J"]:
            fid += band
        out = np.empty(self.num_cells + 2)
        self._ck(self._lib.aio_get(self._h, column, species, fid, out.ctypes.data, out.size))
        return out

    def problem_field(self, name: str, column: int = 0) -> np.ndarray:
        """An array/scalar the front end built for this column, by the reference's member name."""
        n = self._lib.aio_problem_field(self._h, column, name.encode(), None, 0)
        if n < 0:
            raise AiolosError(n, self._lib.aio_last_error().decode())
        out = np.empty(n)
        self._lib.aio_problem_field(self._h, column, name.encode(), out.ctypes.data, n)
        return out

    def upload_state(self, u: np.ndarray):
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert u.shape == (3, self.ncols, self.num_species, self.num_cells + 2)
        self._ck(self._lib.aio_upload_state(self._h, u.ctypes.data))

    def upload_state_async(self, u: np.ndarray):
        """Start copying a (pinned) host state into the staging buffer on the copy stream; pair with commit_upload()."""
        assert u.dtype == np.float64 and u.flags["C_CONTIGUOUS"] and u.shape == (3, self.ncols, self.num_species, self.num_cells + 2)
        self._ck(self._lib.aio_upload_state_async(self._h, u.ctypes.data))

    def commit_upload(self):
        self._ck(self._lib.aio_commit_upload(self._h))

    def download_state(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty((3, self.ncols, self.num_species, self.num_cells + 2))
        self._ck(self._lib.aio_download_state(self._h, out.ctypes.data))
        return out

    def column_mass(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.ncols)
        self._ck(self._lib.aio_column_mass(self._h, out.ctypes.data))
        return out

    def write_outputs(self, column: int, index: int, outdir: str = ""):
        self._ck(self._lib.aio_write_outputs(self._h, column, index, outdir.encode()))

    def write_monitor(self, column: int, outdir: str = "", truncate: bool = False):
        """``c_Sim::print_monitor``: one row of monitor_<stem>.dat for this column's current state."""
        self._ck(self._lib.aio_write_monitor(self._h, column, outdir.encode(), int(truncate)))

    @property
    def kernel_launches(self) -> int:
        return self._lib.aio_kernel_launches(self._h)

    @property
    def total_steps(self) -> int:
        return self._lib.aio_total_steps(self._h)
