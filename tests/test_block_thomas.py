"""Kernel-level known-answer tests of the coupled J_b / T_s block-tridiagonal solve: the product's block-Thomas
sweep on synthetic block rows against the CPU restatement of block_tridiag_solve.h (oracle/block_thomas.py), including the
cases no shipped deck reaches: row exchanges in the PartialPivLU, a zero pivot column, exactly-zero right-hand sides."""
import os
import sys

import numpy as np
import pytest

import aiolos_b200 as A
from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
from block_thomas import dense_blocks, pack_stream, solve_block_tridiag, stream_dims  # noqa: E402


def make_stream(seed, ncols, cells, B, S, mode):
    """Random block rows with the structure of radiation.cpp:407-615.  mode: 'dominant' (no row exchange), 'pivoting' (small
    diagonals: the LU exchanges rows in most steps), 'zeros' (pivoting plus a decoupled all-zero T row with zero right-hand
    side in the last cell, as the reference's last ghost cell has, and exactly-zero right-hand sides)."""
    rng = np.random.default_rng(seed)
    o = stream_dims(B, S)
    n = o["n"]
    coef = np.zeros((ncols, cells, o["NC"]))
    for c in range(ncols):
        for i in range(cells):
            D = rng.uniform(2.0, 4.0, n) * (n if mode == "dominant" else 0.05)
            L = -rng.uniform(0.2, 1.0, B) if i > 0 else np.zeros(B)
            U = -rng.uniform(0.2, 1.0, B) if i + 1 < cells else np.zeros(B)
            Bv = rng.uniform(-1.0, 1.0, n)
            JT = rng.uniform(-1.0, 1.0, (B, S)); TJ = rng.uniform(-1.0, 1.0, (S, B))
            TT = rng.uniform(-0.5, 0.5, (S, S))
            if mode == "zeros":
                if i % 3 == 1:
                    Bv[:] = 0.0
                if i == cells - 1:                     # all-zero T rows with zero right-hand side
                    D[B:] = 0.0; Bv[B:] = 0.0; TJ[:] = 0.0; TT[:] = 0.0
            coef[c, i] = pack_stream(D, L, U, Bv, JT, TJ, TT, B, S)
    return coef


def oracle_solution(coef, B, S):
    out = []
    for c in range(coef.shape[0]):
        l, d, u, rhs = dense_blocks(coef[c], B, S)
        out.append(solve_block_tridiag(l, d, u, rhs))
    return np.array(out)


@pytest.fixture(scope="module")
def lib(product_lib):
    L = A.load_library(product_lib)
    assert L.aio_device_count() > 0, "no CUDA device visible"
    return product_lib


CASES = [(22, 1, "dominant"), (22, 1, "pivoting"), (22, 1, "zeros"), (1, 2, "pivoting"), (5, 3, "pivoting"), (12, 4, "zeros"),
         (7, 1, "dominant"), (15, 1, "pivoting")]


def check(lib, B, S, mode):
    coef = make_stream(1234 + B * 7 + S, ncols=3, cells=6, B=B, S=S, mode=mode)
    ref = oracle_solution(coef, B, S)
    scale = np.abs(ref).max(axis=2, keepdims=True) + 1e-300
    x = A.block_thomas(coef, B, S, lib=lib)
    assert np.isfinite(x[np.isfinite(ref)]).all()
    err = np.abs(x - ref) / scale
    # 'pivoting' blocks are not diagonally dominant: condition numbers of 1e2..1e4 per block, amplified along the sweep
    assert np.nanmax(err) < (1e-11 if mode == "dominant" else 1e-8), (B, S, mode, np.nanmax(err))
    return x


@pytest.mark.parametrize("B,S,mode", CASES)
def test_block_thomas_emulated(emu_lib, B, S, mode):
    check(emu_lib, B, S, mode)


@pytest.mark.gpu
@pytest.mark.parametrize("B,S,mode", CASES)
def test_block_thomas_gpu(lib, B, S, mode):
    check(lib, B, S, mode)
