"""The oracle build links the unmodified reference against a stand-in for Eigen (oracle/shim/Eigen/Core: Eigen is not in the
image).  These tests pin the stand-in's PartialPivLU -- the one piece of it whose rounding order the parity fixtures depend
on -- on known answers, on the rules the reference relies on (first-maximum pivot, a zero pivot column is not scaled,
substitutions skip exactly-zero unknowns), and against the numpy restatement used by tests/test_block_thomas.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
from block_thomas import PartialPivLU  # noqa: E402

SRC = r'''
#include <Eigen/Dense>
#include <cstdio>
#include <cstdlib>
// usage: shim_lu n ncols < numbers (row-major A, then row-major B); prints x = A^-1 B with %.17g
int main(int argc, char** argv) {
    int n = atoi(argv[1]), m = atoi(argv[2]);
    Eigen::MatrixXd a(n, n), b(n, m);
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) if (scanf("%lf", &a(i, j)) != 1) return 2;
    for (int i = 0; i < n; i++) for (int j = 0; j < m; j++) if (scanf("%lf", &b(i, j)) != 1) return 2;
    Eigen::PartialPivLU<Eigen::MatrixXd> lu(a);
    Eigen::MatrixXd x = lu.solve(b);
    for (int i = 0; i < n; i++) { for (int j = 0; j < m; j++) printf("%.17g ", x(i, j)); printf("\n"); }
    return 0;
}
'''


@pytest.fixture(scope="module")
def shim_lu(tmp_path_factory):
    d = tmp_path_factory.mktemp("shim")
    src, exe = d / "shim_lu.cpp", d / "shim_lu"
    src.write_text(SRC)
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "oracle", "shim"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]

    def run(a, b):
        a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
        if b.ndim == 1:
            b = b[:, None]
        txt = " ".join(repr(float(v)) for v in list(a.ravel()) + list(b.ravel()))
        out = subprocess.run([str(exe), str(a.shape[0]), str(b.shape[1])], input=txt, capture_output=True, text=True, check=True).stdout
        return np.array([[float(t) for t in ln.split()] for ln in out.strip().splitlines()])
    return run


def test_known_answers(shim_lu):
    # 2x2 with a row exchange: pivot is row 1
    x = shim_lu([[1., 2.], [4., 3.]], [5., 10.])
    assert np.allclose(x[:, 0], [1., 2.], rtol=0, atol=1e-15)
    # 3x3, exactly representable solution
    a = [[2., 1., 1.], [4., -6., 0.], [-2., 7., 2.]]
    assert np.array_equal(shim_lu(a, [5., -2., 9.])[:, 0], [1., 1., 2.])
    # 4x4 permutation-like matrix: every step exchanges rows
    a = [[0., 0., 0., 2.], [0., 0., 4., 0.], [0., 8., 0., 0.], [16., 0., 0., 0.]]
    assert np.array_equal(shim_lu(a, [2., 4., 8., 16.])[:, 0], [1., 1., 1., 1.])


def test_first_maximum_pivot_rule(shim_lu):
    # rows 0 and 2 tie in |a(r,0)|: Eigen takes the FIRST maximum, i.e. no exchange; the result differs in the last bits from
    # the other choice, so equality with the numpy restatement (np.argmax = first maximum) pins the rule
    a = np.array([[3., 1. / 3., 0.7], [1., 2., 0.1], [-3., 0.3, 1. / 7.]])
    b = np.array([0.1, 0.2, 0.3])
    assert np.array_equal(shim_lu(a, b)[:, 0], PartialPivLU(a).solve(b))


def test_zero_pivot_and_zero_rhs_rules(shim_lu):
    # a zero column: nothing is scaled, no 0/0; the decoupled zero row with zero right-hand side yields exactly 0
    a = np.array([[2., 1., 0.], [1., 3., 0.], [0., 0., 0.]])
    x = shim_lu(a, [3., 4., 0.])[:, 0]
    assert np.isfinite(x).all() and x[2] == 0.0
    assert np.allclose(x[:2], np.linalg.solve(a[:2, :2], [3., 4.]), rtol=1e-15)
    # all-zero right-hand side: substitutions are skipped altogether
    assert np.array_equal(shim_lu(a, [0., 0., 0.])[:, 0], [0., 0., 0.])


@pytest.mark.parametrize("n", [1, 2, 3, 4, 9, 23])
def test_residual_and_restatement(shim_lu, n):
    rng = np.random.default_rng(n)
    a = rng.uniform(-1, 1, (n, n)) + np.diag(rng.uniform(-0.2, 0.2, n))        # not diagonally dominant: rows are exchanged
    b = rng.uniform(-1, 1, (n, 3))
    x = shim_lu(a, b)
    assert np.abs(a @ x - b).max() <= 1e-13 * max(1.0, np.abs(x).max()) * n
    # the numpy restatement performs the same operations in the same order (no FMA contraction on either side)
    assert np.array_equal(x, PartialPivLU(a).solve(b))
