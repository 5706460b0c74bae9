"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's block-tridiagonal solver for the coupled J_b / T_s system.

Follows /root/reference/block_tridiag_solve.h:57-130 (BlockTriDiagSolver::factor_matrix / solve) with Eigen's PartialPivLU
restated as the oracle build's stand-in does (oracle/shim/Eigen/Core: unblocked right-looking LU, first row of the largest
|a(r,k)| as pivot, a zero pivot column is not scaled, triangular solves skip exactly-zero unknowns).  Plain numpy / Python
loops: only for the small systems of tests/test_block_thomas.py.  Parity of this restatement: checked against the oracle
build itself through the c3 / c6 fixtures (same PartialPivLU) and by tests/test_shim_lu.py.
"""
import numpy as np


class PartialPivLU:
    def __init__(self, a):
        a = np.array(a, dtype=np.float64)
        n = a.shape[0]
        self.n, self.piv = n, np.arange(n)
        for k in range(n):
            col = np.abs(a[k:, k])
            p = k + int(np.argmax(col))                 # argmax returns the first maximum
            self.piv[k] = p
            if col[p - k] != 0.0:
                if p != k:
                    a[[k, p], :] = a[[p, k], :]
                a[k + 1:, k] = a[k + 1:, k] / a[k, k]
            for i in range(k + 1, n):
                a[i, k + 1:] -= a[i, k] * a[k, k + 1:]
        self.lu = a

    def solve(self, b):
        x = np.array(b, dtype=np.float64)
        if x.ndim == 1:
            return self.solve(x[:, None])[:, 0]
        n, lu = self.n, self.lu
        for c in range(x.shape[1]):
            for k in range(n):
                if self.piv[k] != k:
                    x[[k, self.piv[k]], c] = x[[self.piv[k], k], c]
            for i in range(n):
                if x[i, c] != 0.0:
                    x[i + 1:, c] -= x[i, c] * lu[i + 1:, i]
            for i in range(n - 1, -1, -1):
                if x[i, c] != 0.0:
                    x[i, c] /= lu[i, i]
                    x[:i, c] -= x[i, c] * lu[:i, i]
        return x


def solve_block_tridiag(l, d, u, rhs):
    """l, d, u: [cells][n][n]; rhs: [cells][n].  block_tridiag_solve.h: factor_matrix (:57-87) then solve (:89-130)."""
    cells, n = rhs.shape
    lus = [PartialPivLU(d[0])]
    for i in range(1, cells):
        lus.append(PartialPivLU(d[i] - l[i] @ lus[i - 1].solve(u[i - 1])))
    a = np.zeros_like(rhs)
    a[0] = lus[0].solve(rhs[0])
    for i in range(1, cells):
        a[i] = lus[i].solve(rhs[i] - l[i] @ a[i - 1])
    x = np.zeros_like(rhs)
    x[cells - 1] = a[cells - 1]
    for i in range(cells - 2, -1, -1):
        x[i] = a[i] - lus[i].solve(u[i] @ x[i + 1])
    return x


def stream_dims(bands_out, num_species):
    """Offsets of the per-cell coefficient stream the device sweep reads (aiolos_b200/csrc/rad_full2.h: full2_dims)."""
    n = bands_out + num_species
    R = 8 if n <= 8 else 16 if n <= 16 else 24
    o = {"n": n, "R": R, "D": 0, "L": R, "U": 2 * R, "B": 3 * R, "JT": 4 * R}
    o["TJ"] = o["JT"] + bands_out * num_species
    o["TT"] = o["TJ"] + num_species * bands_out
    nc = o["TT"] + num_species * num_species
    o["NC"] = nc + (nc & 1)
    return o


def dense_blocks(coef, bands_out, num_species):
    """Expand one column's coefficient stream [cells][NC] into the dense l, d, u blocks and right-hand sides the reference's
    update_fluxes_FLD would hand to BlockTriDiagSolver (radiation.cpp:407-615: l_i / u_i carry only the J-J diagonal)."""
    o = stream_dims(bands_out, num_species)
    cells, n, B, S = coef.shape[0], o["n"], bands_out, num_species
    l = np.zeros((cells, n, n)); d = np.zeros((cells, n, n)); u = np.zeros((cells, n, n)); rhs = np.zeros((cells, n))
    for i in range(cells):
        c = coef[i]
        for r in range(n):
            d[i, r, r] = c[o["D"] + r]
            rhs[i, r] = c[o["B"] + r]
        for r in range(B):
            l[i, r, r] = c[o["L"] + r]
            u[i, r, r] = c[o["U"] + r]
            for s in range(S):
                d[i, r, B + s] = c[o["JT"] + r * S + s]
                d[i, B + s, r] = c[o["TJ"] + s * B + r]
        for s in range(S):
            for t in range(S):
                if s != t:
                    d[i, B + s, B + t] = c[o["TT"] + s * S + t]
    return l, d, u, rhs
