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
    """Geometry of the per-cell coefficient stream the device sweep reads (aiolos_b200/csrc/rad_full2.h, Full2Dims):
    bands_out band records of W doubles [D, L, U, B, JT(s)..., TJ(s)..., pad], then num_species species records of V doubles
    [D, B, TT(t)..., pad]; W and V are multiples of 4."""
    n = bands_out + num_species
    W = (4 + 2 * num_species + 3) & ~3
    V = (2 + num_species + 3) & ~3
    return {"n": n, "W": W, "V": V, "T": bands_out * W, "NC": bands_out * W + num_species * V}


def pack_stream(D, L, U, Bv, JT, TJ, TT, bands_out, num_species):
    """One cell's coefficients (D, Bv: [n]; L, U: [bands_out]; JT: [bands_out][S]; TJ: [S][bands_out]; TT: [S][S]) as the stream."""
    o = stream_dims(bands_out, num_species)
    B, S, W, V = bands_out, num_species, o["W"], o["V"]
    c = np.zeros(o["NC"])
    for b in range(B):
        c[b * W:b * W + 4] = [D[b], L[b], U[b], Bv[b]]
        c[b * W + 4:b * W + 4 + S] = JT[b]
        c[b * W + 4 + S:b * W + 4 + 2 * S] = TJ[:, b]
    for s in range(S):
        r = o["T"] + s * V
        c[r], c[r + 1] = D[B + s], Bv[B + s]
        c[r + 2:r + 2 + S] = TT[s]
        c[r + 2 + s] = 0.0
    return c


def dense_blocks(coef, bands_out, num_species):
    """Expand one column's coefficient stream [cells][NC] into the dense l, d, u blocks and right-hand sides the reference's
    update_fluxes_FLD would hand to BlockTriDiagSolver (radiation.cpp:407-615: l_i / u_i carry only the J-J diagonal)."""
    o = stream_dims(bands_out, num_species)
    cells, n, B, S, W, V = coef.shape[0], o["n"], bands_out, num_species, o["W"], o["V"]
    l = np.zeros((cells, n, n)); d = np.zeros((cells, n, n)); u = np.zeros((cells, n, n)); rhs = np.zeros((cells, n))
    for i in range(cells):
        c = coef[i]
        for b in range(B):
            rec = c[b * W:(b + 1) * W]
            d[i, b, b], l[i, b, b], u[i, b, b], rhs[i, b] = rec[0], rec[1], rec[2], rec[3]
            for s in range(S):
                d[i, b, B + s] = rec[4 + s]
                d[i, B + s, b] = rec[4 + S + s]
        for s in range(S):
            rec = c[o["T"] + s * V:o["T"] + (s + 1) * V]
            d[i, B + s, B + s], rhs[i, B + s] = rec[0], rec[1]
            for t in range(S):
                if s != t:
                    d[i, B + s, B + t] = rec[2 + t]
    return l, d, u, rhs
