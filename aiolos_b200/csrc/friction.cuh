// Implicit multi-species friction (FRICTION_SOLVER != 0), one thread per cell, one CTA per column.
//   fill_alpha_basis_arrays     reference source.cpp:519-529
//   compute_alpha_matrix        source.cpp:536-645   (constant / n-n / i-n / resonant H-H+ / i-i / Epstein dust)
//   compute_friction_numerical  source.cpp:439-512   (T = I - A dt + diag(dt A 1); dense LU; v update; frictional heating)
// The SxS system of a cell lives in the registers of the thread that owns the cell (fully unrolled for S <= 8, so every
// matrix index is a compile-time constant); partial pivoting follows Eigen's PartialPivLU as restated in
// oracle/shim/Eigen/Core: first maximum, zero pivot => no scaling, triangular solves skip exactly-zero unknowns.
#pragma once
#include "common.cuh"
#include "hydro.cuh"

namespace aio {

// In-place LU with partial pivoting + solve of M x = b for one cell.  SS = compile-time capacity, S = runtime size.
// Fully unrolled (matrix in registers) up to SS = 8; the SS = 16 instantiation keeps its loops rolled and the matrix in local
// memory -- capacity for the rare decks with 9..16 species, not speed.
template <int SS>
__device__ __forceinline__ void lu_solve(double (&M)[SS][SS], double (&x)[SS], int S) {
#pragma unroll (SS <= 8 ? SS : 1)
    for (int k = 0; k < SS; k++) {
        if (k < S) {
            int piv = k; double big = fabs(M[k][k]);
#pragma unroll (SS <= 8 ? SS : 1)
            for (int r = k + 1; r < SS; r++) if (r < S) { double v = fabs(M[r][k]); if (v > big) { big = v; piv = r; } }
            if (big != 0.0) {
#pragma unroll (SS <= 8 ? SS : 1)
                for (int r = k + 1; r < SS; r++) {
                    if (r == piv) {
#pragma unroll (SS <= 8 ? SS : 1)
                        for (int c = 0; c < SS; c++) { double t = M[k][c]; M[k][c] = M[r][c]; M[r][c] = t; }
                        double t = x[k]; x[k] = x[r]; x[r] = t;       // permutation applied to the rhs as it is found
                    }
                }
                const double d = M[k][k];
#pragma unroll (SS <= 8 ? SS : 1)
                for (int r = k + 1; r < SS; r++) if (r < S) M[r][k] /= d;
            }
#pragma unroll (SS <= 8 ? SS : 1)
            for (int r = k + 1; r < SS; r++) {
                if (r < S) {
                    const double l = M[r][k];
#pragma unroll (SS <= 8 ? SS : 1)
                    for (int c = k + 1; c < SS; c++) if (c < S) M[r][c] = d_fnma(l, M[k][c], M[r][c]);
                }
            }
        }
    }
    // NOTE: swapping x while factoring is equivalent to applying P afterwards (row k is final once step k is done).
#pragma unroll (SS <= 8 ? SS : 1)
    for (int i = 0; i < SS; i++) {
        if (i < S) {
            const double xi = x[i];
            if (xi != 0.0) {
#pragma unroll (SS <= 8 ? SS : 1)
                for (int r = i + 1; r < SS; r++) if (r < S) x[r] = d_fnma(xi, M[r][i], x[r]);
            }
        }
    }
#pragma unroll (SS <= 8 ? SS : 1)
    for (int i = SS - 1; i >= 0; i--) {
        if (i < S) {
            if (x[i] != 0.0) {
                x[i] /= M[i][i];
                const double xi = x[i];
#pragma unroll (SS <= 8 ? SS : 1)
                for (int r = 0; r < i; r++) x[r] = d_fnma(xi, M[r][i], x[r]);
            }
        }
    }
}

// source.cpp:536-645 for one (si,sj) pair
__device__ __forceinline__ double alpha_pair(const DevParams& p, int si, int sj, int S, const SpeciesConst& ci_, const SpeciesConst& cj_,
                                             double dens_i, double dens_j, double n_j, double T_i, double T_j, double tnow) {
    double alpha_local;
    if (p.collision_model == 'C') {
        alpha_local = p.alpha_collision;
        if (si > sj) alpha_local *= dens_j / dens_i;
        return alpha_local;
    }
    if (si == sj) return 0.;
    const double mi = ci_.mass, mj = cj_.mass;
    const double itm = p.inv_totmasses[si * S + sj];
    const double mumass = mi * mj * itm;
    const double mumass_amu = mumass / C_AMU;
    const double meanT = (mj * T_i + mi * T_j) * itm;
    if (ci_.is_dust_like || cj_.is_dust_like) {
        const double RHO_DUST = 1;
        double s0 = 0, s1 = 0;
        if (ci_.is_dust_like) s0 = pow(3 * mi / (4 * M_PI * RHO_DUST), 1 / 3.);
        if (cj_.is_dust_like) s1 = pow(3 * mj / (4 * M_PI * RHO_DUST), 1 / 3.);
        const double A = M_PI * (s0 + s1) * (s0 + s1);
        const double v_th = sqrt(8 * C_KB * meanT / (M_PI * mumass));
        return 4 / 3. * dens_j * v_th * A * itm;
    }
    const int ci = (int)(ci_.static_charge * 1.01), cj = (int)(cj_.static_charge * 1.01);
    const double qn = (fabs(ci_.static_charge) + fabs(cj_.static_charge));
    if (abs(ci) + abs(cj) == 0) {
        const double coll_b = 5.0e17 * sqrt(sqrt(meanT * meanT * meanT));
        alpha_local = C_KB * meanT * n_j * ci_.inv_mass / coll_b;
    } else if (abs(ci) == 0 || abs(cj) == 0) {
        if (ci_.mass_amu < 1.1 && cj_.mass_amu < 1.1) {
            alpha_local = 2.65e-10 * n_j * sqrt(meanT) * pow((1. - 0.083 * log10(meanT)), 2.);
        } else {
            alpha_local = 1 * 2.21 * 3.141592 * n_j * mj * itm;
            alpha_local *= sqrt(0.66 / mumass) * 1e-12 * qn * C_ELM_CHARGE;
            alpha_local *= p.resonant_pair[si * S + sj];
        }
    } else {
        alpha_local = p.alpha_collision_ions * 1.27 * ci_.static_charge * ci_.static_charge * cj_.static_charge * cj_.static_charge *
                      sqrt(mumass_amu) * C_AMU * ci_.inv_mass;
        alpha_local *= n_j / sqrt(meanT * meanT * meanT);
    }
    double alpha_temp = p.alpha_collision;
    if (tnow < p.coll_rampup_time) alpha_temp = p.alpha_collision * (p.init_coll_factor + (1. - p.init_coll_factor) * tnow / p.coll_rampup_time);
    alpha_local *= alpha_temp;
    return alpha_local;
}

// Friction solver 0 (source.cpp:249-420): closed-form implicit drag for two species, the reference's (self-declared
// experimental) determinant formula for three, nothing for more.  Restated expression by expression, including the dust
// branch's `species[0].is_dust_like` tested twice (source.cpp:283-286) and the `a(0,1)* + a(0,2)` of its v2a line (:377).
template <int SS>
__device__ __forceinline__ void friction_analytic(const DevParams& p, const int S, const double dt, const double* rho, const double* v,
                                                  const double* T, const double* nd, double* x, double* heat) {
    if (S == 2 && SS >= 2) {
        const SpeciesConst s0 = p.sp[0], s1c = p.sp[1];
        const double v1b = v[0], v2b = v[1];
        double alpha;
        if (p.collision_model == 'C') alpha = p.alpha_collision;
        else {
            const double m0 = s0.mass_amu * C_AMU, m1 = s1c.mass_amu * C_AMU;
            const double mumass = m0 * m1 / (m0 + m1);
            const double meanT = (m1 * T[0] + m0 * T[1]) / (m0 + m1);
            if (s0.is_dust_like || s1c.is_dust_like) {
                const double RHO_DUST = 1;
                double r0 = 0, r1 = 0;
                if (s0.is_dust_like) r0 = pow(3 * m0 / (4 * M_PI * RHO_DUST), 1 / 3.);
                if (s0.is_dust_like) r1 = pow(3 * m1 / (4 * M_PI * RHO_DUST), 1 / 3.);
                const double A = M_PI * (r0 + r1) * (r0 + r1);
                const double v_th = sqrt(8 * C_KB * meanT / (M_PI * mumass));
                alpha = 4 / 3. * rho[1] * v_th * A / (m0 + m1);
            } else {
                const double coll_b = 5.0e17 * pow(meanT, 0.75);
                alpha = C_KB * meanT * nd[1] / (m0 * coll_b);
                alpha *= p.alpha_collision;
            }
        }
        const double eps = rho[0] / rho[1];
        const double f1 = (1. + dt * alpha) / (1. + dt * alpha * (1. + eps));
        const double f2 = dt * alpha * eps / (1. + dt * alpha);
        const double v2a = (v2b + v1b * f2) * f1;
        const double v1a = v1b - (v2a - v2b) / eps;
        x[0] = v1a; x[1] = v2a;
        const double v_half = 0.5 * (v1a + v1b) - 0.5 * (v2a + v2b);
        const double v_end = v1a - v2a;
        heat[0] = dt * alpha * (s1c.mass_amu / (s0.mass_amu + s1c.mass_amu)) * v_half * v_end;
        heat[1] = dt * alpha * eps * (s0.mass_amu / (s0.mass_amu + s1c.mass_amu)) * v_half * v_end;
    } else if (S == 3 && SS >= 3) {
        double a[3][3];
        for (int si = 0; si < 3; si++)
            for (int sj = 0; sj < 3; sj++) {
                if (si == sj) a[si][sj] = 0.;
                else if (si > sj) a[si][sj] = p.fric_mask[si * 3 + sj] * p.alpha_collision;
                else a[si][sj] = p.fric_mask[si * 3 + sj] * p.alpha_collision * rho[sj] / rho[si];
            }
        const double v1b = v[0], v2b = v[1], v3b = v[2];
        double det = -(a[0][2] + dt * a[0][2] * a[1][0] + a[0][1] * a[1][2] + dt * a[0][2] * a[1][2]) * a[2][0];
        det += (-a[0][2] * a[1][0] - a[1][2] - dt * a[0][1] * a[1][2] - dt * a[0][2] * a[1][2]) * a[2][1];
        det += (-a[0][1] * a[1][0] + (1. + dt * (a[0][1] + a[0][2])) * (1. + dt * (a[1][0] + a[1][2]))) * (1. + dt * (a[2][0] + a[2][1]));
        double v1a = v1b * (-a[1][2] * a[2][1] + (1. + dt * (a[1][0] + a[1][2])) * (1. + dt * (a[2][0] + a[2][1])));
        v1a += v2b * (a[0][1] + dt * a[0][1] * a[2][0] + dt * a[0][1] * a[2][1] + a[0][2] * a[2][1]);
        v1a += v3b * (a[0][2] + dt * a[0][2] * a[1][0] + a[0][1] * a[1][2] + dt * a[0][2] * a[1][2]);
        double v2a = v1b * (a[1][0] + dt * a[1][0] * a[2][0] + a[1][2] * a[2][0] + dt * a[1][0] * a[2][1]);
        v2a += v2b * (-a[0][2] * a[2][0] + (1. + dt * (a[0][1] * +a[0][2])) * (1. + dt * (a[2][0] + a[2][1])));
        v2a += v3b * (a[0][2] * a[1][0] + a[1][2] + dt * a[0][1] * a[1][2] + dt * a[0][2] * a[1][2]);
        double v3a = v1b * (a[2][0] + dt * a[1][0] * a[2][0] + dt * a[1][2] * a[2][0] + a[1][0] * a[2][1]);
        v3a += v2b * (a[0][1] * a[2][0] + a[2][1] + dt * a[0][1] * a[2][1] + dt * a[0][2] * a[2][1]);
        v3a += v3b * (-a[0][1] * a[1][0] + (1. + dt * (a[0][1] + a[0][2])) * (1. + dt * (a[1][0] + a[1][2])));
        v1a /= det; v2a /= det; v3a /= det;
        x[0] = v1a; x[1] = v2a; x[2] = v3a;
        const double m0 = p.sp[0].mass_amu, m1 = p.sp[1].mass_amu, m2 = p.sp[2].mass_amu;
        const double d12 = (v1a - v2a) * (v1a - v2a), d13 = (v1a - v3a) * (v1a - v3a), d23 = (v2a - v3a) * (v2a - v3a);
        double h0 = 0, h1 = 0, h2 = 0;
        h0 += dt * a[0][1] * (m1 / (m0 + m1)) * d12; h0 += dt * a[0][2] * (m2 / (m0 + m2)) * d13;
        h1 += dt * a[1][0] * (m0 / (m1 + m0)) * d12; h1 += dt * a[1][2] * (m2 / (m1 + m2)) * d23;
        h2 += dt * a[2][0] * (m0 / (m2 + m0)) * d13; h2 += dt * a[2][1] * (m1 / (m2 + m1)) * d23;
        heat[0] = h0; heat[1] = h1; heat[2] = h2;
    }
}

template <int SS>
__device__ __forceinline__ void friction_phase(const DevParams& p, const int col, const double dt, const double tnow, TailAcc& acc) {
    const int tid = threadIdx.x, NT = blockDim.x;
    const int S = p.S, N2 = p.N2;

    for (int j = tid; j < N2; j += NT) {
        double rho[SS], v[SS], e[SS], T[SS], cs[SS], nd[SS], u2o[SS];
#pragma unroll (SS <= 8 ? SS : 1)
        for (int s = 0; s < SS; s++) {
            if (s < S) {
                const SpeciesConst sc = p.sp[s];
                const double a1 = p.u[idx_u(p, 0, col, s, j)], a2 = p.u[idx_u(p, 1, col, s, j)], a3 = p.u[idx_u(p, 2, col, s, j)];
                Prim3 q = cons_to_prim(a1, a2, a3, sc.gm1);           // compute_pressure, advection.cpp:204-205
                rho[s] = q.rho; v[s] = q.v; u2o[s] = a2;
                nd[s] = q.rho / sc.mass;
                e[s] = eint_of(q.p, q.rho, sc.gm1);
                cs[s] = cs_of(e[s], sc.gm1);
                T[s] = e[s] / sc.cv;
            }
        }
        double x[SS], heat[SS];
#pragma unroll (SS <= 8 ? SS : 1)
        for (int si = 0; si < SS; si++) { x[si] = (si < S) ? v[si] : 0.; heat[si] = 0.; }
        if (p.friction_solver == 0) {
            friction_analytic<SS>(p, S, dt, rho, v, T, nd, x, heat);
        } else {
        double A[SS][SS], M[SS][SS];
#pragma unroll (SS <= 8 ? SS : 1)
        for (int si = 0; si < SS; si++) {
#pragma unroll (SS <= 8 ? SS : 1)
            for (int sj = 0; sj < SS; sj++) {
                if (si < S && sj < S)
                    A[si][sj] = p.fric_mask[si * S + sj] *
                                alpha_pair(p, si, sj, S, p.sp[si], p.sp[sj], rho[si], rho[sj], nd[sj], T[si], T[sj], tnow);
                else A[si][sj] = 0;
            }
        }
#pragma unroll (SS <= 8 ? SS : 1)
        for (int si = 0; si < SS; si++) {
            double rowsum = 0;
#pragma unroll (SS <= 8 ? SS : 1)
            for (int sj = 0; sj < SS; sj++) {
                if (si < S && sj < S) { M[si][sj] = d_sub((si == sj) ? 1. : 0., d_mul(A[si][sj], dt)); rowsum = d_add(rowsum, A[si][sj]); }
                else M[si][sj] = (si == sj) ? 1. : 0.;
            }
            M[si][si] = d_add(M[si][si], d_mul(dt, rowsum));
        }
        lu_solve<SS>(M, x, S);
#pragma unroll (SS <= 8 ? SS : 1)
        for (int si = 0; si < SS; si++) {
            if (si < S) {
                double temp = 0;
#pragma unroll (SS <= 8 ? SS : 1)
                for (int sj = 0; sj < SS; sj++) {
                    if (sj < S) {
                        const double v_end = x[si] - x[sj];
                        const double v_half = 0.5 * (v_end + v[si] - v[sj]);
                        const double mj = p.sp[sj].mass_amu;
                        temp += dt * A[si][sj] * (mj / (mj + p.sp[si].mass_amu)) * v_half * v_end;
                    }
                }
                heat[si] = temp;
            }
        }
        }
#pragma unroll (SS <= 8 ? SS : 1)
        for (int si = 0; si < SS; si++) {
            if (si < S) {
                const SpeciesConst sc = p.sp[si];
                const double temp = heat[si];
                const double en = e[si] + temp;
                Prim3 q; q.rho = rho[si]; q.v = x[si]; q.p = en * rho[si] * sc.gm1;      // update_p_from_eint
                double n1, n2, n3;
                prim_to_cons(q, sc.gm1, n1, n2, n3);
                p.u[idx_u(p, 0, col, si, j)] = n1; p.u[idx_u(p, 1, col, si, j)] = n2; p.u[idx_u(p, 2, col, si, j)] = n3;
                if (j == 2) p.prim2v[(size_t)col * S + si] = x[si];
                if (p.last_kernel == 1) tail_cell(p, col, si, j, x[si], cs[si], en, T[si], acc);
                else { p.Tsave[idx_s(p, col, si, j)] = T[si]; p.cssave[idx_s(p, col, si, j)] = cs[si]; }
            }
        }
    }
}

#ifndef FRICTION_MIN_CTAS
#define FRICTION_MIN_CTAS 4            // 128 registers (some spills) and 4 CTAs per SM: 0.67 -> 0.52 ms for 4096 x 102 cells x 6 species (r2;
                                      // 1 / 3 / 4 / 5 / 6 CTAs: 0.67 / 0.57 / 0.52 / 0.63 / 0.61 ms)
#endif
template <int SS>
__global__ void __launch_bounds__(128, FRICTION_MIN_CTAS) k_friction(DevParams p) {
    __shared__ double red[32];
    const int col = blockIdx.x;
    const ColState& c = p.cs[col];
    if (!c.in_step) return;
    TailAcc acc; acc.minstep = 0; acc.maxde = 0; acc.bad = 0; acc.badval = 0;
    friction_phase<SS>(p, col, c.dt, c.t, acc);
    if (p.last_kernel == 1) tail_commit(p, col, acc, red);
}

}  // namespace aio
