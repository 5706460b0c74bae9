// Full FLD solver (PARI_USE_RADIATION 1): coupled J_b / T_s system, one WARP per column.
//   update_opacities (models C/P)      reference opacities.cpp:74-369
//   reset_dS / update_dS               radiation.cpp:15-348     lane = incoming band, sequential top-down optical depth
//   update_fluxes_FLD                  radiation.cpp:353-840    assembly of the block-tridiagonal system, n = Bout + S
//   BlockTriDiagSolver                 block_tridiag_solve.h:57-130   block Thomas, PartialPivLU per block
//   compute_planck_function_integral3  helpers.cpp:227-300
//
// Layout of the work: the off-diagonal blocks l_i, u_i only carry the diagonal of their J-J part (radiation.cpp:431-434),
// so  S_i = d_i - l_i S_{i-1}^{-1} u_{i-1}  needs  X_{i-1} = S_{i-1}^{-1} u_{i-1}  (n x Bout) and touches only J rows.
// Per cell the warp keeps the augmented matrix [S_i | u_i | rhs_i] in a warp-private shared-memory tile (lane = row during
// the elimination, lane = right-hand-side column during the back substitution; odd leading dimension => no bank
// conflicts either way) and runs the LU exactly as Eigen's PartialPivLU does (first-max pivot with physical row swaps,
// zero pivot => no scaling, back substitution skips exactly-zero unknowns -- the all-zero T rows of the last ghost cell
// rely on that).  X_i and S_i^{-1} rhs_i are written to HBM (n*(Bout+1) doubles per cell) for the back sweep
//   x_i = A_i - X_i x_{i+1}.
#pragma once
#include "common.cuh"
#include "hydro.cuh"
#include "rad_simple.cuh"

namespace aio {

// helpers.cpp:227-300
// log_spacing = log(p.lT_spacing), hoisted by the caller (the reference evaluates it in place: same value)
__device__ inline double planck_integral3(const DevParams& p, double lmin, double lmax, double temperature, double log_spacing) {
    if (p.Bout == 1) return 1;
    const double* P = p.planck_matrix;             // [num_plancks][2]: lambda*T, cumulative power
    double lT_min, lT_max;
    if (temperature < 2.71) { lT_min = lmin * 2.71; lT_max = lmax * 2.71; }
    else { lT_min = lmin * temperature; lT_max = lmax * temperature; }
    double m, power_min, power_max;
    const int np = p.num_plancks;
    if (lT_min < P[0]) {
        m = P[1] / P[0];
        power_min = P[1] + m * lT_min;
        if (lT_max < P[0]) return 1.;
    } else {
        const int imin = (int)(log(lT_min / P[0]) / log_spacing);
        m = (P[2 * (imin + 1) + 1] - P[2 * imin + 1]) / (P[2 * (imin + 1)] - P[2 * imin]);
        power_min = P[2 * imin + 1] + m * (lT_min - P[2 * imin]);
    }
    if (lmax * temperature > P[2 * (np - 1)]) {
        power_max = 1.;
        if (lT_min > P[2 * (np - 1)]) return 1;
    } else {
        const int imax = (int)(log(lT_max / P[0]) / log_spacing);
        m = (P[2 * (imax + 1) + 1] - P[2 * imax + 1]) / (P[2 * (imax + 1)] - P[2 * imax]);
        power_max = P[2 * imax + 1] + m * (lT_max - P[2 * imax]);
    }
    return power_max - power_min;
}

__device__ inline double planck_integral3(const DevParams& p, double lmin, double lmax, double temperature) {
    return planck_integral3(p, lmin, lmax, temperature, log(p.lT_spacing));
}

struct FullDims { int n, LD, LDX, per_warp; };
__host__ __device__ inline FullDims full_dims(int Bout, int S, int Bin) {
    FullDims d;
    d.n = Bout + S;
    d.LD = d.n + Bout + 1; if ((d.LD & 1) == 0) d.LD++;      // augmented row: S | u columns | rhs ; odd => conflict-free
    d.LDX = Bout + 1; if ((d.LDX & 1) == 0) d.LDX++;
    // tile, previous X (+A), x of the cell above, fac[S][Bout], contributions[S][Bin], scratch
    d.per_warp = d.n * d.LD + d.n * d.LDX + d.n + S * Bout + S * Bin + 2 * S + 8 + Bout + Bin;
    return d;
}

__device__ __forceinline__ double warp_max(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_or(int v) {
    for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NB>
__global__ void __launch_bounds__(64, (NB > 0) ? 1 : 8) k_rad_full(DevParams p) {
    AIO_DYN_SMEM(double, sm);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int col = blockIdx.x * wpb + wib;
    if (col >= p.ncol) return;
    ColState& cst = p.cs[col];
    if (!cst.in_step) return;
    const double dt = cst.dt, tnow = cst.t;
    const int N = p.N, N2 = p.N2, S = p.S, Bin = p.Bin, Bout = p.Bout;
    const FullDims fd = full_dims(Bout, S, Bin);
    const int n = fd.n, LD = fd.LD, LDX = fd.LDX, NA = n + Bout + 1;      // NA = used columns of the augmented tile
    double* A = sm + (size_t)wib * fd.per_warp;      // [n][LD]
    double* Xp = A + n * LD;                         // [n][LDX]: X_{i-1} (Bout columns) and A_{i-1} (column Bout)
    double* xs = Xp + n * LDX;                       // [n]
    double* facs = xs + n;                           // [S][Bout]
    double* contrib = facs + S * Bout;               // [S][Bin]
    double* misc = contrib + S * Bin;                // [2S + 8]
    double* dcore = misc + 2 * S + 8;                // [Bout] D_core of the surface cell (USE_PLANET_TEMPERATURE, radiation.cpp:464)
    double* ssurf = dcore + Bout;                    // [Bin]  S_band(num_ghosts-1, b): stellar flux reaching the surface
    const int KA = Bin + 2 * Bout;
    const double* kap = p.kappa_avg;
    double* Xg = p.Xstore + (size_t)col * N2 * n * (Bout + 1);

    // ---- stellar heating: lane = incoming band, top-down exactly as radiation.cpp:51-63,154-348 ----------------------------
    {
        const int b = lane;
        double F = 0, rad_above = 0, top_prev = 0;
        if (b < Bin) {
            F = p.solar_heating_final[(size_t)col * Bin + b];
            if (tnow < p.radiation_rampup_time)
                F = F * (p.init_radiation_factor + (1. - p.init_radiation_factor) * tnow / p.radiation_rampup_time);
            top_prev = p.dSband_top[(size_t)col * Bin + b];
        }
        for (int j = N + 1; j > 0; j--) {
            const double dxj = __ldg(p.dx + j);
            if (b < Bin) {
                double tot = 0;
                for (int s = 0; s < S; s++) tot += kapv(p, col, s, b, j) * p.u[idx_u(p, 0, col, s, j)];
                const double cell_tau = tot * dxj;
                const double rad_j = (j == N + 1) ? 0. : rad_above + cell_tau;
                double dSb;
                if (j < N + 1) {
                    const double Sb = F * exp(-rad_above);
                    if (j == p.ng - 1) ssurf[b] = Sb;
                    dSb = 0.25 * Sb / dxj * (1. - p.bond_albedo);
                    const double dtau_tot = -(rad_above - rad_j);
                    double dS_he = dSb;
                    if (dtau_tot > 1e-3) dSb *= (-expm1(-dtau_tot));
                    else dSb *= (-fastexpm1_2(-dtau_tot));
                    dS_he *= -fastexpm1_2(-0.);
                    dSb -= dS_he;
                    dSb = smax(dSb, 0.);
                } else {
                    dSb = top_prev;
                }
                if (j == N) p.dSband_top[(size_t)col * Bin + b] = dSb;
                for (int s = 0; s < S; s++) {
                    const double hsw = p.highenergy_switch[s * Bin + b];
                    const double frac = hsw * kapv(p, col, s, b, j) * p.u[idx_u(p, 0, col, s, j)] * dxj / cell_tau;
                    contrib[s * Bin + b] = hsw * dSb * frac;
                }
                rad_above = rad_j;
            }
            __syncwarp();
            if (lane < S) {
                double acc = 0.;                              // reset_dS
                for (int bb = 0; bb < Bin; bb++) { acc += contrib[lane * Bin + bb]; if (acc < 1e-50) acc = 0.; }
                p.dS[idx_s(p, col, lane, j)] = acc;
            }
            __syncwarp();
        }
        if (lane < S) p.dS[idx_s(p, col, lane, 0)] = 0.;
    }

    // ---- forward sweep over cells: assemble block row i, S_i, LU, X_i = S_i^{-1} u_i, A_i = S_i^{-1}(r_i - l_i A_{i-1}) --------
    const bool isJ = lane < Bout, isT = (lane >= Bout && lane < n);
    const int sT = lane - Bout;
    double D_prev = 0;                  // D_{i-1,b} of the J row this lane owns
    double to_i = 0, to_ip = 0;         // total_opacity(i,b), total_opacity(i+1,b)
    auto total_opacity = [&](int j, int b) -> double {
        if (j < 1 || j > N + 1) return 0.;
        double tot = 0;
        for (int s = 0; s < S; s++) tot = fmax(kapv(p, col, s, Bin + Bout + b, j) * p.u[idx_u(p, 0, col, s, j)], tot);
        return tot;
    };
    if (isJ) to_ip = total_opacity(0, lane);
    double Jcur = 0, Jnext = isJ ? p.J[idx_J(p, col, lane, 0)] : 0.;
    for (int i = 0; i < N2; i++) {
        // zero the tile
        for (int k = lane; k < n * LD; k += 32) A[k] = 0.;
        __syncwarp();
        // fac[s][b] = no_rad_trans * kappa_P * sigma T^3 / pi * planck fraction   (lane b, all species)
        if (isJ && i <= N) {
            for (int s = 0; s < S; s++) {
                const SpeciesConst sc = p.sp[s];
                const double u1 = p.u[idx_u(p, 0, col, s, i)], u2 = p.u[idx_u(p, 1, col, s, i)], u3 = p.u[idx_u(p, 2, col, s, i)];
                Prim3 q = cons_to_prim(u1, u2, u3, sc.gm1);
                const double Ts = p.Tsave ? p.Tsave[idx_s(p, col, s, i)] : eint_of(q.p, q.rho, sc.gm1) / sc.cv;
                facs[s * Bout + lane] = 1. * p.no_rad_trans * kapv(p, col, s, Bin + lane, i) * C_SIGMA_RAD * Ts * Ts * Ts / C_PI *
                                        planck_integral3(p, p.l_i_out[lane], p.l_i_out[lane + 1], Ts);
            }
        }
        __syncwarp();
        double l_rr = 0, u_rr = 0, rhs = 0;
        if (isJ) {
            const int b = lane;
            Jcur = Jnext;
            Jnext = (i + 1 < N2) ? p.J[idx_J(p, col, b, i + 1)] : 0.;
            to_i = to_ip;
            to_ip = total_opacity(i + 1, b);
            double d = 0;
            if (i >= 2 && i <= N + 1) { d = D_prev; l_rr = -D_prev; }
            if (i <= N) {
                const double tv = __ldg(p.vol + i) / (C_CLIGHT * dt);
                d += tv;
                rhs += tv * Jcur;
            }
            double D_i = 0;
            if (i >= 1 && i <= N) {
                const double dxc = (__ldg(p.x_i12 + i + 1) - __ldg(p.x_i12 + i));
                const double rhokr = sqrt(to_i * to_ip);
                const double tau_inv = 1. / (dxc * rhokr);
                const double R = p.xi_rad * tau_inv * fabs(Jnext - Jcur) / (Jcur + 1e-300);
                D_i = 1. * tau_inv * 1. * __ldg(p.surf + i) * flux_limiter(R);
                u_rr = -D_i; d += D_i;
            }
            D_prev = D_i;
            if (i < p.ng) { l_rr = 0; u_rr = -d; rhs = 0; }
            if (i == p.ng - 1 && p.use_planetary_temperature) {      // surface cell radiates at T_planet (radiation.cpp:451-472)
                const double Tp = cst.T_planet;
                double Ssrc = C_SIGMA_RAD * pow(Tp, 4.) * planck_integral3(p, p.l_i_out[b], p.l_i_out[b + 1], fabs(Tp));
                if (Tp < 0) Ssrc *= -1;
                const double dxc = (__ldg(p.x_i12 + i + 1) - __ldg(p.x_i12 + i));
                double rhokr = smax(2. * (to_i * to_ip) / (to_i + to_ip), 4. / 3. / dxc);
                rhokr = smin(0.5 * (to_i + to_ip), rhokr);
                const double tau_inv = 1 / (dxc * rhokr);
                const double R = p.xi_rad * tau_inv * fabs(Jnext - Jcur) / (Jnext + 1e-300);
                const double Dc = flux_limiter(R) * tau_inv;
                dcore[b] = Dc;
                const double Chi = 0.25;
                const double sj = __ldg(p.surf + i);
                l_rr = 0; d = 0.5 * (Dc + Chi) * sj; u_rr = -0.5 * (Dc - Chi) * sj; rhs = sj * Ssrc / (4 * C_PI);
            }
            if (p.closed_radiative_boundaries) {
                if (i >= N2 - p.ng) { l_rr = -d; u_rr = 0; rhs = 0; }
            } else if (i == N + 1) {
                double f = __ldg(p.x_i12 + N) / __ldg(p.x_i12 + N + 1);
                if (p.geometry == 0) f = 1;
                else if (p.geometry == 2) f *= f;
                l_rr = -f * d; u_rr = 0; rhs = 0;
            }
            if (i <= N && !(i < p.ng || i >= N + 2 - p.ng)) {
                const double vj = __ldg(p.vol + i);
                for (int s = 0; s < S; s++) {
                    const SpeciesConst sc = p.sp[s];
                    const double u1 = p.u[idx_u(p, 0, col, s, i)], u2 = p.u[idx_u(p, 1, col, s, i)], u3 = p.u[idx_u(p, 2, col, s, i)];
                    Prim3 q = cons_to_prim(u1, u2, u3, sc.gm1);
                    const double Ts = p.Tsave ? p.Tsave[idx_s(p, col, s, i)] : eint_of(q.p, q.rho, sc.gm1) / sc.cv;
                    const double fac = facs[s * Bout + b];
                    d += vj * q.rho * kapv(p, col, s, Bin + b, i) * p.no_rad_trans;
                    A[b * LD + Bout + s] = -4 * vj * q.rho * fac;
                    rhs -= 3 * vj * q.rho * fac * Ts;
                }
            }
            A[b * LD + b] = d;
        } else if (isT && i <= N) {
            const int s = sT;
            const SpeciesConst sc = p.sp[s];
            const double u1 = p.u[idx_u(p, 0, col, s, i)], u2 = p.u[idx_u(p, 1, col, s, i)], u3 = p.u[idx_u(p, 2, col, s, i)];
            Prim3 q = cons_to_prim(u1, u2, u3, sc.gm1);
            const double Ts = p.Tsave ? p.Tsave[idx_s(p, col, s, i)] : eint_of(q.p, q.rho, sc.gm1) / sc.cv;
            double d = 1 / dt;
            rhs = Ts / dt;
            rhs += (p.dS[idx_s(p, col, s, i)] + 0.) / u1 / sc.cv;
            for (int b = 0; b < Bout; b++) {
                const double fac = facs[s * Bout + b];
                d += 16 * C_PI * fac / sc.cv;
                A[lane * LD + b] = -4 * C_PI * kapv(p, col, s, Bin + b, i) * p.no_rad_trans / sc.cv;
                rhs += 12 * C_PI * fac * Ts / sc.cv;
            }
            A[lane * LD + lane] = d;
            // collisional T_i - T_j coupling (radiation.cpp:589-615, source.cpp:656-674)
            if (p.use_collisional_heating && S > 1) {
                double tot0 = 0;
                if (i >= 1) for (int ss = 0; ss < S; ss++) tot0 = fmax(kapv(p, col, ss, Bin + Bout + 0, i) * p.u[idx_u(p, 0, col, ss, i)], tot0);
                const double tau = tot0 * (__ldg(p.x_i12 + i + 1) - __ldg(p.x_i12 + i));
                const double damp = (tau < 1e3) ? 1. : 1.e-4;
                double diag_sum = 0, self = 0;
                for (int sj = 0; sj < S; sj++) {
                    const SpeciesConst cj = p.sp[sj];
                    const double w1 = p.u[idx_u(p, 0, col, sj, i)], w2 = p.u[idx_u(p, 1, col, sj, i)], w3 = p.u[idx_u(p, 2, col, sj, i)];
                    Prim3 qj = cons_to_prim(w1, w2, w3, cj.gm1);
                    const double Tj = p.Tsave ? p.Tsave[idx_s(p, col, sj, i)] : eint_of(qj.p, qj.rho, cj.gm1) / cj.cv;
                    double fc = p.fric_mask[s * S + sj] * alpha_pair(p, s, sj, S, sc, cj, u1, w1, w1 / cj.mass, Ts, Tj, tnow);
                    fc *= 3 * C_KB / (sc.cv * (sc.mass + cj.mass));
                    diag_sum += fc;
                    if (sj == s) self = fc;
                    else A[lane * LD + Bout + sj] -= fc * damp;
                }
                A[lane * LD + lane] -= (self - diag_sum) * damp;
            }
        }
        // S_i = d_i - l_i X_{i-1} ;  rhs_i - l_i A_{i-1}   (only J rows have l entries)
        if (isJ && i > 0 && l_rr != 0.) {
            for (int c = 0; c < Bout; c++) A[lane * LD + c] -= l_rr * Xp[lane * LDX + c];
            rhs -= l_rr * Xp[lane * LDX + Bout];
        }
        if (lane < n) {
            if (isJ) A[lane * LD + n + lane] = u_rr;
            A[lane * LD + n + Bout] = rhs;
        }
        __syncwarp();

        if constexpr (NB > 0) {
            // ---- register-resident variant: lane r keeps row r of [S | u | rhs] in registers (compile-time indices, fully unrolled);
            //      the pivot row travels by shuffles, the back substitution runs lane = right-hand-side column out of the tile ----------
            double a[NB], z[NB];
#pragma unroll
            for (int c = 0; c < NB; c++) {
                a[c] = (lane < n && c < n) ? A[lane * LD + c] : ((c == lane) ? 1. : 0.);
                z[c] = (lane < n && c <= Bout) ? A[lane * LD + n + c] : 0.;
            }
            // column c of the u block is zero except in row pos[c] until that row becomes the pivot row (uniform bookkeeping)
            int pos[NB]; unsigned active = 0;
#pragma unroll
            for (int c = 0; c < NB; c++) pos[c] = c;
#pragma unroll
            for (int k = 0; k < NB; k++) {
                if (k < n) {
                    const double v = (lane >= k && lane < n) ? fabs(a[k]) : -1.;
                    const double big = warp_max(v);
                    const unsigned m = __ballot_sync(0xffffffffu, v == big);
                    const int piv = __ffs(m) - 1;
                    if (big != 0.0) {
                        if (piv != k) {
                            const int src = (lane == k) ? piv : (lane == piv) ? k : lane;
#pragma unroll
                            for (int c = 0; c < NB; c++) { a[c] = __shfl_sync(0xffffffffu, a[c], src); z[c] = __shfl_sync(0xffffffffu, z[c], src); }
#pragma unroll
                            for (int c = 0; c < NB; c++) { if (pos[c] == k) pos[c] = piv; else if (pos[c] == piv) pos[c] = k; }
                        }
                        const double pv = __shfl_sync(0xffffffffu, a[k], k);
                        if (lane > k && lane < n) a[k] /= pv;
                    }
                    const bool below = (lane > k && lane < n);
                    const double l = below ? a[k] : 0.;
#pragma unroll
                    for (int c = k + 1; c < NB; c++) {
                        if (c < n) {
                            const double rk = __shfl_sync(0xffffffffu, a[c], k);
                            if (below && l != 0.) a[c] = d_fnma(l, rk, a[c]);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < NB; c++) {
                        if (c <= Bout) {
                            const bool is_rhs = (c == Bout);
                            if (!is_rhs && !((active >> c) & 1u) && pos[c] == k) active |= (1u << c);
                            if (is_rhs || ((active >> c) & 1u)) {
                                const double zk = __shfl_sync(0xffffffffu, z[c], k);
                                if (zk != 0. && below && l != 0.) z[c] = d_fnma(l, zk, z[c]);
                            }
                        }
                    }
                }
            }
            // rows back into the tile (U in the upper triangle), then lane = right-hand-side column
#pragma unroll
            for (int c = 0; c < NB; c++) {
                if (lane < n && c < n) A[lane * LD + c] = a[c];
                if (lane < n && c <= Bout) A[lane * LD + n + c] = z[c];
            }
            __syncwarp();
            if (lane < Bout + 1) {
                double y[NB];
#pragma unroll
                for (int k = 0; k < NB; k++) y[k] = (k < n) ? A[k * LD + n + lane] : 0.;
#pragma unroll
                for (int k = NB - 1; k >= 0; k--) {
                    if (k < n && y[k] != 0.) {
                        y[k] /= A[k * LD + k];
#pragma unroll
                        for (int r = 0; r < k; r++) y[r] = d_fnma(y[k], A[r * LD + k], y[r]);
                    }
                }
#pragma unroll
                for (int k = 0; k < NB; k++) if (k < n) A[k * LD + n + lane] = y[k];
            }
            __syncwarp();
        } else {
            // ---- PartialPivLU of the n x n block, the right-hand-side columns ride along (tile row = [S | u | rhs], so one
            //      contiguous sweep over columns k+1..NA-1 updates both; multiplying the structural zeros of the u block is exact) -----
            for (int k = 0; k < n; k++) {
                double v = (lane >= k && lane < n) ? fabs(A[lane * LD + k]) : -1.;
                const double big = warp_max(v);
                // first row (smallest index) holding the maximum
                unsigned m = __ballot_sync(0xffffffffu, v == big);
                const int piv = __ffs(m) - 1;
                if (big != 0.0) {
                    if (piv != k) {
                        for (int c = lane; c < NA; c += 32) { const double t = A[k * LD + c]; A[k * LD + c] = A[piv * LD + c]; A[piv * LD + c] = t; }
                        __syncwarp();
                    }
                    if (lane > k && lane < n) A[lane * LD + k] /= A[k * LD + k];
                }
                if (lane > k && lane < n) {
                    const double l = A[lane * LD + k];
                    if (l != 0.) {
                        const double* __restrict__ rk = A + k * LD;
                        double* __restrict__ rl = A + lane * LD;
#pragma unroll 4
                        for (int c = k + 1; c < NA; c++) rl[c] = d_fnma(l, rk[c], rl[c]);
                    }
                }
                __syncwarp();
            }
            // ---- back substitution, lane = right-hand-side column (Bout columns of u_i and the rhs) ---------------------------------
            if (lane < Bout + 1) {
                double* __restrict__ col_ = A + n + lane;
                for (int k = n - 1; k >= 0; k--) {
                    double z = col_[k * LD];
                    if (z != 0.) {
                        z /= A[k * LD + k];
                        col_[k * LD] = z;
                        const double* uk = A + k;
#pragma unroll 4
                        for (int r = 0; r < k; r++) col_[r * LD] = d_fnma(z, uk[r * LD], col_[r * LD]);
                    }
                }
            }
            __syncwarp();
        }
        // ---- keep X_i / A_i for the next cell and for the back sweep -------------------------------------------------------------
        double* Xi = Xg + (size_t)i * n * (Bout + 1);
        if (lane <= Bout)
            for (int r = 0; r < n; r++) {
                const double v = A[r * LD + n + lane];
                Xp[r * LDX + lane] = v;
                Xi[r * (Bout + 1) + lane] = v;
            }
        __syncwarp();
    }

    // ---- back sweep  x_i = A_i - X_i x_{i+1},  store J, clamp T, e <- T, p <- e, cons <- prim, step tail ---------------------------
    TailAcc acc; acc.minstep = 0; acc.maxde = 0; acc.bad = 0; acc.badval = 0;
    int negJ = 0; double negJval = 0;
    double Jtop = 0;                    // Jrad_FLD(num_cells, b) of this lane's band (advection.cpp:320-322)
    for (int i = N2 - 1; i >= 0; i--) {
        const double* Xi = Xg + (size_t)i * n * (Bout + 1);
        double x = 0;
        if (lane < n) {
            x = Xi[lane * (Bout + 1) + Bout];
            if (i < N2 - 1) {
                double t = 0;
                for (int c = 0; c < Bout; c++) t += Xi[lane * (Bout + 1) + c] * xs[c];
                x = x - t;
            }
        }
        __syncwarp();
        if (lane < n) xs[lane] = x;
        __syncwarp();
        if (isJ) {
            double Jv = x;
            if (p.couple_J_into_T && i <= N - 1 && Jv < 0) {           // advection.cpp:308-338
                if (Jv > -1e-10) Jv *= -1.;
                else if (!(i < N && p.use_planetary_temperature && Jtop > 0)) { negJ = 1; negJval = Jv; }
            }
            if (i == N) Jtop = Jv;
            p.J[idx_J(p, col, lane, i)] = Jv;
        } else if (isT) {
            const int s = sT;
            const SpeciesConst sc = p.sp[s];
            double T = x;
            if (T < p.temperature_floor) T = p.temperature_floor;
            if (T > p.max_temperature) T = p.max_temperature;
            const double u1 = p.u[idx_u(p, 0, col, s, i)], u2 = p.u[idx_u(p, 1, col, s, i)], u3 = p.u[idx_u(p, 2, col, s, i)];
            Prim3 q = cons_to_prim(u1, u2, u3, sc.gm1);
            const double e_old = eint_of(q.p, q.rho, sc.gm1);
            const double cs_old = p.cssave ? p.cssave[idx_s(p, col, s, i)] : cs_of(e_old, sc.gm1);
            const double e = T * sc.cv;
            q.p = e * q.rho * sc.gm1;
            double n1, n2, n3;
            prim_to_cons(q, sc.gm1, n1, n2, n3);
            p.u[idx_u(p, 0, col, s, i)] = n1; p.u[idx_u(p, 1, col, s, i)] = n2; p.u[idx_u(p, 2, col, s, i)] = n3;
            tail_cell(p, col, s, i, q.v, cs_old, e, T, acc);
        }
    }
    // slab model of the planet's surface temperature (radiation.cpp:812-828), bands summed in order by one lane
    __syncwarp();
    if (p.use_planetary_temperature && lane == 0) {
        double F = 0, S_surf = 0;
        for (int b = 0; b < Bout; b++) {
            F -= 4 * C_PI * dcore[b] * (p.J[idx_J(p, col, b, p.ng)] - p.J[idx_J(p, col, b, p.ng - 1)]);
            S_surf += 0.25 * ssurf[b];
        }
        const double S_tot = S_surf - F + C_SIGMA_RAD * (p.T_int * p.T_int) * (p.T_int * p.T_int);
        double Tp = cst.T_planet;
        if (p.core_cv > 0) { const double denom = 4 * C_SIGMA_RAD * Tp * Tp * Tp + p.core_cv / dt; Tp += S_tot / denom; }
        else Tp = pow(pow(Tp, 4) + S_tot / C_SIGMA_RAD, 0.25);
        cst.T_planet = Tp;
    }
    const double ms = warp_max(acc.minstep), md = warp_max(acc.maxde);
    const int bad = warp_or(acc.bad), nj = warp_or(negJ);
    const double njv = -warp_max(negJ ? -negJval : -1e308);
    if (lane == 0) {
        cst.minstep = ms; cst.maxde = md;
        cst.steps += 1;
        if (bad) { cst.status = COL_NEG_T; cst.t = 1.1 * cst.t_max; }
        else if (nj) { cst.status = COL_NEG_J; cst.crash_value = njv; cst.t = 1.1 * cst.t_max; }
    }
}

}  // namespace aio
