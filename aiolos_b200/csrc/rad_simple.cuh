// Fused radiation kernel for PARI_USE_RADIATION 2 (one outgoing band): one CTA advances one column through
//   update_opacities (models C/P)            reference opacities.cpp:74-369
//   reset_dS / update_dS                      radiation.cpp:15-348   (top-down optical depth, stellar heating)
//   update_fluxes_FLD_simple                  radiation_simple.cpp:20-559 (scalar tridiagonal for J, explicit T_s)
//   step tail                                 helpers.cpp:28-71, advection.cpp:88-91,282-346
// The per-column recurrences are re-expressed for a CTA:
//   * reverse cumulative optical depth  -> blocked scan (thread-local serial sums + warp-shuffle scan of the partials)
//   * scalar block-Thomas (block_tridiag_solve.h:57-130 with 1x1 blocks) -> parallel cyclic reduction in shared
//     memory down to stride NT, then a register Thomas over the <=CPT rows a thread owns.
// Both change the association order of the sums/eliminations (round-off level only; parity tolerance in tests/).
#pragma once
#include "common.cuh"
#include "hydro.cuh"
#include "friction.cuh"
#include "opacity.cuh"

namespace aio {

__device__ __forceinline__ double flux_limiter(double R) {        // radiation_simple.cpp:25-30
    if (R <= 2) return 2 / (3 + sqrt(9 + 10 * R * R));
    else return 10 / (10 * R + 9 + sqrt(81 + 180 * R));
}
__device__ __forceinline__ double fastexpm1_2(double x) { return x + 0.5 * x * x; }   // aiolos.h:162-164

// Inclusive scan over q = 0..n-1 of val(q) in shared array `a` (in place), blocked over the CTA.
__device__ inline void block_inclusive_scan(double* a, int n, double* red) {
    const int tid = threadIdx.x, NT = blockDim.x;
    const int C = (n + NT - 1) / NT;
    const int q0 = tid * C, q1 = min(q0 + C, n);
    double loc = 0;
    for (int q = q0; q < q1; q++) { loc += a[q]; a[q] = loc; }
    // exclusive scan of `loc` over threads
    double v = loc;
    const int lane = tid & 31, w = tid >> 5, nw = (NT + 31) >> 5;
    for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    __syncthreads();
    if (lane == 31) red[w] = v;
    __syncthreads();
    double woff = 0;
    for (int k = 0; k < w; k++) woff += red[k];
    (void)nw;
    const double off = woff + (v - loc);
    for (int q = q0; q < q1; q++) a[q] += off;
    __syncthreads();
}

// Solve the tridiagonal system (A sub, B diag, C super, D rhs; n rows) in shared memory; solution left in D.
// PCR with strides 1,2,4,... while stride < NT, then each thread finishes its rows {t, t+NT, ...} serially.
template <int CPT>
__device__ inline void tridiag_pcr(double* A, double* B, double* C, double* D, int n) {
    const int tid = threadIdx.x, NT = blockDim.x;
    double a[CPT], b[CPT], c[CPT], d[CPT];
    for (int s = 1; s < NT && s < n; s <<= 1) {
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int i = tid + k * NT;
            if (i < n) {
                const int im = i - s, ip = i + s;
                double ai = A[i], bi = B[i], ci = C[i], di = D[i];
                double na = 0, nc = 0;
                if (im >= 0) { const double k1 = ai / B[im]; na = -A[im] * k1; bi -= C[im] * k1; di -= D[im] * k1; }
                if (ip < n) { const double k2 = ci / B[ip]; nc = -C[ip] * k2; bi -= A[ip] * k2; di -= D[ip] * k2; }
                a[k] = na; b[k] = bi; c[k] = nc; d[k] = di;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int i = tid + k * NT;
            if (i < n) { A[i] = a[k]; B[i] = b[k]; C[i] = c[k]; D[i] = d[k]; }
        }
        __syncthreads();
    }
    // rows t, t+NT, ... now form an independent tridiagonal system (coupling stride NT)
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        const int i = tid + k * NT;
        if (i < n) { a[k] = A[i]; b[k] = B[i]; c[k] = C[i]; d[k] = D[i]; }
    }
#pragma unroll
    for (int k = 1; k < CPT; k++) {
        const int i = tid + k * NT;
        if (i < n) { const double m = a[k] / b[k - 1]; b[k] -= m * c[k - 1]; d[k] -= m * d[k - 1]; }
    }
#pragma unroll
    for (int k = CPT - 1; k >= 0; k--) {
        const int i = tid + k * NT;
        if (i < n) {
            double x = d[k];
            if (k + 1 < CPT && (i + NT) < n) x -= c[k] * d[k + 1];
            d[k] = x / b[k];
            D[i] = d[k];
        }
    }
    __syncthreads();
}

template <int CPT>
__device__ __forceinline__ void rad_simple_phase(const DevParams& p, const int col, double* sm, const double ddt, const double tnow) {
    const int tid = threadIdx.x, NT = blockDim.x;
    const int N = p.N, N2 = p.N2, S = p.S, Bin = p.Bin;

    double* RHO = sm;                  // [S][N2]
    double* TT = RHO + S * N2;         // [S][N2] temperature
    double* DS = TT + S * N2;          // [S][N2]
    double* E1 = DS + S * N2;          // [S][N2] eta1
    double* DEN = E1 + S * N2;         // [S][N2]
    double* TO = DEN + S * N2;         // total_opacity (out band 0)
    double* JJ = TO + N2;
    double* TAU = JJ + N2;             // scan workspace, one in-band at a time
    double* A = TAU + N2; double* B = A + N2; double* C = B + N2; double* D = C + N2;
    double* red = D + N2;              // 32

    // ---- prims from the post-hydro state (compute_pressure, advection.cpp:204-205) ---------------------------------
    for (int s = 0; s < S; s++) {
        const SpeciesConst sc = p.sp[s];
        for (int j = tid; j < N2; j += NT) {
            Prim3 q = cons_to_prim(p.u[idx_u(p, 0, col, s, j)], p.u[idx_u(p, 1, col, s, j)], p.u[idx_u(p, 2, col, s, j)], sc.gm1);
            RHO[s * N2 + j] = q.rho;
            // prim.temperature is what the post-hydro compute_pressure left; friction does not refresh it (source.cpp:485-505)
            TT[s * N2 + j] = p.Tsave ? p.Tsave[idx_s(p, col, s, j)] : eint_of(q.p, q.rho, sc.gm1) / sc.cv;
            DS[s * N2 + j] = 0.;                       // reset_dS (cell 0 is never touched afterwards either)
        }
    }
    for (int j = tid; j < N2; j += NT) JJ[j] = p.J[idx_J(p, col, 0, j)];
    __syncthreads();

    // ---- stellar heating, band by band (update_tau_s_jb + update_dS_jb) ------------------------------------------------
    const int KA = Bin + 2 * p.Bout;                   // row length of kappa_avg
    for (int b = 0; b < Bin; b++) {
        double F = p.solar_heating_final[(size_t)col * Bin + b];
        if (tnow < p.radiation_rampup_time)
            F = F * (p.init_radiation_factor + (1. - p.init_radiation_factor) * tnow / p.radiation_rampup_time);
        const double top_prev = p.dSband_top[(size_t)col * Bin + b];   // read before this step's cell N overwrites it
        // cell optical depths in scan order q = N - j  (j = N .. 1)
        for (int q = tid; q < N; q += NT) {
            const int j = N - q;
            double tot = 0;
            for (int s = 0; s < S; s++) tot += kapv(p, col, s, b, j) * RHO[s * N2 + j];
            TAU[q] = tot * __ldg(p.dx + j);
        }
        __syncthreads();
        block_inclusive_scan(TAU, N, red);             // TAU[q] = radial_optical_depth_twotemp(N-q, b)
        for (int j = tid + 1; j <= N + 1; j += NT) {
            const double dxj = __ldg(p.dx + j);
            double tot = 0;
            for (int s = 0; s < S; s++) tot += kapv(p, col, s, b, j) * RHO[s * N2 + j];
            const double cell_tau = tot * dxj;
            double dSb;
            if (j < N + 1) {
                const double rad_j = TAU[N - j];
                const double rad_jp = (j == N) ? 0. : TAU[N - j - 1];
                const double Sb = F * exp(-rad_jp);
                dSb = 0.25 * Sb / dxj * (1. - p.bond_albedo);
                const double dtau_tot = -(rad_jp - rad_j);
                double dS_he = dSb;
                if (dtau_tot > 1e-3) dSb *= (-expm1(-dtau_tot));
                else dSb *= (-fastexpm1_2(-dtau_tot));
                dS_he *= -fastexpm1_2(-0.);            // cell_optical_depth_highenergy == 0 without photochemistry
                dSb -= dS_he;
                dSb = smax(dSb, 0.);
            } else {
                dSb = top_prev;                        // dS_band(N+1) = last step's dS_band(N) (radiation.cpp:263-265)
            }
            if (j == N) p.dSband_top[(size_t)col * Bin + b] = dSb;
            for (int s = 0; s < S; s++) {
                const double hsw = p.highenergy_switch[s * Bin + b];
                const double frac = hsw * kapv(p, col, s, b, j) * RHO[s * N2 + j] * dxj / cell_tau;
                double v = DS[s * N2 + j] + hsw * dSb * frac;
                if (v < 1e-50) v = 0.;
                DS[s * N2 + j] = v;
            }
        }
        __syncthreads();
    }

    // ---- thermal band: total opacity (max over species, radiation.cpp:77) -------------------------------------------------
    for (int j = tid; j < N2; j += NT) {
        double tot = 0;
        if (j >= 1) for (int s = 0; s < S; s++) tot = fmax(kapv(p, col, s, Bin + p.Bout + 0, j) * RHO[s * N2 + j], tot);
        // cell 0 is never updated by update_dS (loop j>0) and keeps its initial 0; it is not read by the solver either
        TO[j] = tot;
    }
    __syncthreads();

    // ---- FLD-simple assembly (radiation_simple.cpp:44-233) ------------------------------------------------------------------
    for (int j = tid; j < N2; j += NT) {
        double l = 0, d = 0, u = 0, r = 0;
        auto Dflux = [&](int jj) -> double {           // diffusion coefficient across the right face of cell jj (1 <= jj <= N)
            const double dxc = (__ldg(p.x_i12 + jj + 1) - __ldg(p.x_i12 + jj));
            const double t0 = TO[jj], t1 = TO[jj + 1];
            double rhokr = smax(2. * (t0 * t1) / (t0 + t1), 4. / 3. / dxc);
            rhokr = smin(0.5 * (t0 + t1), rhokr);
            const double tau_inv = 1. / (dxc * rhokr);
            const double dJ = fabs(JJ[jj + 1] - JJ[jj]) / (JJ[jj] + 1e-300);
            const double R = 1. * p.xi_rad * tau_inv * dJ;
            return 1. * tau_inv * __ldg(p.surf + jj) * flux_limiter(R);
        };
        if (j >= 2 && j <= N + 1) { const double Dm = Dflux(j - 1); d = Dm; l = -Dm; }
        if (j <= N) {
            const double tv = __ldg(p.vol + j) / (C_CLIGHT * ddt);
            d += tv;
            r += tv * JJ[j];
        }
        if (j >= 1 && j <= N) { const double Dj = Dflux(j); u = -Dj; d += Dj; }
        if (j < p.ng) { l = 0; u = -d; r = 0; }
        if (p.closed_radiative_boundaries) {
            if (j >= N2 - p.ng) { l = -l; u = 0; r = 0; }
        } else if (j == N) {
            double f = 1. / p.xi_rad * __ldg(p.x_i12 + N) / __ldg(p.x_i12 + N + 1);
            if (p.geometry == 0) f = 1;
            else if (p.geometry == 2) f *= f;
            l = -f * d; u = 0; r = 0;
        }
        if (p.radiation_matter_equilibrium_test <= 2 && j <= N) {
            double exd = 0, exr = 0;
            for (int s = 0; s < S; s++) {
                const SpeciesConst sc = p.sp[s];
                const double Ts = TT[s * N2 + j], Ts3 = Ts * Ts * Ts, rhos = RHO[s * N2 + j];
                const double kappa = kapv(p, col, s, Bin + 0, j);
                const double fac = 1. * ddt * p.no_rad_trans * kappa / sc.cv * C_SIGMA_RAD * Ts3;
                const double u1 = rhos;
                const double moredenom = -p.photocooling_expansion * 0. * ddt / (sc.cv * u1);
                const double denom = 1. + 16. * fac + moredenom;
                double eta1 = Ts * (1. + 12. * fac);
                eta1 += 1. * ddt * (DS[s * N2 + j] + 0. - p.photocooling_expansion * 0. * Ts) / u1 / sc.cv;
                const double eta2 = 4. * C_PI * ddt * kappa * p.no_rad_trans / sc.cv;
                E1[s * N2 + j] = eta1; DEN[s * N2 + j] = denom;
                exd += p.no_rad_trans * rhos * kappa * (1 - 4. * C_SIGMA_RAD * Ts3 / C_PI * eta2 / denom);
                exr += p.no_rad_trans * rhos * kappa * C_SIGMA_RAD * Ts3 / C_PI * (4 * eta1 / denom - 3 * Ts);
            }
            d += __ldg(p.vol + j) * exd;
            r += __ldg(p.vol + j) * exr;
        }
        A[j] = l; B[j] = d; C[j] = u; D[j] = r;
    }
    __syncthreads();
    if (p.tridiag_serial) {
      // Validation mode: the reference's elimination order (block_tridiag_solve.h:57-130 with 1x1 blocks) on one lane.
      // The J system is ill-conditioned once dt reaches the CFL regime; following the same order keeps the round-off
      // correlated with the reference's ("tridiagonal solve").
      if (tid == 0) {
        for (int i = 1; i < N2; i++) { const double m = A[i] / B[i - 1]; B[i] -= m * C[i - 1]; D[i] -= m * D[i - 1]; }
        D[N2 - 1] /= B[N2 - 1];
        for (int i = N2 - 2; i >= 0; i--) D[i] = (D[i] - C[i] * D[i + 1]) / B[i];
      }
      __syncthreads();
    } else {
      tridiag_pcr<CPT>(A, B, C, D, N2);
    }

    // ---- J, T update (radiation_simple.cpp:271-292, 452-479) -------------------------------------------------------------------
    int jneg = 0;
    for (int j = tid; j < N2; j += NT) {
        if (p.solve_for_j) JJ[j] = D[j];
        if (JJ[j] < 0. && p.couple_J_into_T) jneg = 1;
    }
    const int Jswitch = block_or(jneg) || (tnow > 1e15);
    TailAcc acc; acc.minstep = 0; acc.maxde = 0; acc.bad = 0; acc.badval = 0;
    int negJ = 0; double negJval = 0;
    for (int j = tid; j < N2; j += NT) {
        // crash scan on J (advection.cpp:308-338): flip tiny undershoots, flag the rest
        double Jv = JJ[j];
        if (p.couple_J_into_T && j <= N - 1 && Jv < 0) {
            if (Jv > -1e-10) Jv *= -1.;
            else if (!(j < N && p.use_planetary_temperature && JJ[N] > 0)) { negJ = 1; negJval = Jv; }
        }
        p.J[idx_J(p, col, 0, j)] = Jv;
        // ---- T_i - T_j coupling through collisions (radiation_simple.cpp:316-447): one implicit SxS solve per cell -------------
        constexpr int SC8 = 8;
        double Tcoll[SC8];
        const bool collheat = p.use_collisional_heating && S > 1;
        if (collheat && j <= N) {
            double M[SC8][SC8], b[SC8], rho8[SC8], T8[SC8], nd8[SC8];
#pragma unroll
            for (int si = 0; si < SC8; si++) {
                if (si < S) { rho8[si] = RHO[si * N2 + j]; T8[si] = TT[si * N2 + j]; nd8[si] = rho8[si] / p.sp[si].mass; }
                else { rho8[si] = 1; T8[si] = 1; nd8[si] = 0; }
            }
#pragma unroll
            for (int si = 0; si < SC8; si++) {
#pragma unroll
                for (int sj = 0; sj < SC8; sj++) M[si][sj] = (si == sj && si >= S) ? 1. : 0.;
                b[si] = 0.;
                if (si < S) {
                    const SpeciesConst ci = p.sp[si];
                    const double kappa = kapv(p, col, si, Bin + 0, j);
                    const double eta2 = 4. * C_PI * ddt * kappa * p.no_rad_trans / ci.cv;
                    M[si][si] += DEN[si * N2 + j];
                    if (p.couple_J_into_T) b[si] += eta2 * JJ[j];
                    b[si] += E1[si * N2 + j];
                    double diag_sum = 0;
#pragma unroll
                    for (int sj = 0; sj < SC8; sj++) {
                        if (sj < S) {
                            const double a = p.fric_mask[si * S + sj] *
                                             alpha_pair(p, si, sj, S, ci, p.sp[sj], rho8[si], rho8[sj], nd8[sj], T8[si], T8[sj], tnow);
                            const double temp = ddt * a * 3 * C_KB / (ci.cv * (ci.mass + p.sp[sj].mass));
                            diag_sum += temp;
                            M[si][sj] -= temp;
                        }
                    }
                    M[si][si] += diag_sum;
                }
            }
            lu_solve<SC8>(M, b, S);
            double avg_nom = 0, avg_den = 0;
#pragma unroll
            for (int si = 0; si < SC8; si++) {
                if (si < S) {
                    const SpeciesConst ci = p.sp[si];
                    double tt = b[si];
                    if (tt < p.temperature_floor) tt = p.temperature_floor;
                    if (tt > p.max_temperature) tt = p.max_temperature;
                    if (j <= 4) tt = ci.const_T_space;
                    avg_nom += rho8[si] * ci.cv * tt;
                    avg_den += rho8[si] * ci.cv;
                    Tcoll[si] = (tnow > 1e-10) ? tt : ci.const_T_space;
                }
            }
            // radiation_simple.cpp:432-444, literally (with USE_AVG_T 0 the else branch runs, and avg_temperature_t1 defaults to 1e99)
            if (p.use_avg_temperature && tnow < p.avg_temperature_t1) {
                if (tnow < p.avg_temperature_t0) {
#pragma unroll
                    for (int si = 0; si < SC8; si++) if (si < S) Tcoll[si] = avg_nom / avg_den;
                }
            } else {
#pragma unroll
                for (int si = 0; si < SC8; si++)
                    if (si < S) Tcoll[si] = avg_nom / avg_den * (1. - tnow / p.avg_temperature_t1) + Tcoll[si] * tnow / p.avg_temperature_t1;
            }
        }
        for (int s = 0; s < S; s++) {
            const SpeciesConst sc = p.sp[s];
            const double u1 = p.u[idx_u(p, 0, col, s, j)], u2 = p.u[idx_u(p, 1, col, s, j)], u3 = p.u[idx_u(p, 2, col, s, j)];
            Prim3 q = cons_to_prim(u1, u2, u3, sc.gm1);
            const double e_old = eint_of(q.p, q.rho, sc.gm1);
            // stale sound speed seen by the next CFL estimate (neither friction nor radiation refresh prim.sound_speed)
            const double cs_old = p.cssave ? p.cssave[idx_s(p, col, s, j)] : cs_of(e_old, sc.gm1);
            double T = TT[s * N2 + j];
            if (collheat) { if (j <= N) T = Tcoll[s]; }
            else if (j <= N && p.radiation_matter_equilibrium_test <= 2) {
                const double kappa = kapv(p, col, s, Bin + 0, j);
                const double eta2 = 4. * C_PI * ddt * kappa * p.no_rad_trans / sc.cv;
                double tt = E1[s * N2 + j] / DEN[s * N2 + j];
                if (p.couple_J_into_T) tt += eta2 / DEN[s * N2 + j] * JJ[j];
                if (tt < p.temperature_floor) tt = p.temperature_floor;
                if (Jswitch == 0) T = tt;
            }
            const double e = T * sc.cv;                           // update_eint_from_T
            q.p = e * q.rho * sc.gm1;                             // update_p_from_eint
            double n1, n2, n3;
            prim_to_cons(q, sc.gm1, n1, n2, n3);
            p.u[idx_u(p, 0, col, s, j)] = n1; p.u[idx_u(p, 1, col, s, j)] = n2; p.u[idx_u(p, 2, col, s, j)] = n3;
            p.dS[idx_s(p, col, s, j)] = DS[s * N2 + j];
            tail_cell(p, col, s, j, q.v, cs_old, e, T, acc);
        }
    }
    negJ = block_or(negJ);
    tail_commit(p, col, acc, red, negJ != 0, negJval);
}

template <int CPT, int NTB = 256>
__global__ void __launch_bounds__(NTB, NTB == 256 ? 2 : 1) k_rad_simple(DevParams p) {
    AIO_DYN_SMEM(double, sm_dyn);
    const int col = blockIdx.x;
    double* sm = p.scratch_rad ? p.scratch + (size_t)col * p.scratch_stride : sm_dyn;     // see k_hydro
    const ColState& c = p.cs[col];
    if (!c.in_step) return;
    rad_simple_phase<CPT>(p, col, sm, c.dt, c.t);
}

}  // namespace aio
