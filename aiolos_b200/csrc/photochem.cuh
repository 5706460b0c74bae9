// C2Ray photoionisation + heating of a hydrogen-only plasma (PHOTOCHEM_LEVEL 1): c_Sim::do_photochemistry(), reference
// photochem.cpp:430-818, with its helper classes C2Ray_HOnly_ionization (:147-228), C2Ray_HOnly_heating (:241-311), the
// atomic rates of :20-92 and Brent's root finder (brent.h:15-109).  Species 0, 1, 2 are H, H+ and e- (the reference
// hard-wires that order).
//
// The sweep runs from the top cell inwards and is strictly sequential per column -- the ionising flux that reaches cell j
// depends on the optical depth of the cells above it, which depends on their ionisation fraction found by a root solve -- and
// inside a cell Brent's iteration is sequential too.  So one THREAD owns one column; the parallelism is the ensemble
// (4096 columns = 128 warps), every thread walks its own rows of the state (16 cells per cache line).
//
// Built for PARI_USE_RADIATION 0, the combination of the one shipped level-1 deck the unmodified reference can run (with
// non-vanishing ion start fractions, tests/golden/README.md): there the temperature is found by a second root solve that
// balances photo-heating, line / recombination cooling and the ion <-> neutral heat exchange.  (With radiation on, the
// reference hands heating and cooling to the FLD solvers instead; all shipped decks of that kind stop in the reference itself.)
// Reference quirks kept: the collision matrix handed to C2Ray_HOnly_heating is a default-constructed matrix (photochem.cpp:
// 617-621: `dummy`, zero in the oracle build), the top ghost cell reads the optical depth one row past the array (the next
// band's row-0 entry, which is never written: 0), kinetic-energy change dEk is multiplied by 0.
#pragma once
#include "common.cuh"
#include "hydro.cuh"
#include "friction.cuh"

namespace aio {

// ---- atomic data for hydrogen, photochem.cpp:20-92 ------------------------------------------------------------------------------
__device__ inline double H_radiative_recombination(double T_e) {
    if (T_e < 220) T_e = 220.;
    const double x = 2 * 157807 / T_e;
    return 2.753e-14 * pow(x, 1.500) / pow(1 + pow(x / 2.740, 0.407), 2.242);
}
__device__ inline double H_threebody_recombination(double T_e) {
    if (T_e < 220) T_e = 220.;
    const double x = 2 * 157807 / T_e;
    return 1.e0 * (1.005e-14 / (T_e * T_e * T_e)) * pow(x, -1.089) / pow(1 + pow(x / 0.354, 0.874), 1.101);
}
__device__ inline double H_collisional_ionization(double T_e) {
    if (T_e < 220) T_e = 220.;
    const double x = 2 * 157807. / T_e;
    const double term = 21.11 * pow(T_e, -1.5) * pow(x, -1.089) / pow(1 + pow(x / 0.354, 0.874), 1.101);
    return 1e0 * term * exp(-0.5 * x);
}
// cooling rate per electron
__device__ inline double HOnly_cooling(const double (&nX)[3], double Te) {
    const double T_HI = 157807.;
    if (Te < 220) Te = 220.;
    const double x = 2 * T_HI / Te;
    double cooling = 0;
    double term;
    if (x < 1e5) term = 3.435e-30 * Te * x * x / pow(1 + pow(x / 2.250, 0.376), 3.720);
    else term = 7.562e-27 * pow(Te, 0.42872);
    cooling += 1. * nX[1] * term;
    term = C_KB * T_HI * H_collisional_ionization(Te);
    cooling += 1. * nX[0] * term;
    term = 7.5e-19 * exp(-0.75 * T_HI / Te);
    cooling += 1. * nX[0] * term;
    term = 1.426e-27 * 1.3 * sqrt(Te);
    cooling += 1. * nX[1] * term;
    return 1. * cooling;
}

// ---- C2Ray_HOnly_ionization, photochem.cpp:147-228 -------------------------------------------------------------------------------
struct C2RayIon {
    const double* Gamma0; const double* tau0;
    double dt, nH, ne0, x0, R, C, B;
    int nb;
    __device__ C2RayIon(const double* G, const double* t, double dt_, const double (&nX)[3], double Te, int nb_)
        : Gamma0(G), tau0(t), dt(dt_), nH(nX[0] + nX[1]), ne0(nX[2]), x0(nX[1] / (nX[0] + nX[1])),
          R(H_radiative_recombination(Te)), C(H_collisional_ionization(Te)), B(H_threebody_recombination(Te)), nb(nb_) {
        C = smax(C, 1e-15 * R);
    }
    __device__ double photoionization_rate(double x_bar) const {
        double ion = 0;
        for (int b = 0; b < nb; b++) {
            if (tau0[b] * (1 - x_bar) < 1e-10) ion += Gamma0[b] / (nH)*tau0[b] * (1. - 0.5 * tau0[b] * (1 - x_bar));
            else ion += Gamma0[b] / (nH * (1 - x_bar)) * -expm1(-tau0[b] * (1 - x_bar));
        }
        return ion;
    }
    __device__ double xe_average(double x) const {
        const double ne = ne0 + (x - x0) * nH;
        const double ion = photoionization_rate(x);
        const double x_eq = (ion + ne * C) / (ion + ne * (C + R + ne * B));
        const double t_rat = dt * (ion + ne * (C + R + ne * B));
        if (t_rat > 1e-10) return x_eq - (x0 - x_eq) * expm1(-t_rat) / t_rat;
        else return x_eq + (x0 - x_eq) * (1 - 0.5 * t_rat * (1 - t_rat / 3));
    }
    __device__ double operator()(double xe) const { return xe_average(xe) - xe; }
    __device__ void nX_average(double xe, double (&out)[3]) const {
        const double xbar = xe_average(xe);
        out[0] = nH * (1 - xbar); out[1] = nH * xbar; out[2] = nH * (xbar - x0) + ne0;
    }
    __device__ void nX_new(double x_bar, double (&out)[3]) const {
        const double ne = ne0 + (x_bar - x0) * nH;
        const double ion = photoionization_rate(x_bar);
        const double x_eq = (ion + ne * C) / (ion + ne * (C + R + ne * B));
        const double t_rat = dt * (ion + ne * (C + R + ne * B));
        const double x_new = x_eq + (x0 - x_eq) * exp(-t_rat);
        out[0] = nH * (1 - x_new); out[1] = nH * x_new; out[2] = nH * x_new;
    }
};

// ---- C2Ray_HOnly_heating, photochem.cpp:241-311 (collision matrix = identity - 0 * dt, see the header) -----------------------------
struct C2RayHeat {
    double GammaH, dt, ion, recomb;
    double nX[3], T[3];
    double cm[3][3];
    __device__ C2RayHeat(double G, double dt_, const double (&n)[3], const double (&TX)[3], double ion_rate, double recomb_rate)
        : GammaH(G), dt(dt_), ion(ion_rate), recomb(recomb_rate) {
        for (int s = 0; s < 3; s++) { nX[s] = n[s]; T[s] = TX[s]; }
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) cm[a][b] = (a == b) ? 1. : 0.;
        cm[0][0] += ion * dt / nX[0];
        cm[1][0] -= ion * dt / nX[1];
        cm[1][1] += recomb * dt / nX[1];
        cm[0][1] -= recomb * dt / nX[0];
    }
    __device__ void compute_T(double Te, double (&out)[3]) const {
        double M[3][3];
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) M[a][b] = cm[a][b];
        out[0] = T[0]; out[1] = T[1]; out[2] = T[2];
        out[2] += (GammaH / nX[2] - HOnly_cooling(nX, Te)) * dt / (1.5 * C_KB);
        lu_solve<3>(M, out, 3);                       // coll_mat.partialPivLu().solve(RHS)
    }
    __device__ double operator()(double Te) const { double t[3]; compute_T(Te, t); return t[2] - Te; }
};

// ---- Brent's method, brent.h:15-109.  status: 0 ok, 1 root not bracketed, 2 no convergence (the reference throws) -------------------
// (fa0, fb0: sys(a), sys(b) when the caller has just evaluated them -- same arguments, same values, two evaluations saved)
template <class F>
__device__ double brent_solve(double a, double b, const F& sys, double tol, int max_iter, int& status, bool have = false, double fa0 = 0,
                              double fb0 = 0) {
    double fa = have ? fa0 : sys(a);
    double fb = have ? fb0 : sys(b);
    status = 0;
    if (!(fa * fb <= 0)) { status = 1; return b; }
    const double EPS = 2.220446049250313e-16;
    double c = b, fc = fb;
    double d = 0, e = 0;
    for (int i = 0; i < max_iter; i++) {
        if (fb * fc > 0) { c = a; fc = fa; e = d = b - a; }
        if (fabs(fc) < fabs(fb)) { a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
        const double tol1 = 2 * EPS * fabs(b) + 0.5 * tol;
        const double xm = 0.5 * (c - b);
        if (fabs(xm) <= tol1) return b;
        if (fabs(e) >= tol1 && fabs(fa) > fabs(fb)) {
            const double s = fb / fa;
            double pp, q;
            if (a == c) { pp = 2 * xm * s; q = 1 - s; }
            else {
                q = fa / fc;
                const double r = fb / fc;
                pp = s * (2 * xm * q * (q - r) - (b - a) * (r - 1));
                q = (q - 1) * (r - 1) * (s - 1);
            }
            if (pp > 0) q = -q; else pp = -pp;
            if (2 * pp < fmin(3 * xm * q - fabs(tol1 * q), fabs(e * q))) { e = d; d = pp / q; }
            else { d = xm; e = d; }
        } else { d = xm; e = d; }
        a = b; fa = fb;
        if (fabs(d) > tol1) b += d;
        else { if (xm > 0) b += tol1; else b -= tol1; }
        fb = sys(b);
    }
    status = 2;
    return b;
}

// ---- the sweep: one thread per column ----------------------------------------------------------------------------------------------
// Only PHOTOCHEM_LANES lanes of every warp carry a column: the sweep is one long dependent chain per thread, so what hides its
// latency is the number of resident warps, not the number of busy lanes (4096 columns: 512 warps, one per SM sub-partition, instead of 128; measured r2: 57 ms per step with 32 lanes per warp, 61 ms with 4 -- the sweep is bound by its ~60 k instructions per cell, six pow() and two exp() per evaluation of the cooling function inside Brent).
constexpr int PHOTOCHEM_LANES = 8;
__global__ void __launch_bounds__(128) k_photochem(DevParams p) {
    const int lane = threadIdx.x & 31;
    if (lane >= PHOTOCHEM_LANES) return;
    const int col = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * PHOTOCHEM_LANES + lane;
    if (col >= p.ncol) return;
    ColState& cst = p.cs[col];
    if (!cst.in_step) return;
    const double dt = cst.dt, tnow = cst.t;
    const int N = p.N, nb = p.num_he_bands, Bin = p.Bin;
    const double dlognu = 1.;
    double F[MAXB], rad_above[MAXB];                 // solar_heating(b) after reset_dS (radiation.cpp:20-31); optical depth of the cells above
    for (int b = 0; b < nb; b++) {
        double f = p.solar_heating_final[(size_t)col * Bin + b];
        if (tnow < p.radiation_rampup_time)
            f = f * (p.init_radiation_factor + (1. - p.init_radiation_factor) * tnow / p.radiation_rampup_time);
        F[b] = f;
        rad_above[b] = 0.;
    }
    const double mX[3] = {p.sp[0].mass_amu * C_AMU, p.sp[1].mass_amu * C_AMU, p.sp[2].mass_amu * C_AMU};
    const double gm1[3] = {p.sp[0].gm1, p.sp[1].gm1, p.sp[2].gm1};
    const double cv[3] = {p.sp[0].cv, p.sp[1].cv, p.sp[2].cv};
    const double mass[3] = {p.sp[0].mass, p.sp[1].mass, p.sp[2].mass};
    TailAcc acc; acc.minstep = 0; acc.maxde = 0; acc.bad = 0; acc.badval = 0;
    int fail = 0;
    for (int j = N + 1; j >= 0; j--) {
        double rho[3], vX[3], TX[3], nX[3];
        for (int s = 0; s < 3; s++) {
            const double a1 = p.u[idx_u(p, 0, col, s, j)], a2 = p.u[idx_u(p, 1, col, s, j)];
            rho[s] = a1;
            vX[s] = a2 / a1;                                       // prim.speed as the friction update left it
            TX[s] = p.Tsave[idx_s(p, col, s, j)];                  // prim.temperature of the post-hydro compute_pressure (stale on purpose)
            nX[s] = a1 / mass[s];                                  // prim.number_density
        }
        double Tnew[3] = {TX[0], TX[1], TX[2]};
        if (j >= 1) {
            for (int s = 0; s < 3; s++) { nX[s] = smax(nX[s], 0.0); TX[s] = smax(TX[s], 1.0); }
            const double dxj = __ldg(p.dx + j);
            double Gamma0[MAXB], dtaus[MAXB];
            for (int b = 0; b < nb; b++) {
                Gamma0[b] = 0.25 * F[b] / p.photon_energy0 * exp(-rad_above[b]) / dxj * dlognu;
                dtaus[b] = kapv(p, col, 0, b, j) * (nX[0] + nX[1]) * mX[0] * dxj;
            }
            C2RayIon ion(Gamma0, dtaus, dt, nX, TX[2], nb);
            const double xmin = smax((nX[1] - nX[2]) / (nX[0] + nX[1]), 0.0);
            double x_bar = 1e-10;
            {
                int st;
                const double xb = brent_solve(xmin, 1., ion, p.ion_precision, p.ion_maxiter, st);
                if (st == 0) x_bar = xb;                           // try { } catch (...) { }: a failed solve keeps 1e-10
            }
            double nX_new[3], nX_bar[3];
            ion.nX_new(x_bar, nX_new);
            ion.nX_average(x_bar, nX_bar);
            for (int b = 0; b < nb; b++) dtaus[b] *= (1 - x_bar);
            // momentum conservation
            const double ne = nX_bar[2], nH = nX_bar[0];
            const double dn_R = 1. * (ion.R + ion.B * ne) * ne * ne * dt;
            const double dn_I = 1. * (ion.C * ne + ion.photoionization_rate(x_bar)) * nH * dt;
            double mom[3] = {nX[0] * mX[0] * vX[0], nX[1] * mX[1] * vX[1], nX[2] * mX[2] * vX[2]};
            const double fe = 1 / (1 + mX[2] / mX[1]), fp = 1 / (1 + mX[1] / mX[2]);
            const double f = nX_new[0] + dn_I - dn_I * dn_R * (fp / (nX_new[1] + dn_R) + fe / (nX_new[2] + dn_R));
            mom[0] = (mom[0] + dn_R * (mom[1] / (nX_new[1] + dn_R) + mom[2] / (nX_new[2] + dn_R))) / f;
            mom[1] = (mom[1] + dn_I * mom[0] * fp) / (nX_new[1] + dn_R);
            mom[2] = (mom[2] + dn_I * mom[0] * fe) / (nX_new[2] + dn_R);
            const double dEk = 0.0 * 0.5 * (mom[0] * mom[0] * nX_new[0] / mX[0] - mX[0] * nX[0] * vX[0] * vX[0] +
                                            mom[1] * mom[1] * nX_new[1] / mX[1] - mX[1] * nX[1] * vX[1] * vX[1] +
                                            mom[2] * mom[2] * nX_new[2] / mX[2] - mX[2] * nX[2] * vX[2] * vX[2]);
            vX[0] = mom[0] / mX[0]; vX[1] = mom[1] / mX[1]; vX[2] = mom[2] / mX[2];
            for (int s = 0; s < 3; s++) rho[s] = nX_new[s] * mX[s];
            // heating
            double GammaH = 0.;
            for (int b = 0; b < nb; b++) {
                const double heatingrate = 0.25 * F[b] * exp(-rad_above[b]) * (-expm1(-dtaus[b])) *
                                           (1 - 13.6 * C_EV_TO_K * C_KB / p.photon_energy0) / dxj * dlognu;
                GammaH += heatingrate;
            }
            {
                C2RayHeat heat(GammaH - dEk / (dt + 1e-300), dt, nX_new, TX, (ion.R + ion.B * ne) * ne * ne,
                               (ion.C * ne + ion.photoionization_rate(x_bar)) * nH);
                // bracket the temperature (photochem.cpp:624-639); h1 / h2 follow heat(Te1) / heat(Te2) so that Brent does not
                // evaluate them again
                double Te1 = TX[2], Te2, h1 = heat(Te1), h2;
                if (h1 < 0) {
                    Te2 = Te1; h2 = h1; Te1 = Te2 / 1.4;
                    int guard = 0;
                    while ((h1 = heat(Te1)) < 0 && ++guard < 4096) { Te2 = Te1; h2 = h1; Te1 /= 1.4; }
                } else {
                    Te2 = Te1 * 1.4;
                    int guard = 0;
                    while ((h2 = heat(Te2)) > 0 && ++guard < 4096) { Te1 = Te2; h1 = h2; Te2 *= 1.4; }
                }
                int st;
                const double Te = brent_solve(Te1, Te2, heat, p.ion_heating_precision, p.ion_heating_maxiter, st, true, h1, h2);
                if (st != 0) fail = 1;                             // the reference throws here and ends the run
                heat.compute_T(Te, Tnew);
            }
            // optical depth through this cell (three species: no non-ionising absorbers), photochem.cpp:703-740
            for (int b = 0; b < nb; b++) {
                const double dtauplus = dtaus[b];
                if (j < N + 1) rad_above[b] = rad_above[b] + dtauplus;
                else rad_above[b] = dtauplus;
            }
        }
        // update_eint_from_T, update_p_from_eint, compute_auxillary, compute_conserved (photochem.cpp:808-813), and the step tail
        for (int s = 0; s < 3; s++) {
            const double T = Tnew[s];
            const double e = T * cv[s];
            const double pres = e * rho[s] * gm1[s];
            const double eint = (pres / rho[s]) / gm1[s];
            const double cs = cs_of(eint, gm1[s]);
            const double Tfin = eint / cv[s];
            Prim3 q; q.rho = rho[s]; q.v = vX[s]; q.p = pres;
            double n1, n2, n3;
            prim_to_cons(q, gm1[s], n1, n2, n3);
            p.u[idx_u(p, 0, col, s, j)] = n1; p.u[idx_u(p, 1, col, s, j)] = n2; p.u[idx_u(p, 2, col, s, j)] = n3;
            if (j == 2) p.prim2v[(size_t)col * 3 + s] = vX[s];
            tail_cell(p, col, s, j, vX[s], cs, eint, Tfin, acc);
        }
    }
    cst.minstep = acc.minstep; cst.maxde = acc.maxde;
    cst.steps += 1;
    if (acc.bad || fail) { cst.status = fail ? COL_NAN : COL_NEG_T; cst.t = 1.1 * cst.t_max; }
}

}  // namespace aio
