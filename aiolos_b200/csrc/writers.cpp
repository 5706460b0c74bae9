// Output files in the reference's layout (io.cpp:704-778 per-species output, io.cpp:787-932 diagnostic file): same file names,
// `setprecision(15)`, tab separators, ghost rows with '-' placeholders (23 tokens) and 24-column interior rows, so that the
// reference's python checkers (test_files/load_aiolos.py) read them unchanged.  Runs on the host from a downloaded column;
// it is off the time-step path.  write_monitor() adds one row to monitor_<stem>.dat (io.cpp:939-1072).
//
// Columns of output_<stem>_<species>_t<k>.dat (1-based): 1 x_i12, 2-4 u1 u2 u3, 5-7 flux, 8-10 balance, 11 pres, 12 u2/u1,
// 13 T, 14 c_s/dx, 15 cflfactor/(|v|/dx), 16 c_s, 17 dt_dE, 18 u_analytic, 19 alpha_01, 20 phi_s, 21 M_enc, 22 -dG+dGdT*T,
// 23 dS, 24 cell optical depth of outgoing band 0.
// Not reproduced (written as 0): columns 5-10 (the reference prints the interface fluxes and balances its last hydro stage
// left behind; the kernels keep them in shared memory only) and 19 (alpha_01 sample); 18 is 0 in the reference too
// (compute_analytic_solution is commented out, helpers.cpp:411-434); 22 is 0 without photochemistry.  Columns 14-17 are
// evaluated from the state being written; the reference prints the values its last dt selection stored one step earlier.
#include "engine.h"
#include <cmath>
#include <fstream>
#include <iomanip>
#include <sstream>

namespace aio {

namespace {
std::string stem_of(const std::string& simname) {
    const size_t p = simname.find('.');
    return p == std::string::npos ? simname : simname.substr(0, p);
}
}  // namespace

void write_outputs(Engine& eng, const Problem& q, int col, int index, const std::string& outdir) {
    const int N = q.num_cells, N2 = N + 2, S = q.num_species, Bin = q.num_bands_in, Bout = q.num_bands_out;
    const double c_light = 2.99792458e10, pi = 3.141592653589793, sigma_rad = 5.670374419e-5;
    const std::string dir = outdir.empty() ? q.workingdir : outdir;
    double t = 0, dt = 0; long long steps = 0; int state = 0;
    eng.status(col, &t, &dt, &steps, &state);
    const double cflfactor = eng.col_scalar(col, 0);
    std::vector<double> phi(N2), J((size_t)Bout * N2, 0.);
    eng.get_field(col, 0, 4, phi.data(), N2);
    for (int b = 0; b < Bout; b++) eng.get_field(col, 0, 16 + b, J.data() + (size_t)b * N2, N2);
    std::vector<std::vector<double>> u1(S), u2(S), u3(S), phis(S), dS(S), el(S), pres(S), T(S), cs(S);
    std::vector<double> total_opacity((size_t)Bout * N2, 0.);
    for (int s = 0; s < S; s++) {
        u1[s].resize(N2); u2[s].resize(N2); u3[s].resize(N2); phis[s].resize(N2); dS[s].resize(N2); el[s].resize(N2);
        pres[s].resize(N2); T[s].resize(N2); cs[s].resize(N2);
        eng.get_field(col, s, 0, u1[s].data(), N2); eng.get_field(col, s, 1, u2[s].data(), N2); eng.get_field(col, s, 2, u3[s].data(), N2);
        eng.get_field(col, s, 5, phis[s].data(), N2); eng.get_field(col, s, 6, dS[s].data(), N2); eng.get_field(col, s, 3, el[s].data(), N2);
        const SpeciesData& d = q.species[s];
        const double gm1 = 2. / d.degrees_of_freedom;
        for (int i = 0; i < N2; i++) {                    // eos.cpp:19-63
            const double v = u2[s][i] / u1[s][i];
            pres[s][i] = gm1 * (u3[s][i] - 0.5 * u2[s][i] * v);
            const double e = (pres[s][i] / u1[s][i]) / gm1;
            cs[s][i] = std::sqrt((gm1 + 1) * gm1 * e);
            T[s][i] = e / d.cv;
        }
        if (q.use_rad_fluxes && (q.opacity_model == 'C' || q.opacity_model == 'P'))
            for (int b = 0; b < Bout; b++) {
                double kR = q.const_opacity_rosseland_factor * d.const_opacity;
                if (q.opacity_model == 'P') kR = d.is_dust_like ? d.opacity_avg_rosseland[b] : q.const_opacity_rosseland_factor * d.opacity_avg_rosseland[b];
                for (int i = 1; i < N2; i++) total_opacity[(size_t)b * N2 + i] = std::fmax(kR * u1[s][i], total_opacity[(size_t)b * N2 + i]);
            }
    }
    // ---- per-species output (io.cpp:704-778) ----------------------------------------------------------------------------------
    for (int s = 0; s < S; s++) {
        const SpeciesData& d = q.species[s];
        std::ostringstream name;
        name << dir << "output_" << stem_of(q.simname) << "_" << d.speciesname << "_t" << index << ".dat";
        std::ofstream out(name.str(), std::ios::out);
        if (!out.is_open()) throw std::runtime_error("cannot open output file " + name.str());
        out << std::setprecision(15);
        const char tb = '\t';
        out << q.x_i12[0] << tb << u1[s][0] << tb << u2[s][0] << tb << u3[s][0] << tb << 0. << tb << 0. << tb << 0. << tb << '-' << tb << '-' << tb
            << '-' << tb << pres[s][0] << tb << u2[s][0] / u1[s][0] << tb << '-' << tb << '-' << tb << phi[0];
        for (int k = 0; k < 8; k++) out << tb << '-';
        out << std::endl;
        std::vector<double> menc(N2, 0.);
        for (int i = 2; i <= N; i++)
            menc[i] = menc[i - 1] + 4. * 3.141592 * (std::pow(q.x_i[i], 3.) - std::pow(q.x_i[i - 1], 3.)) / 3. * u1[s][i];
        for (int i = 1; i <= N; i++) {
            const double v = u2[s][i] / u1[s][i];
            const double e = (pres[s][i] / u1[s][i]) / (2. / d.degrees_of_freedom);
            const double de_e = std::abs(el[s][i] - e) / e;
            out << q.x_i12[i] << tb << u1[s][i] << tb << u2[s][i] << tb << u3[s][i] << tb << 0. << tb << 0. << tb << 0. << tb << 0. << tb << 0. << tb
                << 0. << tb << pres[s][i] << tb << v << tb << T[s][i] << tb << cs[s][i] / q.dx[i] << tb << cflfactor / std::abs(v / q.dx[i]) << tb
                << cs[s][i] << tb << dt / de_e * q.energy_epsilon << tb << 0. << tb << 0. << tb << phis[s][i] << tb << menc[i] << tb << -0. + 0. * T[s][i]
                << tb << dS[s][i] << tb << total_opacity[i] * q.dx[i] << std::endl;
        }
        out << q.x_i12[N + 1] << tb << u1[s][N + 1] << tb << u2[s][N + 1] << tb << u3[s][N + 1];
        for (int k = 0; k < 6; k++) out << tb << '-';
        out << tb << pres[s][N + 1] << tb << u2[s][N + 1] / u1[s][N + 1] << tb << T[s][N + 1] << tb << '-' << tb << phi[N + 1];
        for (int k = 0; k < 8; k++) out << tb << '-';
        out << std::endl;
    }
    // ---- diagnostic file (io.cpp:787-932): x, J_tot, S_tot, J_b, S_b, dS_b, rho*kappa_R, dtau_b, tau_in_b, kappa_S[b][s],
    //      {kappa_P[b][s], kappa_R[b][s]}_b, T_rad, T_s, sum p_s, F_b ------------------------------------------------------------------
    {
        std::ostringstream name;
        name << dir << "diagnostic_" << stem_of(q.simname) << "_t" << index << ".dat";
        std::ofstream out(name.str(), std::ios::out);
        if (!out.is_open()) throw std::runtime_error("cannot open output file " + name.str());
        out << std::setprecision(15);
        const bool constk = (q.opacity_model == 'C' || q.opacity_model == 'P');
        auto kS = [&](int s, int b) { const SpeciesData& d = q.species[s]; return (q.opacity_model == 'P' && d.is_dust_like) ? d.opacity_avg_solar[b] : q.const_opacity_solar_factor * d.opacity_avg_solar[b]; };
        auto kP = [&](int s, int b) { const SpeciesData& d = q.species[s]; if (q.opacity_model == 'P') return d.is_dust_like ? d.opacity_avg_planck[b] : q.const_opacity_planck_factor * d.opacity_avg_planck[b]; return q.const_opacity_planck_factor * d.const_opacity; };
        auto kR = [&](int s, int b) { const SpeciesData& d = q.species[s]; if (q.opacity_model == 'P') return d.is_dust_like ? d.opacity_avg_rosseland[b] : q.const_opacity_rosseland_factor * d.opacity_avg_rosseland[b]; return q.const_opacity_rosseland_factor * d.const_opacity; };
        // incoming bands: top-down optical depth, S_b, dS_b (radiation.cpp:154-265) from the state being written
        std::vector<double> tau((size_t)Bin * N2, 0.), Sb((size_t)Bin * N2, 0.), dSb((size_t)Bin * N2, 0.);
        const bool rad = q.use_rad_fluxes > 0 && constk;
        if (rad)
            for (int b = 0; b < Bin; b++) {
                double F = q.solar_heating_final[b];
                if (t < q.radiation_rampup_time) F = F * (q.init_radiation_factor + (1. - q.init_radiation_factor) * t / q.radiation_rampup_time);
                for (int j = N + 1; j > 0; j--) {
                    double tot = 0;
                    for (int s = 0; s < S; s++) tot += kS(s, b) * u1[s][j];
                    const double cell = tot * q.dx[j];
                    tau[(size_t)b * N2 + j] = (j == N + 1) ? 0. : tau[(size_t)b * N2 + j + 1] + cell;
                    Sb[(size_t)b * N2 + j] = (j >= N + 1) ? F : F * std::exp(-tau[(size_t)b * N2 + j + 1]);
                    if (j < N + 1) {
                        double v = 0.25 * Sb[(size_t)b * N2 + j] / q.dx[j] * (1. - q.bond_albedo);
                        const double dtau = -(tau[(size_t)b * N2 + j + 1] - tau[(size_t)b * N2 + j]);
                        v *= (dtau > 1e-3) ? -std::expm1(-dtau) : -(-dtau + 0.5 * dtau * dtau);
                        dSb[(size_t)b * N2 + j] = std::max(v, 0.);
                    }
                }
            }
        for (int i = 1; i <= N; i++) {
            double Jtot = 0., Stot = 0.;
            for (int b = 0; b < Bin; b++) Stot += Sb[(size_t)b * N2 + i];
            for (int b = 0; b < Bout; b++) Jtot += J[(size_t)b * N2 + i] / c_light * 4. * pi;
            out << q.x_i12[i] << '\t' << Jtot << '\t' << Stot;
            for (int b = 0; b < Bout; b++) out << '\t' << J[(size_t)b * N2 + i];
            for (int b = 0; b < Bin; b++) out << '\t' << Sb[(size_t)b * N2 + i];
            for (int b = 0; b < Bin; b++) out << '\t' << dSb[(size_t)b * N2 + i];
            for (int b = 0; b < Bout; b++) out << '\t' << total_opacity[(size_t)b * N2 + i];
            for (int b = 0; b < Bout; b++) out << '\t' << total_opacity[(size_t)b * N2 + i] * q.dx[i];
            for (int b = 0; b < Bin; b++) out << '\t' << tau[(size_t)b * N2 + i];
            for (int b = 0; b < Bin; b++) for (int s = 0; s < S; s++) out << '\t' << (rad ? kS(s, b) : 0.);
            for (int b = 0; b < Bout; b++) {
                for (int s = 0; s < S; s++) out << '\t' << (rad ? kP(s, b) : 0.);
                for (int s = 0; s < S; s++) out << '\t' << (rad ? kR(s, b) : 0.);
            }
            out << '\t' << std::pow(Jtot / sigma_rad * c_light / 4., 0.25);
            for (int s = 0; s < S; s++) out << '\t' << T[s][i];
            double tempP = 0.;
            for (int s = 0; s < S; s++) tempP += pres[s][i];
            out << '\t' << tempP;
            for (int b = 0; b < Bout; b++) {
                const double dx = (q.x_i12[i + 1] - q.x_i12[i]);
                const double t0 = total_opacity[(size_t)b * N2 + i], t1 = total_opacity[(size_t)b * N2 + i + 1];
                double rhokr = std::max(2. * (t0 * t1) / (t0 + t1), 4. / 3. / dx);
                rhokr = std::min(0.5 * (t0 + t1), rhokr);
                const double tau_inv = 1. / (dx * rhokr);
                const double Ji = J[(size_t)b * N2 + i], Jp = J[(size_t)b * N2 + i + 1];
                const double R = q.xi_rad * tau_inv * std::abs(Jp - Ji) / (Ji + 1e-300);
                const double fl = (R <= 2) ? 2 / (3 + std::sqrt(9 + 10 * R * R)) : 10 / (10 * R + 9 + std::sqrt(81 + 180 * R));
                const double D = q.surf[i] * fl * tau_inv;
                out << '\t' << -4. * pi * D * (Jp - Ji);
            }
            out << std::endl;
        }
    }
}

// One row of monitor_<stem>.dat (io.cpp:939-1072): time, dt, then quantities sampled in cell 5 (and c_v T of cell 1), all with
// setprecision(24).  `prev` carries previous_monitor_J / previous_monitor_T (B_out + S entries, zero before the first row as in
// init_and_bounds.cpp:758-759, so the first row's convergence columns are inf/nan exactly like the reference's).
// truncate != 0 starts a new file (the reference does that when steps == 0).
void write_monitor(Engine& eng, const Problem& q, int col, const std::string& outdir, bool truncate, std::vector<double>& prev) {
    const int N2 = q.num_cells + 2, S = q.num_species, Bout = q.num_bands_out;
    const double c_light = 2.99792458e10, pi = 3.141592653589793, sigma_rad = 5.670374419e-5;
    if (N2 < 6) throw std::invalid_argument("write_monitor: fewer than 6 array entries");
    if ((int)prev.size() != Bout + S) prev.assign(Bout + S, 0.0);
    const std::string dir = outdir.empty() ? q.workingdir : outdir;
    const std::string name = dir + "monitor_" + stem_of(q.simname) + ".dat";
    double t = 0, dt = 0; long long steps = 0; int state = 0;
    eng.status(col, &t, &dt, &steps, &state);
    std::vector<double> buf(N2), J5(Bout, 0.), v5(S), rho5(S), T5(S), cs5(S), T1(S);
    for (int b = 0; b < Bout; b++) { eng.get_field(col, 0, 16 + b, buf.data(), N2); J5[b] = buf[5]; }
    for (int s = 0; s < S; s++) {
        std::vector<double> u1(N2), u2(N2), u3(N2);
        eng.get_field(col, s, 0, u1.data(), N2); eng.get_field(col, s, 1, u2.data(), N2); eng.get_field(col, s, 2, u3.data(), N2);
        const SpeciesData& d = q.species[s];
        const double gm1 = 2. / d.degrees_of_freedom;
        auto temperature = [&](int i, double* cs) {        // eos.cpp:19-63
            const double v = u2[i] / u1[i], p = gm1 * (u3[i] - 0.5 * u2[i] * v), e = (p / u1[i]) / gm1;
            if (cs) *cs = std::sqrt((gm1 + 1) * gm1 * e);
            return e / d.cv;
        };
        v5[s] = u2[5] / u1[5]; rho5[s] = u1[5]; T5[s] = temperature(5, &cs5[s]); T1[s] = temperature(1, nullptr);
    }
    std::ofstream f(name, truncate ? std::ios_base::trunc : std::ios_base::app);
    if (!f.is_open()) throw std::runtime_error("cannot open " + name);
    f << std::setprecision(24);
    f << t << '\t' << dt << '\t';
    for (int s = 0; s < S; s++) f << v5[s] << '\t';
    double erad = 0;
    for (int b = 0; b < Bout; b++) erad += J5[b] * 4 * pi / c_light;
    f << erad << '\t';
    double eint = 0;
    for (int s = 0; s < S; s++) eint += rho5[s] * q.species[s].cv * T5[s];
    f << eint << '\t';
    f << std::pow(c_light * erad / sigma_rad / 4., 0.25) << '\t';
    for (int s = 0; s < S; s++) f << T5[s] << '\t';
    for (int s = 0; s < S; s++) f << cs5[s] / q.dx[5] << '\t';
    for (int b = 0; b < Bout; b++) f << J5[b] * 4 * pi / c_light << '\t';
    for (int b = 0; b < Bout; b++) { f << ((J5[b] - prev[b]) / prev[b]) << '\t'; prev[b] = J5[b]; }
    for (int s = 0; s < S; s++) {
        const double B = sigma_rad / pi * std::pow(T5[s], 4.);
        f << ((B - prev[Bout + s]) / prev[Bout + s]) << '\t';
        prev[Bout + s] = B;
    }
    for (int s = 0; s < S; s++) f << q.species[s].cv * T1[s] << '\t';
    f << std::endl;
}

}  // namespace aio
