// Host front end: turns the reference's input files (.par / .spc / .opa) into a `Problem`, i.e. it restates what
// c_Sim's and c_Species' constructors compute once at start-up (reference init_and_bounds.cpp:26-1300, 1303-1674,
// io.cpp:25-693, source.cpp:26-48,120-146, helpers.cpp:194-373).  Nothing here runs per time step.
// File formats are kept exactly: parameters are found by SUBSTRING match on every line of the
// .par file, the value is the second blank-separated token parsed with operator>>, a key present on two lines is an
// error; species rows start with '@'; opacity files are opened as inputdata/<name> relative to the CWD.
#include "problem.h"
#include "hooks.h"
#include <cmath>
#include <cstdio>
#include <fstream>
#include <limits>
#include <map>
#include <tuple>
#include <sstream>
#include <stdexcept>

namespace aio {
namespace {

// physical constants, verbatim from aiolos.h:47-77
const double G = 6.678e-8, pi = 3.141592653589793, c_light = 2.99792458e10, amu = 1.66054e-24, h_planck = 6.62607015e-27;
const double Rgas = 8.31446261815324e7, kb = 1.380649e-16, mearth = 5.98e27, msolar = 2.0e33, au = 1.495978707e13;
const double rsolar = 695510 * 1e5, angstroem = 1e-4, sigma_rad = 5.670374419e-5;
const double sigma_rad2 = 2 * h_planck * c_light * c_light / std::pow(angstroem, 4.);
const double K_to_eV = 8.621738e-5, ev_to_K = 1. / K_to_eV;
const double lminglobal = 1e-6, lmaxglobal = 1e14;

std::vector<std::string> stringsplit(const std::string& str, const std::string& delim) {      // helpers.cpp:125-139
    std::vector<std::string> tokens;
    size_t prev = 0, pos = 0;
    do {
        pos = str.find(delim, prev);
        if (pos == std::string::npos) pos = str.length();
        std::string token = str.substr(prev, pos - prev);
        if (!token.empty()) tokens.push_back(token);
        prev = pos + delim.length();
    } while (pos < str.length() && prev < str.length());
    return tokens;
}

// The .par file as the reference reads it (io.cpp:595-674), with the per-column overrides applied to the text first:
// every line that contains the key is replaced by "KEY value" (or the pair is appended if no line has it).
struct ParFile {
    std::vector<std::string> lines;
    std::string name;
    void load(const std::string& path) {
        std::ifstream f(path);
        if (!f) throw std::runtime_error("cannot open parameter file " + path);
        name = path;
        std::string line;
        while (std::getline(f, line)) lines.push_back(line);
    }
    void override_key(const std::string& key, const std::string& value) {
        bool hit = false;
        for (auto& l : lines)
            if (l.find(key) != std::string::npos) {
                if (!hit) { l = key + " " + value; hit = true; }
                else l = "#";
            }
        if (!hit) lines.push_back(key + " " + value);
    }
    template <typename T> bool find(const std::string& key, T& out) const {
        int found = 0;
        for (const auto& l : lines)
            if (l.find(key) != std::string::npos) {
                auto tok = stringsplit(l, " ");
                if (tok.size() < 2) throw std::runtime_error("parameter " + key + " has no value in " + name);
                std::stringstream ss(tok[1]);
                T val = T();
                ss >> val;
                out = val;
                found++;
            }
        if (found > 1) throw std::runtime_error("ERROR: Variable " + key + " defined more than once in parameterfile!");
        return found == 1;
    }
    template <typename T> T get(const std::string& key, T dflt) const { T v = dflt; if (!find(key, v)) v = dflt; return v; }
    template <typename T> T require(const std::string& key) const {
        T v = T();
        if (!find(key, v)) throw std::runtime_error("ERROR: Variable " + key + " not found in parameterfile " + name);
        return v;
    }
};

struct AOSp { double density = 0, speed = 0, pres = 0, sound_speed = 0, internal_energy = 0, temperature = 0, number_density = 0; };

struct Sim;   // forward

// host mirror of the per-species construction state
struct Spec {
    SpeciesData d;
    double gm1 = 0, mass = 0;                    // IdealGas_EOS members (eos.h:28-47)
    std::vector<AOSp> prim, primlast;
    std::vector<double> phi_s, K_zzf;
    double SHOCK_TUBE_MID = 0.5;
    int index = 0;
    std::string opacity_data_string, opacity_corrk_string;
    void cons_from_prim(const AOSp& p, double& u1, double& u2, double& u3) const {
        u1 = p.density; u2 = p.density * p.speed; u3 = p.pres / gm1 + 0.5 * u2 * p.speed;
    }
    void compute_pressure() {                   // compute_primitive + compute_auxillary, eos.cpp:19-63
        const size_t n = d.u1.size();
        for (size_t i = 0; i < n; i++) {
            AOSp& q = prim[i];
            q.density = d.u1[i];
            q.speed = d.u2[i] / d.u1[i];
            q.pres = gm1 * (d.u3[i] - 0.5 * d.u2[i] * q.speed);
            q.number_density = q.density / mass;
            q.internal_energy = (q.pres / q.density) / gm1;
            q.sound_speed = std::sqrt((gm1 + 1) * gm1 * q.internal_energy);
            q.temperature = q.internal_energy / d.cv;
        }
    }
};

struct Sim {
    Problem q;
    std::vector<Spec> sp;
    double rs_at_moment = 0.2, cells_per_decade = 0, grid2_transition = 99.e99, grid2_cells_per_decade = 10;
    int grid2_transition_i = 0, reverse_hydrostat = 0, use_init_discont_smoothing = 0;
    double init_sonic_radius = -1e10, density_floor = 1e-20, rad_energy_multiplier = 1, rhill = 0;
    char temperature_model = 'P';
    double T_other = 0, R_other = 0, d_other = 1;
    std::string fluxfile = "---"; double fluxmultiplier = 1;

    double get_phi_grav(double r, double mass) {                     // source.cpp:120-146 at globalTime = 0
        if (0. < q.rs_time) rs_at_moment = 0.5 - (0.5 + q.rs) * (0. / q.rs_time);
        else rs_at_moment = q.rs;
        if (q.use_linear_gravity) return -G * mass * std::abs(q.domain_max - r);
        if (std::abs(r - q.planet_position) > rs_at_moment) return -G * mass / std::abs(r - q.planet_position);
        return -G * mass * (std::pow(r - q.planet_position, 3.) / std::pow(rs_at_moment, 4.) -
                            2. * std::pow(r - q.planet_position, 2.) / std::pow(rs_at_moment, 3.) + 2. / rs_at_moment);
    }
    double planck_integral2(double lmin, double lmax, double temperature) const {         // helpers.cpp:194-225
        const int num_steps = 10000;
        const double dloggrid = std::pow(lmax / lmin, 1. / ((double)num_steps));
        double tempresult = 0;
        for (int i = 0; i < num_steps; i++) {
            const double l1 = lmin * std::pow(dloggrid, (double)i);
            const double l2 = l1 * dloggrid;
            const double l_avg = 0.5 * (l1 + l2);
            const double l_avginv = 1. / l_avg;
            const double expfactor = h_planck * c_light / (l_avg * angstroem * kb * temperature);
            tempresult += sigma_rad2 * (l2 - l1) / (std::exp(expfactor) - 1.) * l_avginv * l_avginv * l_avginv * l_avginv * l_avginv;
        }
        return tempresult;
    }
    // helpers.cpp:227-373: integral3 returns 1 for a single outgoing band, integral4 for a single incoming band
    double planck_integral34(double lmin, double lmax, double temperature, int nbands) const {
        if (nbands == 1) return 1;
        const std::vector<double>& P = q.planck_matrix;
        const int np = q.num_plancks;
        double lT_min, lT_max;
        if (temperature < 2.71) { lT_min = lmin * 2.71; lT_max = lmax * 2.71; }
        else { lT_min = lmin * temperature; lT_max = lmax * temperature; }
        double m, power_min, power_max;
        if (lT_min < P[0]) {
            m = P[1] / P[0];
            power_min = P[1] + m * lT_min;
            if (lT_max < P[0]) return 1.;
        } else {
            const int imin = std::log(lT_min / P[0]) / std::log(q.lT_spacing);
            m = (P[2 * (imin + 1) + 1] - P[2 * imin + 1]) / (P[2 * (imin + 1)] - P[2 * imin]);
            power_min = P[2 * imin + 1] + m * (lT_min - P[2 * imin]);
        }
        if (lmax * temperature > P[2 * (np - 1)]) {
            power_max = 1.;
            if (lT_min > P[2 * (np - 1)]) return 1;
        } else {
            const int imax = std::log(lT_max / P[0]) / std::log(q.lT_spacing);
            m = (P[2 * (imax + 1) + 1] - P[2 * imax + 1]) / (P[2 * (imax + 1)] - P[2 * imax]);
            power_max = P[2 * imax + 1] + m * (lT_max - P[2 * imax]);
        }
        return power_max - power_min;
    }
};

// ---- .spc and .opa readers (io.cpp:25-406) -------------------------------------------------------------------------------
void read_species_data(Sim& S, Spec& s, const std::string& filename, int species_index) {
    Problem& q = S.q;
    SpeciesData& d = s.d;
    d.opacity_avg_solar.assign(q.num_bands_in, 0.);
    d.opacity_avg_planck.assign(q.num_bands_out, 0.);
    d.opacity_avg_rosseland.assign(q.num_bands_out, 0.);
    std::ifstream file(filename);
    if (!file) throw std::runtime_error("cannot open species file " + filename);
    std::string line;
    double temp_static_charge = 0.;
    int found = 0;
    while (std::getline(file, line)) {
        std::vector<std::string> t = stringsplit(line, " ");
        if (t.empty()) continue;
        if (t[0].find("@") != std::string::npos) {
            if (t.size() < 2) continue;
            if (std::stoi(t[1]) == species_index) {
                if (t.size() < 10) throw std::runtime_error("species row " + std::to_string(species_index) + " of " + filename +
                                                            " has fewer than 10 columns (io.cpp:73 reads column 10)");
                found = 1;
                d.speciesname = t[2];
                d.mass_amu = std::stod(t[3]);
                d.degrees_of_freedom = std::stod(t[4]);
                temp_static_charge = std::stod(t[5]);
                d.initial_fraction = std::stod(t[6]);
                d.density_excess = std::stod(t[7]);
                d.is_dust_like = (int)std::stod(t[8]);
                s.opacity_data_string = t[9];
                if (t.size() == 11) s.opacity_corrk_string = t[10];
                d.inv_mass = 1. / (d.mass_amu * amu);
            }
        }
    }
    if (!found) throw std::runtime_error("species number " + std::to_string(species_index) + " not found in " + filename);
    d.static_charge = (std::abs(temp_static_charge) < 0.1) ? 0 : std::floor(temp_static_charge);

    const char om = q.opacity_model;
    if (om == 'T' || om == 'K') {                  // read_opacity_table, io.cpp:421-566
        const std::string tpath = "inputdata/" + s.opacity_data_string;
        std::ifstream tf(tpath);
        if (!tf) throw std::runtime_error("cannot open opacity table " + tpath);
        auto next_tokens = [&]() { std::string l; if (!std::getline(tf, l)) throw std::runtime_error("opacity table " + tpath + " ends early"); return stringsplit(l, " "); };
        auto t = next_tokens();
        if (t.size() < 2) throw std::runtime_error("malformed header in " + tpath);
        const int np = (int)std::stod(t[0]), nT = (int)std::stod(t[1]);
        if (np <= 0 || nT <= 0 || np > 9999 || nT > 9999) throw std::runtime_error("bad grid size in " + tpath);
        d.opa_pgrid_size = np; d.opa_tgrid_size = nT;
        d.opa_pgrid.assign(np, 0.); d.opa_tgrid.assign(nT, 0.); d.opa_pgrid_log.assign(np, 0.); d.opa_tgrid_log.assign(nT, 0.);
        d.opa_grid_solar_log.assign((size_t)np * nT * q.num_bands_in, 0.);
        d.opa_grid_planck_log.assign((size_t)np * nT * q.num_bands_out, 0.);
        d.opa_grid_rosseland_log.assign((size_t)np * nT * q.num_bands_out, 0.);
        next_tokens();                               // one empty line
        t = next_tokens();
        if ((int)t.size() < np) throw std::runtime_error("short pressure grid in " + tpath);
        for (int i = 0; i < np; i++) { d.opa_pgrid[i] = std::stod(t[i]); d.opa_pgrid_log[i] = std::log10(d.opa_pgrid[i]); }
        t = next_tokens();
        if ((int)t.size() < nT) throw std::runtime_error("short temperature grid in " + tpath);
        for (int i = 0; i < nT; i++) { d.opa_tgrid[i] = std::stod(t[i]); d.opa_tgrid_log[i] = std::log10(d.opa_tgrid[i]); }
        auto read_blocks = [&](std::vector<double>& dst, int nb) {
            next_tokens();                           // header line of the block group
            for (int b = 0; b < nb; b++)
                for (int i = 0; i < np; i++) {
                    auto r = next_tokens();
                    if ((int)r.size() < nT) throw std::runtime_error("short opacity row in " + tpath);
                    for (int j = 0; j < nT; j++) dst[(size_t)j + (size_t)i * nT + (size_t)b * np * nT] = std::log10(std::stod(r[j]));
                }
        };
        read_blocks(d.opa_grid_solar_log, 1);        // only the first incoming band is read (io.cpp:503)
        read_blocks(d.opa_grid_planck_log, q.num_bands_out);
        read_blocks(d.opa_grid_rosseland_log, q.num_bands_out);
    }
    if (!(om == 'P' || om == 'M' || om == 'C' || om == 'K')) return;
    if (om == 'K') s.opacity_data_string = s.opacity_corrk_string;
    const std::string path = "inputdata/" + s.opacity_data_string;
    // The band means depend on the file, the band edges and PARI_CONST_OPAC only: columns of a sweep share them, and the
    // re-gridding below walks ~1e5 points per file, so the result is cached per process.
    struct OpaKey { std::string path; std::vector<double> lin, lout; double copac; int idx7;
                    bool operator<(const OpaKey& o) const { return std::tie(path, lin, lout, copac, idx7) < std::tie(o.path, o.lin, o.lout, o.copac, o.idx7); } };
    struct OpaVal { std::vector<double> s, p, r; };
    static std::map<OpaKey, OpaVal> opa_cache;
    const OpaKey key{path, q.l_i_in, q.l_i_out, d.const_opacity, species_index == 7 ? 1 : 0};
    {
        auto it = opa_cache.find(key);
        if (it != opa_cache.end()) { d.opacity_avg_solar = it->second.s; d.opacity_avg_planck = it->second.p; d.opacity_avg_rosseland = it->second.r; return; }
    }
    std::ifstream f2(path);
    if (!f2) throw std::runtime_error("cannot open opacity file " + path + " (opened relative to the working directory, io.cpp:143)");
    int ncolsr = 1;
    {
        auto e = stringsplit(path, ".");
        if (e.size() > 1 && e[1] == "op2") ncolsr = 2;
        if (e.size() > 1 && e[1] == "op3") ncolsr = 3;
    }
    std::vector<std::vector<double>> fd;            // rows: wl, opa...
    while (std::getline(f2, line)) {
        auto t = stringsplit(line, " ");
        std::vector<double> row(ncolsr + 1, 0.);
        if ((int)t.size() >= ncolsr + 1) for (int c = 0; c <= ncolsr; c++) row[c] = std::stod(t[c]);
        else if (t.size() >= 2) for (size_t c = 0; c < t.size() && c <= (size_t)ncolsr; c++) row[c] = std::stod(t[c]);
        else throw std::runtime_error("malformed line in " + path);
        fd.push_back(row);
    }
    const int nd = (int)fd.size();
    if (nd < 2) throw std::runtime_error("opacity file " + path + " has fewer than two rows");
    const int C = ncolsr + 1;
    // The reference indexes its column-major Eigen matrix one row past the end at the last grid point (io.cpp:236-241);
    // this accessor reproduces what that read returns (the next column's first entry) without leaving the storage.
    auto F = [&](int r, int c) -> double {
        const long lin = (long)c * nd + r;
        if (lin < 0 || lin >= (long)nd * C) return 0.0;
        return fd[lin % nd][lin / nd];
    };
    double minDeltaL = 9999999999.;
    for (int j = 0; j < nd - 1; j++) { const double dl = F(j + 1, 0) - F(j, 0); minDeltaL = (dl < minDeltaL) ? dl : minDeltaL; }
    const int ntl = (F(nd - 1, 0) - F(0, 0)) / minDeltaL + 2;
    std::vector<std::vector<double>> od(ntl, std::vector<double>(C, 0.));
    for (int col = 1; col < C; col++) {
        int j = 0;
        for (int i = 0; i < ntl; i++) {
            od[i][0] = F(0, 0) + ((double)i) * minDeltaL;
            const double wl = od[i][0];
            double lmin = F(j, 0), lmax = F(j + 1, 0);
            if (wl > lmax) { j++; lmin = F(j, 0); lmax = F(j + 1, 0); }
            if (i == 0) od[i][col] = F(0, col);
            else if (i == ntl - 1) od[i][col] = F(nd - 1, col);
            else {
                const double m = std::log10(F(j + 1, col) / F(j, col)) / std::log10(lmax / lmin);
                od[i][col] = std::pow(10., std::log10(F(j, col)) + m * std::log10(wl / lmin));
            }
        }
    }
    for (int col = 1; col < C; col++) {
        const int nb = (col == 1) ? q.num_bands_in : q.num_bands_out;
        const std::vector<double>& lgrid = (col == 1) ? q.l_i_in : q.l_i_out;
        std::vector<double>& avg = (col == 1) ? d.opacity_avg_solar : (col == 2) ? d.opacity_avg_planck : d.opacity_avg_rosseland;
        for (int b = 0; b < nb; b++) {
            int wlcount = 0;
            const double lmin = lgrid[b], lmax = lgrid[b + 1];
            avg[b] = 1e-10;
            for (int i = 0; i < ntl; i++) {
                const double wl = od[i][0];
                if (wl < lmax && wl > lmin) { avg[b] += od[i][col]; wlcount++; }
            }
            if (wlcount > 1) avg[b] /= (double)wlcount;
            if (lmax < od[0][0]) avg[b] = od[0][col];
            if (lmin > od[ntl - 1][0]) avg[b] = od[ntl - 1][col];
        }
    }
    if (ncolsr < 3) for (int b = 0; b < q.num_bands_out; b++) d.opacity_avg_rosseland[b] = d.const_opacity;
    if (ncolsr < 2) for (int b = 0; b < q.num_bands_out; b++) d.opacity_avg_planck[b] = d.const_opacity;
    auto fix = [](std::vector<double>& v) { for (auto& x : v) { if (std::isnan(x)) x = 1e-10; if (x < 1e-10) x = 1e-10; } };
    fix(d.opacity_avg_solar); fix(d.opacity_avg_planck); fix(d.opacity_avg_rosseland);
    if (species_index == 7 && q.num_bands_in >= 10) d.opacity_avg_solar[10] = 1e-10;
    opa_cache[key] = OpaVal{d.opacity_avg_solar, d.opacity_avg_planck, d.opacity_avg_rosseland};
}

// ---- link-time problem plugin (hooks.h): plain-data view of one species for the glue in engine.cu -------------------------
HookHost make_hook_host(Sim& S, Spec& s, int si) {
    Problem& q = S.q; SpeciesData& d = s.d;
    HookHost h{};
    h.species_index = si; h.num_cells = q.num_cells; h.num_ghosts = q.num_ghosts; h.num_species = q.num_species;
    h.num_bands_in = q.num_bands_in; h.num_bands_out = q.num_bands_out; h.is_dust_like = d.is_dust_like;
    h.mass_amu = d.mass_amu; h.inv_mass = d.inv_mass; h.initial_fraction = d.initial_fraction; h.density_excess = d.density_excess;
    h.static_charge = d.static_charge; h.cv = d.cv; h.gamma_adiabat = d.gamma_adiabat; h.degrees_of_freedom = d.degrees_of_freedom;
    h.const_T_space = d.const_T_space; h.const_opacity = d.const_opacity; h.gm1 = s.gm1; h.speciesname = d.speciesname.c_str();
    h.planet_mass = q.planet_mass; h.planet_semimajor = q.planet_semimajor; h.star_mass = q.star_mass; h.T_star = q.T_star;
    h.R_star = q.R_star; h.UV_star = q.UV_star; h.X_star = q.X_star; h.domain_min = q.domain_min; h.domain_max = q.domain_max;
    h.phi = q.phi.data(); h.x_i = q.x_i.data(); h.x_i12 = q.x_i12.data(); h.x_iVC = q.x_iVC.data(); h.dx = q.dx.data();
    h.surf = q.surf.data(); h.vol = q.vol.data();
    h.u1 = d.u1.data(); h.u2 = d.u2.data(); h.u3 = d.u3.data();
    return h;
}

// ---- boundary conditions on the host (applied once by the constructor, init_and_bounds.cpp:1275-1276,1779-1974) ------------
void apply_boundaries_host(Sim& S, Spec& s) {
    Problem& q = S.q;
    SpeciesData& d = s.d;
    const int ng = q.num_ghosts, N = q.num_cells, Ncell = N - 2 * (ng - 1);
    auto prim_of = [&](int i, AOSp& p) { p.density = d.u1[i]; p.speed = d.u2[i] / d.u1[i]; p.pres = s.gm1 * (d.u3[i] - 0.5 * d.u2[i] * p.speed); };
    auto set = [&](int i, const AOSp& p) { s.cons_from_prim(p, d.u1[i], d.u2[i], d.u3[i]); };
    switch (d.boundary_left) {
    case 0: { HookHost h = make_hook_host(S, s, s.index); hook_boundary_host(h, 0); break; }    // throws in the stock library
    case 1:
        for (int i = ng; i > 0; i--) {
            AOSp p; prim_of(i, p);
            double dphi = (q.phi[i] * s.K_zzf[i] - q.phi[i - 1] * s.K_zzf[i - 1]) / (q.dx[i - 1] + q.dx[i]);
            dphi *= (q.omegaplus[i] * q.dx[i] + q.omegaminus[i - 1] * q.dx[i - 1]);
            p.pres = p.pres + p.density * dphi;
            p.pres = std::max(p.pres, 0.0);
            p.speed = d.u2[i] / d.u1[i];
            set(i - 1, p);
        }
        break;
    case 2: case 5:
        for (int i = 0; i < ng; i++) {
            const int igh = ng - 1 - i, iact = ng + i;
            d.u1[igh] = d.u1[iact]; d.u2[igh] = d.u2[iact] * -1; d.u3[igh] = d.u3[iact];
            q.phi[igh] = q.phi[iact];
        }
        break;
    case 3:
        for (int i = 0; i < ng; i++) {
            AOSp p;
            if (q.problem_number != 1) {
                p.density = d.BACKGROUND_U[0];
                p.speed = (s.prim[2].speed > 0) ? s.prim[2].speed : 0.;
                p.temperature = d.const_T_space;
            }
            p.internal_energy = p.temperature * d.cv;
            p.pres = p.internal_energy * p.density * s.gm1;
            set(i, p);
        }
        break;
    case 4:
        for (int i = 0; i < ng; i++) {
            const int iact = Ncell + i;
            d.u1[i] = d.u1[iact]; d.u2[i] = d.u2[iact]; d.u3[i] = d.u3[iact];
            q.phi[i] = q.phi[iact];
        }
        break;
    default: break;
    }
    switch (d.boundary_right) {
    case 0: { HookHost h = make_hook_host(S, s, s.index); hook_boundary_host(h, 1); break; }
    case 1:
        for (int i = Ncell + ng; i < Ncell + 2 * ng; i++) {
            AOSp p; prim_of(i - 1, p);
            double dphi = (q.phi[i] * s.K_zzf[i] - q.phi[i - 1] * s.K_zzf[i - 1]) / (q.dx[i - 1] + q.dx[i]);
            dphi *= (q.omegaplus[i] * q.dx[i] + q.omegaminus[i - 1] * q.dx[i - 1]);
            p.pres = p.pres - q.right_extrap_press_multiplier * p.density * dphi;
            p.pres = std::max(p.pres, 1e-1 * p.pres);
            set(i, p);
        }
        break;
    case 2:
        for (int i = 0; i < ng; i++) {
            const int iact = Ncell + ng - 1 - i, igh = Ncell + ng + i;
            d.u1[igh] = d.u1[iact]; d.u2[igh] = d.u2[iact] * -1; d.u3[igh] = d.u3[iact];
            q.phi[igh] = q.phi[iact];
        }
        break;
    case 3:
        for (int i = Ncell + ng; i < Ncell + 2 * ng; i++) { d.u1[i] = d.SHOCK_TUBE_UR[0]; d.u2[i] = d.SHOCK_TUBE_UR[1]; d.u3[i] = d.SHOCK_TUBE_UR[2]; }
        break;
    case 4:
        for (int i = Ncell + ng; i < Ncell + 2 * ng; i++) {
            const int iact = i - Ncell;
            d.u1[i] = d.u1[iact]; d.u2[i] = d.u2[iact]; d.u3[i] = d.u3[iact];
            q.phi[i] = q.phi[iact];
        }
        break;
    case 5:
        for (int i = Ncell + ng; i < Ncell + 2 * ng; i++) {
            AOSp p; prim_of(i - 1, p);
            const double r = q.x_i12[i];
            const double mdot = 4. * 3.141592 * p.density * p.speed * r * r;
            const double freefallv = -std::sqrt(2. * G * q.planet_mass / r);
            double mdotlimit = q.max_mdot;
            if (d.is_dust_like) mdotlimit = 1e-4 * mdotlimit;
            if (mdot < mdotlimit) { p.speed = freefallv; p.density = mdotlimit / freefallv / r / r; }
            else p.pres = std::min(p.pres * 2., 1.);
            set(i, p);
        }
        break;
    default: break;
    }
}

// ---- hydrostatic atmosphere (init_and_bounds.cpp:1303-1674) -------------------------------------------------------------------
void initialize_hydrostatic_atmosphere(Sim& S, Spec& s, int species_index, const ParFile& par) {
    Problem& q = S.q;
    SpeciesData& d = s.d;
    const int N = q.num_cells;
    const double dphi_factor = par.get<double>("PARI_DPHI_FACTOR", 1.);
    if (q.photochemistry_level > 0 && std::abs(d.static_charge) > 0) S.density_floor = par.get<double>("ION_FLOOR", 1.e-20);
    else S.density_floor = par.get<double>("DENSITY_FLOOR", 1.e-20);
    for (int i = N + 1; i >= 0; i--) {
        if (S.temperature_model == 'P') {
            if (q.x_i12[i] < 1.2e9) s.prim[i].temperature = -dphi_factor * q.phi[i] / (d.cv * d.gamma_adiabat) + d.const_T_space;
            else s.prim[i].temperature = -dphi_factor * S.get_phi_grav(q.x_i12[i], q.planet_mass) / (d.cv * d.gamma_adiabat) + d.const_T_space;
        } else if (S.temperature_model == '0') {
            if (species_index == 0) {
                if (q.x_i12[i] < 1.e99) s.prim[i].temperature = -dphi_factor * q.phi[i] / (d.cv * d.gamma_adiabat) + d.const_T_space;
                else s.prim[i].temperature = -dphi_factor * S.get_phi_grav(q.x_i12[i], q.planet_mass) / (d.cv * d.gamma_adiabat) + d.const_T_space;
            } else s.prim[i].temperature = S.sp[0].prim[i].temperature;
        } else s.prim[i].temperature = d.const_T_space;
    }
    const int iter_start = N;
    auto p_over_rho = [&](double T) { return T * s.gm1 * d.cv; };            // eos.cpp:99
    if (S.reverse_hydrostat == 0) {
        for (int i = iter_start; i >= 0; i--) {
            const double T_outer = s.prim[i + 1].temperature, T_inner = s.prim[i].temperature;
            const double factor_outer = p_over_rho(T_outer), factor_inner = p_over_rho(T_inner);
            const double dphi = s.phi_s[i + 1] - s.phi_s[i];
            const double metric_outer = dphi * q.omegaplus[i + 1] * q.dx[i + 1] / (q.dx[i + 1] + q.dx[i]);
            const double metric_inner = dphi * q.omegaminus[i] * q.dx[i] / (q.dx[i + 1] + q.dx[i]);
            const double rho = d.u1[i + 1] * (factor_outer + metric_outer) / (factor_inner - metric_inner);
            d.u1[i] = rho; d.u2[i] = 0.; d.u3[i] = d.cv * rho * T_inner;
            if (rho < 0) throw std::runtime_error("NEGATIVE DENSITY IN INIT HYDROSTATIC (the reference prompts on stdin here)");
        }
    } else {
        for (int i = 0; i < 2; i++) { d.u2[i] = 0.; d.u3[i] = d.cv * d.u1[i] * s.prim[i].temperature; }
        for (int i = 1; i <= iter_start; i++) {
            const double T_outer = s.prim[i + 1].temperature, T_inner = s.prim[i].temperature;
            const double factor_outer = p_over_rho(T_outer), factor_inner = p_over_rho(T_inner);
            const double dphi = s.phi_s[i + 1] - s.phi_s[i];
            const double metric_outer = dphi * q.omegaplus[i + 1] * q.dx[i + 1] / (q.dx[i + 1] + q.dx[i]);
            const double metric_inner = dphi * q.omegaminus[i] * q.dx[i] / (q.dx[i + 1] + q.dx[i]);
            double rho = d.u1[i] * (factor_inner - metric_inner) / (factor_outer + metric_outer);
            if (dphi < 0.) rho = d.u1[i];
            if (d.mass_amu < 0.90) {
                int p_index = -1;
                for (const char* nm : {"S1", "H+", "p+", "p"}) {
                    for (size_t k = 0; k < S.sp.size() && p_index < 0; k++) if (S.sp[k].d.speciesname == nm) p_index = (int)k;
                    if (p_index >= 0) break;
                }
                if (p_index < 0) throw std::runtime_error("hydrostatic construction for electrons: proton species not found");
                rho = S.sp[p_index].d.u1[i + 1] * d.initial_fraction / S.sp[p_index].d.initial_fraction;
            }
            const double lastval_scaled = d.u1[i] * d.u1[i] / d.u1[i - 1];
            if (rho < 0.) rho = lastval_scaled;
            d.u1[i + 1] = rho; d.u2[i + 1] = 0.; d.u3[i + 1] = d.cv * rho * T_outer;
        }
    }
    for (int i = 0; i < N + 1; i++) {
        const double floor = S.density_floor * std::pow(q.x_i12[i] / q.x_i12[1], -4.) * d.initial_fraction;
        if (d.u1[i] < floor) { d.u1[i] = floor; d.u2[i] = 0.; d.u3[i] = d.cv * floor * s.prim[i].temperature; }
    }
    s.compute_pressure();
    for (int i = N; i >= 0; i--) s.primlast[i].internal_energy = s.prim[i].internal_energy;
    if (q.init_wind == 1) {
        const double csi = s.prim[N - 1].sound_speed;
        const double bondi_radius = (S.init_sonic_radius > 0) ? S.init_sonic_radius : q.planet_mass * G / (std::pow(csi, 2.));
        for (int i = 1; i < N + 1; i++) {
            const double smoothed = d.density_excess * (1. / (1. + std::exp(-(q.x_i12[i] - bondi_radius) / (0.01 * bondi_radius))));
            d.u1[i] = d.u1[i] * (1. + smoothed);
            d.u3[i] = d.u3[i] * (1. + smoothed);
        }
    }
}

// ---- c_Species constructor (init_and_bounds.cpp:1076-1290) ----------------------------------------------------------------------
void construct_species(Sim& S, int si, const ParFile& par, const std::string& species_filename) {
    Problem& q = S.q;
    Spec& s = S.sp[si];
    s.index = si;
    SpeciesData& d = s.d;
    const int N2 = q.num_cells + 2;
    d.boundary_left = par.get<int>("PARI_BOUND_TYPE_LEFT", 3);
    d.boundary_right = par.get<int>("PARI_BOUND_TYPE_RIGHT", 3);
    d.const_T_space = par.get<double>("PARI_CONST_TEMP", 1.);
    d.const_opacity = par.get<double>("PARI_CONST_OPAC", 1.);
    read_species_data(S, s, species_filename, si);
    d.u1.assign(N2, 0.); d.u2.assign(N2, 0.); d.u3.assign(N2, 0.);
    if (d.is_dust_like == 0) {
        d.cv = 0.5 * d.degrees_of_freedom * Rgas / d.mass_amu;
        d.gamma_adiabat = (d.degrees_of_freedom + 2.) / d.degrees_of_freedom;
    } else {
        d.cv = 1.5 * Rgas / d.mass_amu;
        d.mass_amu = d.mass_amu * d.degrees_of_freedom / 3;
        d.gamma_adiabat = (d.degrees_of_freedom + 2.) / d.degrees_of_freedom;
    }
    s.gm1 = 2. / d.degrees_of_freedom;             // IdealGas_EOS(dof, cv, mass): _gamma_m1 = 2/dof (eos.h:31)
    s.mass = d.mass_amu * amu;
    s.prim.assign(N2, AOSp()); s.primlast.assign(N2, AOSp());
    s.phi_s.assign(N2, 0.); s.K_zzf.assign(N2, 1.);
    {   // update_kzz_and_gravpot(species_index), source.cpp:174-219 (called by the species constructor, init_and_bounds.cpp:1162)
        const double mu = S.sp[0].d.mass_amu, mi = d.mass_amu;
        int homopause = 0;
        if (si > 0 && mi > 1e-3 && q.K_zz_init != 0) {
            for (int i = 0; i < N2; i++) {
                const double n = S.sp[0].prim[i].number_density;
                s.K_zzf[i] = (q.K_zz[i] * n < 1.) ? 1. : mu / mi;
                if (i >= 1 && s.K_zzf[i] > 0.99999 && s.K_zzf[i - 1] < 0.99999) homopause = i;
            }
        }
        for (int i = 0; i < N2; i++) s.phi_s[i] = q.phi[i] * s.K_zzf[i];
        const double corr = (homopause == 0) ? 0. : q.phi[homopause] * (1. - mu / mi);
        for (int i = 0; i < homopause; i++) s.phi_s[i] += corr;
    }
    if (q.problem_number == 1) {
        double u1l = par.get<double>("PARI_INIT_DATA_U1L", 1.), u2l = par.get<double>("PARI_INIT_DATA_U2L", 0.), u3l = par.get<double>("PARI_INIT_DATA_U3L", 1.);
        double u1r = par.get<double>("PARI_INIT_DATA_U1R", 1.), u2r = par.get<double>("PARI_INIT_DATA_U2R", 0.), u3r = par.get<double>("PARI_INIT_DATA_U3R", 1.);
        s.SHOCK_TUBE_MID = par.get<double>("PARI_INIT_SHOCK_MID", 0.5);
        u1l *= d.initial_fraction; u1r *= d.initial_fraction;
        u3l *= d.initial_fraction; u3r *= d.initial_fraction;
        if (d.is_dust_like) { u3l *= 3 / d.degrees_of_freedom; u3r *= 3 / d.degrees_of_freedom; }
        if (q.init_wind) { u2l = d.density_excess; u2r = d.density_excess; }
        AOSp pl; pl.density = u1l; pl.speed = u2l; pl.pres = u3l;
        AOSp pr; pr.density = u1r; pr.speed = u2r; pr.pres = u3r;
        s.cons_from_prim(pl, d.SHOCK_TUBE_UL[0], d.SHOCK_TUBE_UL[1], d.SHOCK_TUBE_UL[2]);
        s.cons_from_prim(pr, d.SHOCK_TUBE_UR[0], d.SHOCK_TUBE_UR[1], d.SHOCK_TUBE_UR[2]);
        for (int i = 0; i < N2; i++) {
            const double* v = (q.x_i12[i] < s.SHOCK_TUBE_MID) ? d.SHOCK_TUBE_UL : d.SHOCK_TUBE_UR;
            d.u1[i] = v[0]; d.u2[i] = v[1]; d.u3[i] = v[2];
        }
    } else if (q.problem_number == 2) {
        const double u1 = par.get<double>("PARI_INIT_DATA_U1", 1.), u2 = par.get<double>("PARI_INIT_DATA_U2", 0.), u3 = par.get<double>("PARI_INIT_DATA_U3", 1.);
        const int init_static = par.get<int>("PARI_INIT_STATIC", 1);
        AOSp p; p.density = u1 * d.initial_fraction; p.speed = u2; p.pres = u3;
        s.cons_from_prim(p, d.BACKGROUND_U[0], d.BACKGROUND_U[1], d.BACKGROUND_U[2]);
        for (int i = 0; i < N2; i++) { d.u1[i] = d.BACKGROUND_U[0]; d.u2[i] = d.BACKGROUND_U[1]; d.u3[i] = d.BACKGROUND_U[2]; }
        if (init_static == 1) initialize_hydrostatic_atmosphere(S, s, si, par);
        else if (init_static == 2) throw std::runtime_error("unsupported configuration: PARI_INIT_STATIC 2 (exponential atmosphere)");
        if (S.init_sonic_radius > 0. && q.init_wind == 2) throw std::runtime_error("unsupported configuration: PARI_INIT_WIND 2 (analytic wind)");
    } else if (q.problem_number == 3) {            // sound wave, init_and_bounds.cpp:1752-1770
        const double amp = 1e-6, rho0 = 1.0 * d.initial_fraction, cs = 1.0 * (q.domain_max - q.domain_min);
        double p0 = rho0 * cs * cs / (d.gamma_adiabat);
        if (d.is_dust_like) p0 *= 3 / d.degrees_of_freedom;
        for (int i = 0; i < N2; i++) {
            const double kx = 2 * M_PI * (i + 0.5 - q.num_ghosts) / double(q.num_cells - 2 * (q.num_ghosts - 1));
            const double f = amp * std::sin(kx);
            AOSp p; p.density = rho0 * (1 + f); p.speed = cs * f; p.pres = p0 * (1 + d.gamma_adiabat * f);
            s.cons_from_prim(p, d.u1[i], d.u2[i], d.u3[i]);
        }
    } else if (q.problem_number == 0) {
        HookHost h = make_hook_host(S, s, si);                                  // init_and_bounds.cpp:1256-1257
        hook_initial_conditions(h);                                             // throws in the stock library
    } else {
        for (int i = 0; i < N2; i++) { d.u1[i] = 1; d.u2[i] = 1; d.u3[i] = 1; }
    }
    d.USE_WAVE = par.get<int>("WAVE_USE", 0);
    if (d.USE_WAVE == 1) { d.WAVE_AMPLITUDE = par.require<double>("WAVE_AMPLITUDE"); d.WAVE_PERIOD = par.require<double>("WAVE_PERIOD"); }
    apply_boundaries_host(S, s);
}

}  // namespace

bool problem_from_files(const std::string& dir, const std::string& par_solo, const std::string& spc_solo_in,
                        const std::vector<std::pair<std::string, std::string>>& overrides, Problem& out, std::string& err) {
    try {
        Sim S;
        Problem& q = S.q;
        ParFile par;
        const std::string filename = dir + par_solo;
        par.load(filename);
        for (auto& kv : overrides) par.override_key(kv.first, kv.second);
        q.simname = par_solo; q.workingdir = dir;
        std::string spc_solo = spc_solo_in.empty() ? "default.spc" : spc_solo_in;
        if (spc_solo == "default.spc") spc_solo = par.get<std::string>("SPECIES_FILE", "default.spc");
        const std::string speciesfile = dir + spc_solo;

        // ---- numerical / physical parameters (init_and_bounds.cpp:55-253) ----
        q.type_of_grid = par.get<int>("PARI_GRID_TYPE", 0);
        q.domain_min = par.require<double>("PARI_DOMAIN_MIN");
        q.domain_max = par.require<double>("PARI_DOMAIN_MAX");
        q.geometry = par.get<int>("PARI_GEOMETRY", 0);
        q.order = par.get<int>("PARI_ORDER", 2);
        q.num_bands_in = par.get<int>("PARI_NUM_BANDS", 1);
        q.num_bands_out = par.get<int>("NUM_BANDS_OUT", q.num_bands_in);
        const double lambda_min_in = par.get<double>("PARI_LAM_MIN", 1e-1), lambda_max_in = par.get<double>("PARI_LAM_MAX", 10.);
        const double lambda_per_decade_in = par.get<double>("PARI_LAM_PER_DECADE", 10.);
        const double lambda_min_out = par.get<double>("PARI_LAM_MIN_OUT", lambda_min_in), lambda_max_out = par.get<double>("PARI_LAM_MAX_OUT", lambda_max_in);
        (void)par.get<double>("PARI_LAM_PER_DECADE_OUT", lambda_per_decade_in);
        S.fluxfile = par.get<std::string>("FLUX_FILE", "---");
        S.fluxmultiplier = par.get<double>("FLUX_MULTIPLIER", 1.);
        q.star_mass = par.get<double>("PARI_MSTAR", 1.);
        q.init_star_mass = msolar * par.get<double>("INIT_MSTAR", q.star_mass);
        q.star_mass *= msolar;
        q.ramp_star_mass_t0 = par.get<double>("RAMP_MSTAR_T0", 0.);
        q.ramp_star_mass_t1 = par.get<double>("RAMP_MSTAR_T1", 1.);
        q.T_star = par.get<double>("PARI_TSTAR", 5777.);
        q.R_star = par.get<double>("PARI_RSTAR", 0.);
        q.UV_star = par.get<double>("PARI_UVSTAR", 0.);
        q.X_star = par.get<double>("PARI_XSTAR", 0.);
        q.planet_semimajor = par.get<double>("PARI_PLANET_DIST", 1.);
        S.R_other = par.get<double>("R_OTHER", 0.); S.T_other = par.get<double>("T_OTHER", 0.); S.d_other = par.get<double>("D_OTHER", 1.);
        q.T_int = par.get<double>("PARI_TINT", 0.);
        if (q.T_int == 0) q.T_int = par.get<double>("PARI_TPLANET", 0.);
        q.use_planetary_temperature = par.get<int>("USE_PLANET_TEMPERATURE", 0);
        q.core_cv = par.get<double>("PARI_CORE_CV", 0);
        q.T_planet = par.get<double>("PARI_TPLANET_INIT", 0);
        q.radiation_matter_equilibrium_test = par.get<int>("RAD_MATTER_EQUI_TEST", 0);
        q.couple_J_into_T = par.get<int>("COUPLE_J_INTO_T", 1);
        q.num_ghosts = (q.order == 1) ? 1 : 2;
        double dx0;
        if (q.type_of_grid == 0) {
            dx0 = par.require<double>("PARI_DOMAIN_DX");
            q.num_cells = (int)((q.domain_max - q.domain_min) / dx0);
        } else {
            S.cells_per_decade = par.require<double>("PARI_CELLS_PER_DECADE");
            q.num_cells = (int)((log10f(q.domain_max) - log10f(q.domain_min)) * S.cells_per_decade);       // float, as the reference
            dx0 = q.domain_min;
            if (q.type_of_grid == 2) {
                S.grid2_transition = par.get<double>("GRID2_TRANSITION", 99.e99);
                S.grid2_cells_per_decade = par.get<double>("GRID2_CELLS_PER_DECADE", 10);
                q.num_cells = (int)((log10f(S.grid2_transition) - log10f(q.domain_min)) * S.cells_per_decade);
                S.grid2_transition_i = q.num_cells + q.num_ghosts;
                q.num_cells += (int)((log10f(q.domain_max) - log10f(S.grid2_transition)) * S.grid2_cells_per_decade);
            }
        }
        S.reverse_hydrostat = par.get<int>("REVERSE_HYDROSTAT_CONSTRUCTION", 0);
        q.init_wind = par.get<int>("PARI_INIT_WIND", 0);
        S.init_sonic_radius = par.get<double>("INIT_SONIC_RADIUS", -1e10);
        if (!(q.num_cells > 0 && q.num_cells < 999999)) throw std::runtime_error("WARNING! Something seems wrong with the number of cells required.");
        q.cflfactor = par.get<double>("PARI_CFLFACTOR", 0.9);
        q.t_max = par.get<double>("PARI_TIME_TMAX", 1e0);
        q.dt_max = par.get<double>("PARI_DTMAX", 1e99);
        q.max_timestep_change = par.get<double>("MAX_TIMESTEP_CHANGE", 1.1);
        q.dt_min_init = par.get<double>("DT_MIN_INIT", 1e-20);
        q.output_time = par.get<double>("PARI_TIME_OUTPUT", 1e99);
        q.output_time_offset = par.get<double>("TIME_OUTPUT_OFFSET", 0.);
        q.monitor_time = par.require<double>("PARI_TIME_DT");
        q.CFL_break_time = par.get<double>("CFL_BREAK_TIME", std::numeric_limits<double>::max());
        q.energy_epsilon = par.get<double>("ENERGY_EPSILON", 0.01);
        q.problem_number = par.require<int>("PARI_PROBLEM_NUMBER");
        q.use_self_gravity = par.get<int>("PARI_SELF_GRAV_SWITCH", 0);
        q.use_tides = par.get<int>("USE_TIDES", 0);
        q.use_linear_gravity = par.get<int>("PARI_LINEAR_GRAV", 0);
        q.use_rad_fluxes = par.get<int>("PARI_USE_RADIATION", 0);
        q.use_convective_fluxes = par.get<int>("USE_CONVECTION", 0);
        q.K_zz_init = par.get<double>("KZZ_INIT", 0.);
        q.use_collisional_heating = par.get<int>("PARI_USE_COLL_HEAT", 1);
        q.use_drag_predictor_step = par.get<int>("PARI_SECONDORDER_DRAG", 0);
        q.alpha_collision = par.get<double>("PARI_ALPHA_COLL", 1.);
        q.alpha_collision_ions = par.get<double>("PARI_ALPHA_IONS", 1.);
        q.max_mdot = par.get<double>("MAX_MDOT", -1.);
        S.rad_energy_multiplier = par.get<double>("PARI_RAD_MULTIPL", 1.);
        q.collision_model = par.get<char>("PARI_COLL_MODEL", 'P');
        q.opacity_model = par.get<char>("PARI_OPACITY_MODEL", 'C');
        q.use_avg_temperature = par.get<int>("USE_AVG_T", 0);
        q.avg_temperature_t0 = par.get<double>("avg_temperature_t0", 1e99);
        q.avg_temperature_t1 = par.get<double>("avg_temperature_t1", 1e99);
        if (q.opacity_model == 'M') throw std::runtime_error("unsupported configuration: PARI_OPACITY_MODEL M (Malygin/Semenov)");
        q.no_rad_trans = par.get<double>("NO_RAD_TRANS", 1.);
        q.solve_for_j = par.get<int>("SOLVE_FOR_J", 1);
        q.photocooling_expansion = par.get<double>("PHOTOCOOL_EXPANSION", 1.);
        q.radiation_rampup_time = par.get<double>("RAD_RAMPUP_TIME", 0.);
        q.init_radiation_factor = par.get<double>("INIT_RAD_FACTOR", 0.);
        q.closed_radiative_boundaries = par.get<int>("PARI_CLOSED_RADIATIVE_BOUND", 0);
        q.const_opacity_solar_factor = par.get<double>("CONSTOPA_SOLAR_FACTOR", 1.);
        q.const_opacity_rosseland_factor = par.get<double>("CONSTOPA_ROSS_FACTOR", 1.);
        q.const_opacity_planck_factor = par.get<double>("CONSTOPA_PLANCK_FACTOR", 1.);
        S.temperature_model = par.get<char>("INIT_TEMPERATURE_MODEL", 'P');
        q.friction_solver = par.get<int>("FRICTION_SOLVER", 0);
        q.update_coll_frequently = par.get<int>("UPDATE_COLL_FREQUENTLY", 1);
        q.do_hydrodynamics = par.get<int>("DO_HYDRO", 1);
        q.start_hydro_time = par.get<double>("START_HYDRO_TIME", -1.);
        q.photochemistry_level = par.get<int>("PHOTOCHEM_LEVEL", 0);
        q.ion_precision = par.get<double>("ION_PRECISION", 1e-12); q.ion_heating_precision = par.get<double>("ION_HEATING_PRECISION", 1e-12);
        q.ion_maxiter = par.get<int>("ION_MAXITER", 128); q.ion_heating_maxiter = par.get<int>("ION_HEATING_MAXITER", 128);
        q.temperature_floor = par.get<double>("TEMPERATURE_FLOOR", 0.);
        q.max_temperature = par.get<double>("TEMPERATURE_MAX", 9e99);
        q.use_total_pressure = par.get<int>("USE_TOTAL_PRESSURE", 0);
        q.coll_rampup_time = par.get<double>("COLL_RAMPUP_TIME", 0.);
        q.init_coll_factor = par.get<double>("INIT_COLL_FACTOR", 0.);
        q.right_extrap_press_multiplier = par.get<double>("BOUND_EXTRAP_MUL", 1.0);
        q.xi_rad = par.get<double>("XI_RAD", 2.);
        q.bond_albedo = par.get<double>("BOND_ALBEDO", 0.);
        S.use_init_discont_smoothing = par.get<int>("INIT_DISCONT_SMOOTHING", 0);
        q.planet_mass = par.get<double>("PARI_PLANET_MASS", (q.problem_number == 2) ? 1. : 0.);
        q.planet_mass *= mearth;
        q.planet_position = par.get<double>("PARI_PLANET_POS", 0.);
        q.rs = par.get<double>("PARI_SMOOTHING_LENGTH", 0.);
        q.rs_time = par.get<double>("PARI_SMOOTHING_TIME", 0.);
        q.rs_at_moment = 0.2;

        // ---- grid (init_and_bounds.cpp:293-590) ----
        q.num_cells += 2 * (q.num_ghosts - 1);
        const int N = q.num_cells, N2 = N + 2;
        q.x_i.assign(N + 1, 0.); q.x_i12.assign(N2, 0.); q.x_iVC.assign(N2, 0.); q.surf.assign(N + 1, 0.); q.vol.assign(N2, 0.);
        q.dx.assign(N2, 0.); q.omegaplus.assign(N2, 0.); q.omegaminus.assign(N2, 0.);
        q.source_pressure_prefactor_left.assign(N2, 0.); q.source_pressure_prefactor_right.assign(N2, 0.);
        q.phi.assign(N2, 0.);
        if (q.num_bands_in < 1) q.num_bands_in = 1;
        if (q.num_bands_out < 1) q.num_bands_out = 1;
        const int Bin = q.num_bands_in, Bout = q.num_bands_out;
        q.l_i_in.assign(Bin + 1, 0.); q.l_i_out.assign(Bout + 1, 0.);
        q.BAND_IS_HIGHENERGY.assign(Bin, 0);
        q.l_i_in[0] = lminglobal / 100.; q.l_i_in[Bin] = lmaxglobal / 100.;
        q.l_i_out[0] = lminglobal / 100.; q.l_i_out[Bout] = lmaxglobal / 100.;
        auto band_grid = [](std::vector<double>& l, int nb, double lmin, double lmax) {
            if (nb == 2) l[1] = 0.5 * (lmax + lmin);
            else if (nb > 2) {
                l[1] = lmin; l[nb - 1] = lmax;
                const double dlogl2 = std::pow(lmax / lmin, 1. / (nb - 2));
                for (int b = 2; b < nb - 1; b++) l[b] = l[b - 1] * dlogl2;
            }
        };
        band_grid(q.l_i_in, Bin, lambda_min_in, lambda_max_in);
        band_grid(q.l_i_out, Bout, lambda_min_out, lambda_max_out);
        for (int b = 0; b < Bin; b++)
            if (q.photochemistry_level > 0 && q.l_i_in[b + 1] < 0.09116) q.BAND_IS_HIGHENERGY[b] = 1;
        q.photon_energies.assign(Bin, 20. * ev_to_K * kb);
        for (int b = 0; b < Bin; b++) q.photon_energies[b] = 1.24 / (q.l_i_in[b + 1]) * ev_to_K * kb;

        if (q.type_of_grid == 1) {
            const double dlogx = std::pow(10., 1. / S.cells_per_decade);
            q.x_i[0] = q.domain_min / std::pow(dlogx, (q.num_ghosts - 1));
            for (int i = 1; i <= N; i++) q.x_i[i] = q.x_i[i - 1] * dlogx;
            q.domain_max = q.x_i[N];
        } else if (q.type_of_grid == 2) {
            const double dlogx = std::pow(10., 1. / S.cells_per_decade), dlogx2 = std::pow(10., 1. / S.grid2_cells_per_decade);
            double templogx = dlogx;
            q.x_i[0] = q.domain_min / std::pow(dlogx, (q.num_ghosts - 1));
            for (int i = 1; i <= N; i++) {
                if (i <= S.grid2_transition_i) templogx = dlogx;
                else if (i >= (S.grid2_transition_i + 5)) templogx = dlogx2;
                else templogx = dlogx + double(i - S.grid2_transition_i) / 5. * (dlogx2 - dlogx);
                q.x_i[i] = q.x_i[i - 1] * templogx;
            }
            q.domain_max = q.x_i[N];
        } else {
            q.x_i[0] = q.domain_min - dx0 * (q.num_ghosts - 1);
            for (int i = 1; i <= N; i++) q.x_i[i] = q.x_i[i - 1] + dx0;
        }
        for (int i = 1; i < N + 1; i++) q.dx[i] = q.x_i[i] - q.x_i[i - 1];
        q.dx[0] = q.dx[1] * q.dx[1] / q.dx[2];
        q.dx[N + 1] = q.dx[N] * q.dx[N] / q.dx[N - 1];
        for (int i = 0; i < N + 1; i++)
            q.surf[i] = (q.geometry == 0) ? 1 : (q.geometry == 1) ? 2 * M_PI * q.x_i[i] : 4 * M_PI * q.x_i[i] * q.x_i[i];
        for (int i = 0; i <= N + 1; i++) {
            const double xr = (i < N) ? q.x_i[i] : q.x_i[i - 1] + q.dx[i];
            const double xl = (i > 0) ? q.x_i[i - 1] : q.x_i[0] - q.dx[0];
            if (q.geometry == 0) { q.vol[i] = xr - xl; q.x_iVC[i] = (xr * xr - xl * xl) / (2 * q.vol[i]); }
            else if (q.geometry == 1) { q.vol[i] = M_PI * (xr * xr - xl * xl); q.x_iVC[i] = 2 * M_PI * (xr * xr * xr - xl * xl * xl) / (3 * q.vol[i]); }
            else { q.vol[i] = (4 * M_PI / 3) * (xr * xr * xr - xl * xl * xl); q.x_iVC[i] = (4 * M_PI) * (xr * xr * xr * xr - xl * xl * xl * xl) / (4 * q.vol[i]); }
        }
        for (int i = 1; i < N + 1; i++) q.x_i12[i] = 0.5 * (q.x_i[i] + q.x_i[i - 1]);
        q.x_i12[0] = q.x_i[0] - q.dx[0] / 2;
        q.x_i12[N + 1] = q.x_i[N] + q.dx[N + 1] / 2;
        for (int i = 0; i <= N + 1; i++) if (q.dx[i] < 0) throw std::runtime_error("WARNING! Problem with grid values (negative dx)");
        for (int i = 1; i < N + 1; i++) {
            q.omegaminus[i] = 2. * (q.dx[i - 1] + q.dx[i]) / (q.dx[i - 1] + 2. * q.dx[i] + q.dx[i + 1]);
            q.omegaplus[i] = 2. * (q.dx[i + 1] + q.dx[i]) / (q.dx[i - 1] + 2. * q.dx[i] + q.dx[i + 1]);
        }
        q.omegaminus[0] = 2 * q.omegaminus[1] - q.omegaminus[2];
        q.omegaplus[0] = 2 * q.omegaplus[1] - q.omegaplus[2];
        q.omegaminus[N + 1] = 2 * q.omegaminus[N] - q.omegaminus[N - 1];
        q.omegaplus[N + 1] = 2 * q.omegaplus[N] - q.omegaplus[N - 1];
        for (int i = 1; i < N + 1; i++) {
            q.source_pressure_prefactor_left[i] = (q.surf[i - 1] / q.vol[i] - 1. / q.dx[i]);
            q.source_pressure_prefactor_right[i] = (q.surf[i] / q.vol[i] - 1. / q.dx[i]);
        }
        // ---- gravity (source.cpp:26-48) ----
        for (int i = 1; i <= N + 1; i++) {
            q.phi[i] = S.get_phi_grav(q.x_i12[i], q.planet_mass);
            if (q.use_tides == 1) q.phi[i] -= 0.5 * 3. * G * q.init_star_mass * q.x_i12[i] * q.x_i12[i] / std::pow(q.planet_semimajor * au, 3.);
        }
        q.phi[0] = S.get_phi_grav(q.x_i12[1], q.planet_mass);
        S.rhill = q.planet_semimajor * au * std::pow(q.planet_mass / 3. / q.star_mass, 0.333333333333333333333);
        q.K_zz.assign(N2, q.K_zz_init);

        // ---- species ----
        q.num_species = par.get<int>("PARI_NUM_SPECIES", 1);
        if (q.num_species < 1) throw std::runtime_error(" Number of species has to be >= 1!");
        q.init_T_temp = par.get<double>("INIT_T_TEMP", 1.);
        const int NS = q.num_species;
        S.sp.resize(NS);
        for (int s = 0; s < NS; s++) construct_species(S, s, par, speciesfile);
        if (q.use_self_gravity) {                              // init_and_bounds.cpp:645-662: relax the atmosphere in its own potential
            std::vector<double> enclosed(N2, 0.);
            for (int cnt = 0; cnt < 50; cnt++) {
                enclosed[0] = q.planet_mass; enclosed[1] = q.planet_mass;       // update_mass_and_pot(), source.cpp:50-109, at globalTime 0
                for (int i = 2; i <= N; i++) {
                    double dens = 0.;
                    for (auto& sp : S.sp) dens += sp.d.u1[i];
                    double xn3 = q.x_i[i]; xn3 = xn3 * xn3 * xn3;
                    double xl3 = q.x_i[i - 1]; xl3 = xl3 * xl3 * xl3;
                    enclosed[i] = enclosed[i - 1] + 4. * 3.141592 * (xn3 - xl3) / 3. * dens;
                    q.phi[i] = S.get_phi_grav(q.x_i12[i], enclosed[i]);
                    if (q.use_tides == 1) {
                        double d3 = q.planet_semimajor * au;
                        d3 = d3 * d3 * d3;
                        q.phi[i] -= 0.5 * 3. * G * q.star_mass * q.x_i12[i] * q.x_i12[i] / d3;
                    }
                }
                for (int s = 0; s < NS; s++) {
                    for (int i = 0; i < N2; i++) S.sp[s].phi_s[i] = q.phi[i];
                    initialize_hydrostatic_atmosphere(S, S.sp[s], s, par);
                }
            }
        }
        q.friction_coeff_mask.assign((size_t)NS * NS, 1.); q.inv_totmasses.assign((size_t)NS * NS, 0.); q.resonant_pair_matrix.assign((size_t)NS * NS, 1.);
        for (int si = 0; si < NS; si++)
            for (int sj = 0; sj < NS; sj++) {
                q.inv_totmasses[si * NS + sj] = 1. / (S.sp[si].d.mass_amu * amu + S.sp[sj].d.mass_amu * amu);
                if (S.sp[si].d.is_dust_like == 1 && S.sp[sj].d.is_dust_like == 1) { q.friction_coeff_mask[si * NS + sj] = 0.; q.friction_coeff_mask[sj * NS + si] = 0.; }
            }
        // ---- first primitives and the constructor's dt (init_and_bounds.cpp:678-682, helpers.cpp:23-87) ----
        for (auto& s : S.sp) s.compute_pressure();
        {
            q.dt = q.t_max;
            double maxde = 0;
            for (auto& s : S.sp)
                for (int i = N - 1; i > 0; i--)
                    maxde = std::max(std::abs(s.primlast[i].internal_energy - s.prim[i].internal_energy) / s.prim[i].internal_energy, maxde);
            q.timestep_rad2 = q.dt / maxde * q.energy_epsilon;
            double minstep = 0.;
            for (auto& s : S.sp)
                for (int i = 1; i <= N; i++) minstep = std::max(minstep, std::abs(s.prim[i].speed / q.dx[i]) + s.prim[i].sound_speed / q.dx[i]);
            if (0. > q.CFL_break_time) q.cflfactor = 0.9;
            const double cfl_step = q.cflfactor / minstep;
            if (q.do_hydrodynamics) q.dt = std::min(cfl_step, q.dt * q.max_timestep_change);
            else q.dt = std::min(std::min(q.timestep_rad2, q.dt * q.max_timestep_change), q.dt_max);
        }
        // ---- Planck table, band fluxes (init_and_bounds.cpp:767-895) ----
        q.num_plancks = 512;
        {
            // the table depends on compile-time constants only (512 bins x 10000-point quadrature): built once per process
            static std::vector<double> cached;
            const double dlogx = std::pow(lmaxglobal / lminglobal, 1. / ((double)q.num_plancks));
            q.lT_spacing = dlogx;
            if (cached.empty()) {
                std::vector<double> tab((size_t)q.num_plancks * 2, 0.);
                double sum_planck = 0;
                const double lumi100 = sigma_rad * std::pow(100., 4.) / pi;
                for (int p = 0; p < q.num_plancks; p++) {
                    const double lmin = lminglobal * std::pow(dlogx, (double)p), lmax = lmin * dlogx;
                    sum_planck += S.planck_integral2(lmin, lmax, 100.) / lumi100;
                    tab[2 * p] = lmax * 100.;
                    tab[2 * p + 1] = sum_planck;
                }
                cached.swap(tab);
            }
            q.planck_matrix = cached;
        }
        q.solar_heating.assign(Bin, 0.); q.solar_heating_final.assign(Bin, 0.); q.solar_heating_thermal.assign(Bin, 0.);
        if (S.fluxfile != "---") throw std::runtime_error("unsupported configuration: FLUX_FILE");
        for (int b = 0; b < Bin; b++) {
            double F;
            if (Bin == 1) {
                F = sigma_rad * std::pow(q.T_star, 4.) * std::pow(q.R_star * rsolar, 2.) / std::pow(q.planet_semimajor * au, 2.);
                F += sigma_rad * std::pow(S.T_other, 4.) * std::pow(S.R_other * rsolar, 2.) / std::pow(S.d_other * au, 2.);
                q.solar_heating_thermal[b] = F;
                F += q.UV_star / (4. * pi * std::pow(q.planet_semimajor * au, 2.));
                F += q.X_star / (4. * pi * std::pow(q.planet_semimajor * au, 2.));
            } else {
                F = sigma_rad * std::pow(q.T_star, 4.) * std::pow(q.R_star * rsolar, 2.) / std::pow(q.planet_semimajor * au, 2.) *
                    S.planck_integral34(q.l_i_in[b], q.l_i_in[b + 1], q.T_star, Bin);
                F += sigma_rad * std::pow(S.T_other, 4.) * std::pow(S.R_other * rsolar, 2.) / std::pow(S.d_other * au, 2.) *
                     S.planck_integral34(q.l_i_in[b], q.l_i_in[b + 1], S.T_other, Bin);
                q.solar_heating_thermal[b] = F;
                if (q.l_i_in[b + 1] <= 0.09161) {
                    if (q.l_i_in[b + 1] <= 0.0161) { F += q.X_star / (4. * pi * std::pow(q.planet_semimajor * au, 2.)); q.BAND_IS_HIGHENERGY[b] = 0; }
                    else F += q.UV_star / (4. * pi * std::pow(q.planet_semimajor * au, 2.));
                }
            }
            q.solar_heating[b] = F; q.solar_heating_final[b] = F;
        }
        if (q.T_planet == 0 && q.use_planetary_temperature) {
            double T = sigma_rad * (q.T_int * q.T_int) * (q.T_int * q.T_int);
            for (int b = 0; b < Bin; b++) T += 0.25 * q.solar_heating[b];
            q.T_planet = std::pow(T / sigma_rad, 0.25);
        }
        q.highenergy_switch.assign((size_t)NS * Bin, 1.);
        q.dS_band.assign((size_t)N2 * Bin, 0.);
        // ---- initial J (init_and_bounds.cpp:942-1006) ----
        q.init_J_factor = par.get<double>("INIT_J_FACTOR", -1.);
        q.Jrad_FLD.assign((size_t)N2 * Bout, 0.); q.Jrad_init.assign((size_t)N2 * Bout, 0.);
        if (q.radiation_matter_equilibrium_test != 0) throw std::runtime_error("unsupported configuration: RAD_MATTER_EQUI_TEST != 0");
        for (int j = 0; j <= N + 1; j++) {
            if (Bout == 1) {
                for (int s = 0; s < NS; s++) {
                    q.Jrad_init[j] = S.rad_energy_multiplier * c_light / 4. / pi;
                    q.Jrad_FLD[j] = S.rad_energy_multiplier * sigma_rad * std::pow(S.sp[s].prim[j].temperature, 4) / pi;
                }
            } else {
                for (int b = 0; b < Bout; b++)
                    for (int s = 0; s < NS; s++) {
                        const double T = S.sp[s].prim[j].temperature;
                        double J = S.rad_energy_multiplier * S.planck_integral34(q.l_i_out[b], q.l_i_out[b + 1], T, Bout) * sigma_rad * std::pow(T, 4) / pi;
                        if (J < 1e-50) J = 1e-50;
                        q.Jrad_FLD[(size_t)j * Bout + b] = J;
                        q.Jrad_init[(size_t)j * Bout + b] = J;
                    }
            }
            if (j > 10) q.Jrad_FLD[(size_t)j * Bout + 0] = q.Jrad_FLD[(size_t)10 * Bout + 0] * (q.x_i12[10] * q.x_i12[10]) / (q.x_i12[j] * q.x_i12[j]);
        }
        // ---- hand the species over ----
        q.species.resize(NS);
        for (int s = 0; s < NS; s++) {
            SpeciesData& d = S.sp[s].d;
            d.eint_last.assign(N2, 0.);
            for (int i = 0; i < N2; i++) d.eint_last[i] = S.sp[s].primlast[i].internal_energy;
            d.prim2_speed_init.assign(1, N2 > 2 ? S.sp[s].prim[2].speed : 0.0);
            q.species[s] = d;
        }
        out = q;
        return true;
    } catch (const std::exception& e) {
        err = e.what();
        return false;
    }
}

}  // namespace aio
