// Reader for the "aiodump" container and construction of a Problem from a reference init dump.
// Used by aio_create_from_dump() so that the device engine can be driven from exactly the state the
// reference's constructor produced (golden fixtures under tests/golden/), independent of frontend.cpp.
#include "problem.h"
#include <cstdio>
#include <cstring>
#include <cmath>

namespace aio {

bool read_dump(const std::string& path, DumpMap& out, std::string& err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open dump " + path; return false; }
    char line[512];
    while (fgets(line, sizeof line, f)) {
        char name[400]; unsigned long long n = 0;
        if (sscanf(line, "%399s %llu", name, &n) != 2) { err = "bad dump header in " + path; fclose(f); return false; }
        std::vector<double> v(n);
        if (n && fread(v.data(), sizeof(double), n, f) != n) { err = "short dump record " + std::string(name); fclose(f); return false; }
        out[name] = std::move(v);
    }
    fclose(f);
    return true;
}

namespace {
struct Getter {
    const DumpMap& d; std::string& err; bool ok = true;
    const std::vector<double>* find(const std::string& k) {
        auto it = d.find(k);
        if (it == d.end()) { if (ok) err = "dump lacks record '" + k + "'"; ok = false; return nullptr; }
        return &it->second;
    }
    double s(const std::string& k) { auto v = find(k); return (v && !v->empty()) ? (*v)[0] : 0.0; }
    int i(const std::string& k) { return (int)std::llround(s(k)); }
    std::vector<double> v(const std::string& k) { auto p = find(k); return p ? *p : std::vector<double>(); }
    std::vector<double> vopt(const std::string& k) { auto it = d.find(k); return it == d.end() ? std::vector<double>() : it->second; }
};
}  // namespace

bool problem_from_dump(const DumpMap& d, Problem& p, std::string& err) {
    Getter g{d, err};
#define GI(x) p.x = g.i(#x);
#define GD(x) p.x = g.s(#x);
#define GV(x) p.x = g.v(#x);
    GI(num_cells) GI(num_species) GI(num_bands_in) GI(num_bands_out) GI(num_ghosts) GI(order) GI(geometry)
    GI(problem_number) GI(type_of_grid) GI(use_rad_fluxes) GI(use_tides) GI(use_self_gravity) GI(use_linear_gravity)
    GI(friction_solver) GI(photochemistry_level) GI(use_collisional_heating) GI(use_drag_predictor_step)
    GI(do_hydrodynamics) GI(closed_radiative_boundaries) GI(use_planetary_temperature) GI(couple_J_into_T)
    GI(solve_for_j) GI(radiation_matter_equilibrium_test) GI(use_convective_fluxes) GI(update_coll_frequently)
    GI(use_avg_temperature) GI(use_total_pressure) GI(init_wind) GI(num_plancks)
    p.collision_model = (char)g.i("collision_model"); p.opacity_model = (char)g.i("opacity_model");
    GD(dt) GD(cflfactor) GD(t_max) GD(max_timestep_change) GD(dt_min_init) GD(dt_max) GD(timestep_rad2)
    GD(energy_epsilon) GD(CFL_break_time) GD(start_hydro_time) GD(output_time) GD(output_time_offset) GD(monitor_time)
    GD(planet_mass) GD(planet_position) GD(planet_semimajor) GD(star_mass) GD(init_star_mass)
    GD(ramp_star_mass_t0) GD(ramp_star_mass_t1) GD(rs) GD(rs_time) GD(rs_at_moment) GD(domain_min) GD(domain_max) GD(K_zz_init)
    GD(alpha_collision) GD(alpha_collision_ions) GD(coll_rampup_time) GD(init_coll_factor)
    GD(radiation_rampup_time) GD(init_radiation_factor) GD(bond_albedo) GD(xi_rad) GD(no_rad_trans)
    GD(photocooling_expansion) GD(temperature_floor) GD(max_temperature)
    GD(const_opacity_solar_factor) GD(const_opacity_rosseland_factor) GD(const_opacity_planck_factor)
    GD(UV_star) GD(X_star) GD(R_star) GD(T_star) GD(T_int) GD(T_planet) GD(core_cv) GD(right_extrap_press_multiplier) GD(max_mdot)
    GD(lT_spacing) GD(init_J_factor) GD(init_T_temp) GD(avg_temperature_t0) GD(avg_temperature_t1)
    GV(x_i) GV(x_i12) GV(x_iVC) GV(dx) GV(surf) GV(vol) GV(omegaplus) GV(omegaminus)
    GV(source_pressure_prefactor_left) GV(source_pressure_prefactor_right) GV(phi) GV(K_zz)
    GV(l_i_in) GV(l_i_out) GV(photon_energies) GV(solar_heating) GV(solar_heating_final)
    p.solar_heating_thermal = g.vopt("solar_heating_thermal"); { auto f = g.vopt("fluxfile_used"); p.fluxfile_used = (!f.empty() && f[0] != 0); }
    GV(planck_matrix) GV(Jrad_FLD) GV(Jrad_init) GV(highenergy_switch) GV(dS_band)
    GV(friction_coeff_mask) GV(inv_totmasses) GV(resonant_pair_matrix)
    { auto b = g.v("BAND_IS_HIGHENERGY"); p.BAND_IS_HIGHENERGY.assign(b.size(), 0); for (size_t k = 0; k < b.size(); k++) p.BAND_IS_HIGHENERGY[k] = (int)b[k]; }
    p.species.resize(p.num_species);
    for (int k = 0; k < p.num_species; k++) {
        SpeciesData& sp = p.species[k];
        std::string q = "s" + std::to_string(k) + ".";
        sp.mass_amu = g.s(q + "mass_amu"); sp.inv_mass = g.s(q + "inv_mass"); sp.static_charge = g.s(q + "static_charge");
        sp.initial_fraction = g.s(q + "initial_fraction"); sp.density_excess = g.s(q + "density_excess");
        sp.is_dust_like = g.i(q + "is_dust_like"); sp.cv = g.s(q + "cv"); sp.gamma_adiabat = g.s(q + "gamma_adiabat");
        sp.degrees_of_freedom = g.s(q + "degrees_of_freedom"); sp.const_T_space = g.s(q + "const_T_space");
        sp.const_opacity = g.s(q + "const_opacity"); sp.boundary_left = g.i(q + "boundary_left");
        sp.boundary_right = g.i(q + "boundary_right"); sp.USE_WAVE = g.i(q + "USE_WAVE");
        sp.WAVE_AMPLITUDE = g.s(q + "WAVE_AMPLITUDE"); sp.WAVE_PERIOD = g.s(q + "WAVE_PERIOD");
        auto c3 = [&](const char* nm, double* dst) { auto v = g.v(q + nm); for (int c = 0; c < 3 && c < (int)v.size(); c++) dst[c] = v[c]; };
        c3("SHOCK_TUBE_UL", sp.SHOCK_TUBE_UL); c3("SHOCK_TUBE_UR", sp.SHOCK_TUBE_UR); c3("BACKGROUND_U", sp.BACKGROUND_U);
        sp.opacity_avg_solar = g.v(q + "opacity_avg_solar"); sp.opacity_avg_planck = g.v(q + "opacity_avg_planck");
        sp.opacity_avg_rosseland = g.v(q + "opacity_avg_rosseland");
        sp.u1 = g.v(q + "u1"); sp.u2 = g.v(q + "u2"); sp.u3 = g.v(q + "u3"); sp.eint_last = g.v(q + "eint_last");
        { auto v = g.v(q + "speed"); sp.prim2_speed_init.assign(1, v.size() > 2 ? v[2] : 0.0); }
        sp.opa_pgrid = g.vopt(q + "opa_pgrid"); sp.opa_tgrid = g.vopt(q + "opa_tgrid");
        sp.opa_pgrid_log = g.vopt(q + "opa_pgrid_log"); sp.opa_tgrid_log = g.vopt(q + "opa_tgrid_log");
        sp.opa_grid_solar_log = g.vopt(q + "opa_grid_solar_log"); sp.opa_grid_planck_log = g.vopt(q + "opa_grid_planck_log");
        sp.opa_grid_rosseland_log = g.vopt(q + "opa_grid_rosseland_log");
        sp.opa_pgrid_size = (int)sp.opa_pgrid.size(); sp.opa_tgrid_size = (int)sp.opa_tgrid.size();
        for (auto& kv : d) if (kv.first.compare(0, q.size() + 5, q + "name:") == 0) sp.speciesname = kv.first.substr(q.size() + 5);
    }
#undef GI
#undef GD
#undef GV
    return g.ok;
}

// Named access to a Problem's arrays and scalars (same names as the oracle dump records): lets the tests compare what the
// front end built from the input files with what the reference's constructor built.
bool problem_field(const Problem& p, const std::string& name, std::vector<double>& out) {
#define V(x) if (name == #x) { out = p.x; return true; }
#define D(x) if (name == #x) { out.assign(1, (double)p.x); return true; }
    V(x_i) V(x_i12) V(x_iVC) V(dx) V(surf) V(vol) V(omegaplus) V(omegaminus) V(source_pressure_prefactor_left)
    V(source_pressure_prefactor_right) V(phi) V(K_zz) V(l_i_in) V(l_i_out) V(photon_energies) V(solar_heating)
    V(solar_heating_final) V(solar_heating_thermal) V(planck_matrix) V(Jrad_FLD) V(Jrad_init) V(highenergy_switch) V(dS_band)
    V(friction_coeff_mask) V(inv_totmasses) V(resonant_pair_matrix)
    D(num_cells) D(num_species) D(num_bands_in) D(num_bands_out) D(num_ghosts) D(order) D(geometry) D(problem_number)
    D(use_rad_fluxes) D(use_tides) D(friction_solver) D(photochemistry_level) D(use_collisional_heating) D(dt) D(cflfactor)
    D(t_max) D(max_timestep_change) D(dt_min_init) D(timestep_rad2) D(energy_epsilon) D(CFL_break_time) D(planet_mass)
    D(star_mass) D(init_star_mass) D(planet_semimajor) D(rs) D(domain_min) D(domain_max) D(alpha_collision) D(alpha_collision_ions)
    D(bond_albedo) D(xi_rad) D(no_rad_trans) D(temperature_floor) D(max_temperature) D(const_opacity_solar_factor)
    D(const_opacity_rosseland_factor) D(const_opacity_planck_factor) D(T_star) D(UV_star) D(R_star) D(lT_spacing) D(init_wind)
    D(collision_model) D(opacity_model) D(closed_radiative_boundaries) D(right_extrap_press_multiplier) D(couple_J_into_T)
#undef V
#undef D
    if (name == "BAND_IS_HIGHENERGY") { out.assign(p.BAND_IS_HIGHENERGY.begin(), p.BAND_IS_HIGHENERGY.end()); return true; }
    if (name.size() > 2 && name[0] == 's') {
        const size_t dot = name.find('.');
        if (dot == std::string::npos) return false;
        const int k = atoi(name.substr(1, dot - 1).c_str());
        if (k < 0 || k >= (int)p.species.size()) return false;
        const SpeciesData& sp = p.species[k];
        const std::string f = name.substr(dot + 1);
#define V(x) if (f == #x) { out = sp.x; return true; }
#define D(x) if (f == #x) { out.assign(1, (double)sp.x); return true; }
        V(u1) V(u2) V(u3) V(eint_last) V(opacity_avg_solar) V(opacity_avg_planck) V(opacity_avg_rosseland)
        D(mass_amu) D(inv_mass) D(static_charge) D(initial_fraction) D(density_excess) D(is_dust_like) D(cv) D(gamma_adiabat)
        D(degrees_of_freedom) D(const_T_space) D(const_opacity) D(boundary_left) D(boundary_right) D(USE_WAVE)
#undef V
#undef D
        if (f == "SHOCK_TUBE_UL") { out.assign(sp.SHOCK_TUBE_UL, sp.SHOCK_TUBE_UL + 3); return true; }
        if (f == "SHOCK_TUBE_UR") { out.assign(sp.SHOCK_TUBE_UR, sp.SHOCK_TUBE_UR + 3); return true; }
        if (f == "BACKGROUND_U") { out.assign(sp.BACKGROUND_U, sp.BACKGROUND_U + 3); return true; }
    }
    return false;
}

}  // namespace aio
