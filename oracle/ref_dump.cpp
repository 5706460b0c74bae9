// Oracle dump harness (TEST INFRASTRUCTURE ONLY; never linked into the product).
//
// Links against the UNMODIFIED reference sources compiled from /root/reference (see Makefile) and
// drives them through their own public API: c_Sim ctor (init_and_bounds.cpp:26) + c_Sim::execute()
// (advection.cpp:18).  All c_Sim / c_Species members are public (aiolos.h:363,843), so the state can be
// dumped without touching the reference.  Writes
//   <out>_init.bin   problem description + initial state right after the constructor
//   <out>_final.bin  state after execute() returned (run to t_max)
// in the "aiodump" container format read by tests/aiodump.py and aiolos_b200/csrc/dumpio.h:
//   repeated records:  "<name> <count>\n"  followed by <count> little-endian float64.
// The per-step dt sequence is captured through a forwarding EOS object (the reference calls the EOS
// through a virtual base pointer, eos.h:13-26): the first compute_primitive() of every step sees the
// dt chosen for that step (advection.cpp:73-77 run before :152).
//
// usage: aiolos_ref_dump -dir <deckdir/> -par <file.par> [-spc <file.spc>] [-tmax T] [-nsteps K] -out <prefix>
#include "aiolos.h"
#include <cstdio>

static FILE* g_out = nullptr;
static void put(const std::string& name, const double* v, size_t n) {
    fprintf(g_out, "%s %zu\n", name.c_str(), n);
    if (n) fwrite(v, sizeof(double), n, g_out);
}
static void put1(const std::string& name, double v) { put(name, &v, 1); }
static void putv(const std::string& name, const std::vector<double>& v) { put(name, v.data(), v.size()); }
template <typename M> static void putm(const std::string& name, const M& m) {
    std::vector<double> t((size_t)m.rows() * m.cols());
    for (int i = 0; i < m.rows(); i++) for (int j = 0; j < m.cols(); j++) t[(size_t)i * m.cols() + j] = m(i, j);
    put(name, t.data(), t.size());
}

struct HookEOS final : public EOS_Base {
    EOS_Base* inner; c_Sim* sim; std::vector<double>* dts; long long last = -1; long long stop_after = -1; int ncall = 0;
    HookEOS(EOS_Base* e, c_Sim* s, std::vector<double>* d) : inner(e), sim(s), dts(d) {}
    ~HookEOS() { delete inner; }
    void compute_primitive(const AOS* a, AOS_prim* b, int n) const override {
        HookEOS* self = const_cast<HookEOS*>(this);
        if (sim->steps != self->last) { self->last = sim->steps; self->ncall = 0; }
        self->ncall++;
        if (sim->steps >= 0) {
            if ((size_t)sim->steps >= dts->size()) dts->resize(sim->steps + 1, 0.0);
            (*dts)[sim->steps] = sim->dt;      // overwritten by every call of the step: the last one saw this step's dt
            // -nsteps K: let step K-1 run with its natural dt, then leave the loop (its condition is globalTime < t_max,
            // advection.cpp:60).  dt = min(dt, t_max - globalTime) (advection.cpp:75) must already have been evaluated:
            // in step 0 the first call comes from the steps==0 priming (advection.cpp:67-71), before dt is chosen.
            if (stop_after > 0 && sim->steps == stop_after - 1 && self->ncall >= (sim->steps == 0 ? 2 : 1))
                sim->t_max = sim->globalTime;
        }
        inner->compute_primitive(a, b, n);
    }
    void compute_conserved(const AOS_prim* a, AOS* b, int n) const override { inner->compute_conserved(a, b, n); }
    void compute_auxillary(AOS_prim* a, int n) const override { inner->compute_auxillary(a, n); }
    void update_p_from_eint(AOS_prim* a, int n) const override { inner->update_p_from_eint(a, n); }
    void update_eint_from_T(AOS_prim* a, int n) const override { inner->update_eint_from_T(a, n); }
    void get_p_over_rho_analytic(const double* a, double* b) const override { inner->get_p_over_rho_analytic(a, b); }
    void get_p_from_rhoT(const double* a, const double* b, double* c) const override { inner->get_p_from_rhoT(a, b, c); }
};

static void dump_state(c_Sim& s) {
    const int N2 = s.num_cells + 2;
    put1("steps", (double)s.steps); put1("globalTime", s.globalTime); put1("dt", s.dt);
    put1("cflfactor", s.cflfactor); put1("timestep_rad2", s.timestep_rad2); put1("cfl_step", s.cfl_step);
    put1("T_planet", s.T_planet); put1("max_snd_crs_time", s.max_snd_crs_time);
    putv("phi", s.phi); putv("enclosed_mass", s.enclosed_mass);
    putm("Jrad_FLD", s.Jrad_FLD); putm("S_band", s.S_band); putm("dS_band", s.dS_band);
    putm("solar_heating", s.solar_heating);
    putm("total_opacity", s.total_opacity); putm("cell_optical_depth", s.cell_optical_depth);
    putm("radial_optical_depth", s.radial_optical_depth);
    putm("total_opacity_twotemp", s.total_opacity_twotemp);
    putm("cell_optical_depth_twotemp", s.cell_optical_depth_twotemp);
    putm("radial_optical_depth_twotemp", s.radial_optical_depth_twotemp);
    putm("alphas_sample", s.alphas_sample);
    for (int k = 0; k < s.num_species; k++) {
        c_Species& sp = s.species[k];
        std::string p = "s" + std::to_string(k) + ".";
        std::vector<double> a(N2);
#define FIELD(nm, expr) for (int i = 0; i < N2; i++) a[i] = (expr); putv(p + nm, a);
        FIELD("u1", sp.u[i].u1) FIELD("u2", sp.u[i].u2) FIELD("u3", sp.u[i].u3)
        FIELD("density", sp.prim[i].density) FIELD("speed", sp.prim[i].speed) FIELD("pres", sp.prim[i].pres)
        FIELD("sound_speed", sp.prim[i].sound_speed) FIELD("internal_energy", sp.prim[i].internal_energy)
        FIELD("temperature", sp.prim[i].temperature) FIELD("number_density", sp.prim[i].number_density)
        FIELD("eint_last", sp.primlast[i].internal_energy) FIELD("de_e", sp.de_e[i])
        FIELD("phi_s", sp.phi_s[i]) FIELD("K_zzf", sp.K_zzf[i])
        FIELD("dudt0_1", sp.dudt[0][i].u1) FIELD("dudt0_2", sp.dudt[0][i].u2) FIELD("dudt0_3", sp.dudt[0][i].u3)
#undef FIELD
        if (sp.dS.size() == N2) { putm(p + "dS", sp.dS); putm(p + "dG", sp.dG); putm(p + "dGdT", sp.dGdT); }
        putm(p + "opacity", sp.opacity); putm(p + "opacity_planck", sp.opacity_planck);
        putm(p + "opacity_twotemp", sp.opacity_twotemp);
        putm(p + "fraction_total_solar_opacity", sp.fraction_total_solar_opacity);
    }
}

static void dump_problem(c_Sim& s) {
#define I(x) put1(#x, (double)s.x);
    I(num_cells) I(num_species) I(num_bands_in) I(num_bands_out) I(num_ghosts) I(problem_number)
    I(use_rad_fluxes) I(use_tides) I(use_self_gravity) I(use_linear_gravity) I(friction_solver)
    I(collision_model) I(opacity_model) I(photochemistry_level) I(use_collisional_heating)
    I(use_drag_predictor_step) I(do_hydrodynamics) I(closed_radiative_boundaries) I(use_planetary_temperature)
    I(couple_J_into_T) I(solve_for_j) I(radiation_matter_equilibrium_test) I(use_convective_fluxes)
    I(update_coll_frequently) I(use_avg_temperature) I(use_total_pressure) I(type_of_grid) I(init_wind)
    I(t_max) I(max_timestep_change) I(dt_min_init) I(dt_max) I(energy_epsilon) I(CFL_break_time)
    I(start_hydro_time) I(output_time) I(output_time_offset) I(monitor_time)
    I(planet_mass) I(planet_position) I(planet_semimajor) I(star_mass) I(init_star_mass)
    I(ramp_star_mass_t0) I(ramp_star_mass_t1) I(rs) I(rs_time) I(rs_at_moment) I(domain_min) I(domain_max)
    I(alpha_collision) I(alpha_collision_ions) I(coll_rampup_time) I(init_coll_factor)
    I(radiation_rampup_time) I(init_radiation_factor) I(bond_albedo) I(xi_rad) I(no_rad_trans)
    I(photocooling_expansion) I(photocooling_multiplier) I(temperature_floor) I(max_temperature)
    I(const_opacity_solar_factor) I(const_opacity_rosseland_factor) I(const_opacity_planck_factor)
    I(T_star) I(R_star) I(UV_star) I(X_star) I(T_int) I(core_cv) I(R_core) I(right_extrap_press_multiplier) I(max_mdot)
    I(K_zz_init) I(lT_spacing) I(init_J_factor) I(init_T_temp) I(avg_temperature_t0) I(avg_temperature_t1)
    I(density_floor) I(rhill) I(num_plancks) I(rad_energy_multiplier)
#undef I
    put1("order", (double)(int)s.order); put1("geometry", (double)(int)s.geometry);
    putv("x_i", s.x_i); putv("x_i12", s.x_i12); putv("x_iVC", s.x_iVC); putv("dx", s.dx);
    putv("surf", s.surf); putv("vol", s.vol); putv("omegaplus", s.omegaplus); putv("omegaminus", s.omegaminus);
    putv("source_pressure_prefactor_left", s.source_pressure_prefactor_left);
    putv("source_pressure_prefactor_right", s.source_pressure_prefactor_right);
    putv("K_zz", s.K_zz); putv("l_i_in", s.l_i_in); putv("l_i_out", s.l_i_out);
    putv("photon_energies", s.photon_energies);
    { std::vector<double> t(s.BAND_IS_HIGHENERGY.begin(), s.BAND_IS_HIGHENERGY.end()); putv("BAND_IS_HIGHENERGY", t); }
    putm("solar_heating_final", s.solar_heating_final); putm("planck_matrix", s.planck_matrix);
    {   // thermal (black-body) part of the top-of-atmosphere band fluxes, same expressions as init_and_bounds.cpp:808-826,
        // so that a per-column PARI_UVSTAR / PARI_XSTAR override can be applied to a dump exactly as the constructor would
        std::vector<double> th(s.num_bands_in, 0.0);
        for (int b = 0; b < s.num_bands_in; b++) {
            if (s.num_bands_in == 1) {
                th[b]  = sigma_rad * pow(s.T_star,4.) * pow(s.R_star*rsolar,2.)/pow(s.planet_semimajor*au,2.);
                th[b] += sigma_rad * pow(s.T_other,4.) * pow(s.R_other*rsolar,2.)/pow(s.d_other*au,2.);
            } else {
                th[b]  = sigma_rad * pow(s.T_star,4.) * pow(s.R_star*rsolar,2.)/pow(s.planet_semimajor*au,2.) * s.compute_planck_function_integral4(s.l_i_in[b], s.l_i_in[b+1], s.T_star);
                th[b] += sigma_rad * pow(s.T_other,4.) * pow(s.R_other*rsolar,2.)/pow(s.d_other*au,2.) * s.compute_planck_function_integral4(s.l_i_in[b], s.l_i_in[b+1], s.T_other);
            }
        }
        putv("solar_heating_thermal", th);
        put1("fluxfile_used", s.fluxfile.compare("---") != 0 ? 1.0 : 0.0);
    }
    putm("Jrad_init", s.Jrad_init); putm("highenergy_switch", s.highenergy_switch);
    putm("cell_optical_depth_highenergy", s.cell_optical_depth_highenergy);
    putm("friction_coeff_mask", s.friction_coeff_mask); putm("inv_totmasses", s.inv_totmasses);
    putm("resonant_pair_matrix", s.resonant_pair_matrix);
    for (int k = 0; k < s.num_species; k++) {
        c_Species& sp = s.species[k];
        std::string p = "s" + std::to_string(k) + ".";
#define S(x) put1(p + #x, (double)sp.x);
        S(mass_amu) S(inv_mass) S(static_charge) S(initial_fraction) S(density_excess) S(is_dust_like)
        S(cv) S(gamma_adiabat) S(degrees_of_freedom) S(const_T_space) S(const_opacity)
        S(USE_WAVE) S(pressure_broadening_factor) S(pressure_broadening_exponent)
#undef S
        put1(p + "WAVE_AMPLITUDE", sp.USE_WAVE == 1 ? sp.WAVE_AMPLITUDE : 0.0);
        put1(p + "WAVE_PERIOD", sp.USE_WAVE == 1 ? sp.WAVE_PERIOD : 1.0);
        put1(p + "boundary_left", (double)(int)sp.boundary_left);
        put1(p + "boundary_right", (double)(int)sp.boundary_right);
        double t3[3];
        t3[0] = sp.SHOCK_TUBE_UL.u1; t3[1] = sp.SHOCK_TUBE_UL.u2; t3[2] = sp.SHOCK_TUBE_UL.u3; put(p + "SHOCK_TUBE_UL", t3, 3);
        t3[0] = sp.SHOCK_TUBE_UR.u1; t3[1] = sp.SHOCK_TUBE_UR.u2; t3[2] = sp.SHOCK_TUBE_UR.u3; put(p + "SHOCK_TUBE_UR", t3, 3);
        t3[0] = sp.BACKGROUND_U.u1; t3[1] = sp.BACKGROUND_U.u2; t3[2] = sp.BACKGROUND_U.u3; put(p + "BACKGROUND_U", t3, 3);
        putm(p + "opacity_avg_solar", sp.opacity_avg_solar); putm(p + "opacity_avg_planck", sp.opacity_avg_planck);
        putm(p + "opacity_avg_rosseland", sp.opacity_avg_rosseland);
        if (s.opacity_model == 'T' || s.opacity_model == 'K') {
            putm(p + "opa_pgrid", sp.opa_pgrid); putm(p + "opa_tgrid", sp.opa_tgrid);
            putm(p + "opa_pgrid_log", sp.opa_pgrid_log); putm(p + "opa_tgrid_log", sp.opa_tgrid_log);
            putm(p + "opa_grid_solar_log", sp.opa_grid_solar_log);
            putm(p + "opa_grid_planck_log", sp.opa_grid_planck_log);
            putm(p + "opa_grid_rosseland_log", sp.opa_grid_rosseland_log);
        }
        fprintf(g_out, "%sname:%s 0\n", p.c_str(), sp.speciesname.c_str());
    }
    dump_state(s);
}

int main(int argc, char** argv) {
    std::string dir = "./", par = "", spc = "default.spc", out = "dump";
    double tmax = -1; int have_tmax = 0; long long nsteps = -1;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "-dir" && i + 1 < argc) dir = argv[++i];
        else if (a == "-par" && i + 1 < argc) par = argv[++i];
        else if (a == "-spc" && i + 1 < argc) spc = argv[++i];
        else if (a == "-out" && i + 1 < argc) out = argv[++i];
        else if (a == "-tmax" && i + 1 < argc) { tmax = atof(argv[++i]); have_tmax = 1; }
        else if (a == "-nsteps" && i + 1 < argc) { nsteps = atoll(argv[++i]); }
    }
    try {
        c_Sim sim(par, spc, dir, 0, 1, 99999999);
        sim.set_suppress_warnings(1);
        g_out = fopen((out + "_init.bin").c_str(), "wb");
        if (!g_out) { perror("open"); return 2; }
        dump_problem(sim);
        fclose(g_out);
        std::vector<double> dts;
        HookEOS* hook = new HookEOS(sim.species[0].eos, &sim, &dts);
        hook->stop_after = nsteps;
        sim.species[0].eos = hook;
        const double tmax_deck = sim.t_max;
        if (have_tmax) sim.t_max = tmax;
        sim.execute();
        g_out = fopen((out + "_final.bin").c_str(), "wb");
        if (!g_out) { perror("open"); return 2; }
        putv("dt_sequence", dts);
        put1("t_max", sim.t_max); put1("t_max_deck", tmax_deck);
        dump_state(sim);
        fclose(g_out);
    } catch (std::exception& e) {
        std::cerr << "ref_dump: exception: " << e.what() << std::endl; return 1;
    } catch (...) { std::cerr << "ref_dump: unknown exception" << std::endl; return 1; }
    return 0;
}
