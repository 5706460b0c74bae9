// Host-side description of one 1-D column ("problem"): everything c_Sim's constructor
// (reference init_and_bounds.cpp:26-1050) leaves behind that the per-timestep solver core reads.
// It is the interface between the host front end (frontend.cpp: .par/.spc/.opa readers + init) and the
// device engine (engine.cu).  Field names follow the reference members they mirror (aiolos.h:361-1076).
#pragma once
#include <string>
#include <vector>
#include <map>

namespace aio {

struct SpeciesData {                     // mirrors c_Species (aiolos.h:841-1076)
    std::string speciesname;
    double mass_amu = 1, inv_mass = 0, static_charge = 0, initial_fraction = 1, density_excess = 0;
    int    is_dust_like = 0;
    double cv = 0, gamma_adiabat = 0, degrees_of_freedom = 3;
    double const_T_space = 1, const_opacity = 1;
    int    boundary_left = 3, boundary_right = 3;      // BoundaryType (enum.h:35-37)
    int    USE_WAVE = 0; double WAVE_AMPLITUDE = 0, WAVE_PERIOD = 1;
    double SHOCK_TUBE_UL[3] = {0, 0, 0}, SHOCK_TUBE_UR[3] = {0, 0, 0}, BACKGROUND_U[3] = {0, 0, 0};
    std::vector<double> opacity_avg_solar, opacity_avg_planck, opacity_avg_rosseland;   // per band
    std::vector<double> u1, u2, u3;                   // conserved state, N+2 entries
    std::vector<double> eint_last;                    // primlast.internal_energy
    std::vector<double> prim2_speed_init;             // prim[2].speed seen by the fixed BC (1 entry)
    // tabulated opacities (models T/K, io.cpp:421-566)
    int opa_pgrid_size = 0, opa_tgrid_size = 0;
    std::vector<double> opa_pgrid, opa_tgrid, opa_pgrid_log, opa_tgrid_log;
    std::vector<double> opa_grid_solar_log, opa_grid_planck_log, opa_grid_rosseland_log;
};

struct Problem {                         // mirrors c_Sim (aiolos.h:361-828)
    // sizes / switches
    int num_cells = 0, num_species = 0, num_bands_in = 1, num_bands_out = 1, num_ghosts = 2;
    int order = 2, geometry = 0, problem_number = 0, type_of_grid = 0;
    int use_rad_fluxes = 0, use_tides = 0, use_self_gravity = 0, use_linear_gravity = 0;
    int friction_solver = 0, photochemistry_level = 0, use_collisional_heating = 1;
    int use_drag_predictor_step = 0, do_hydrodynamics = 1, closed_radiative_boundaries = 0;
    int use_planetary_temperature = 0, couple_J_into_T = 1, solve_for_j = 1;
    int radiation_matter_equilibrium_test = 0, use_convective_fluxes = 0, update_coll_frequently = 1;
    int use_avg_temperature = 0, use_total_pressure = 0, init_wind = 0;
    char collision_model = 'P', opacity_model = 'C';
    // time stepping
    double dt = 0, cflfactor = 0.9, t_max = 1, max_timestep_change = 1.1, dt_min_init = 1e-20, dt_max = 1e99;
    double timestep_rad2 = 0, energy_epsilon = 0.01, CFL_break_time = 0, start_hydro_time = -1;
    double output_time = 1e99, output_time_offset = 0, monitor_time = 1;
    // gravity
    double planet_mass = 0, planet_position = 0, planet_semimajor = 1, star_mass = 0, init_star_mass = 0;
    double ramp_star_mass_t0 = 0, ramp_star_mass_t1 = 1, rs = 0, rs_time = 0, rs_at_moment = 0.2;
    double domain_min = 0, domain_max = 0, K_zz_init = 0;
    // friction
    double alpha_collision = 1, alpha_collision_ions = 1, coll_rampup_time = 0, init_coll_factor = 0;
    // radiation
    double radiation_rampup_time = 0, init_radiation_factor = 0, bond_albedo = 0, xi_rad = 2, no_rad_trans = 1;
    double photocooling_expansion = 1, temperature_floor = 0, max_temperature = 9e99;
    double const_opacity_solar_factor = 1, const_opacity_rosseland_factor = 1, const_opacity_planck_factor = 1;
    double UV_star = 0, X_star = 0, R_star = 0;
    std::vector<double> solar_heating_thermal;       // black-body part of the band fluxes (init_and_bounds.cpp:808-826)
    int fluxfile_used = 0;
    double T_star = 0, T_int = 0, T_planet = 0, core_cv = 0, right_extrap_press_multiplier = 1, max_mdot = -1;
    double ion_precision = 1e-12, ion_heating_precision = 1e-12; int ion_maxiter = 128, ion_heating_maxiter = 128;   // C2Ray Brent solves
    double lT_spacing = 1, init_J_factor = -1, init_T_temp = 1, avg_temperature_t0 = 1e99, avg_temperature_t1 = 1e99;
    // grid (init_and_bounds.cpp:421-590)
    std::vector<double> x_i, x_i12, x_iVC, dx, surf, vol, omegaplus, omegaminus;
    std::vector<double> source_pressure_prefactor_left, source_pressure_prefactor_right;
    std::vector<double> phi, K_zz;
    // bands
    std::vector<double> l_i_in, l_i_out, photon_energies, solar_heating, solar_heating_final;
    std::vector<int>    BAND_IS_HIGHENERGY;
    std::vector<double> planck_matrix;               // num_plancks x 2 row-major (helpers.cpp:227)
    int num_plancks = 512;
    std::vector<double> Jrad_FLD, Jrad_init;         // (N+2) x B_out row-major
    std::vector<double> dS_band;                     // (N+2) x B_in row-major
    std::vector<double> highenergy_switch;           // S x B_in
    std::vector<double> friction_coeff_mask, inv_totmasses, resonant_pair_matrix;   // S x S row-major
    std::vector<SpeciesData> species;
    std::string simname, workingdir;
};

typedef std::map<std::string, std::vector<double>> DumpMap;
// aiodump container (oracle/ref_dump.cpp): "<name> <count>\n" + count float64.
bool read_dump(const std::string& path, DumpMap& out, std::string& err);
bool problem_from_dump(const DumpMap& d, Problem& p, std::string& err);
bool problem_field(const Problem& p, const std::string& name, std::vector<double>& out);

// Front end: same three strings as the reference's c_Sim ctor (init_and_bounds.cpp:26).
// overrides: "KEY value" pairs applied on top of the .par file for this column (ensemble sweeps).
bool problem_from_files(const std::string& dir, const std::string& par, const std::string& spc,
                        const std::vector<std::pair<std::string, std::string>>& overrides,
                        Problem& p, std::string& err);

}  // namespace aio
