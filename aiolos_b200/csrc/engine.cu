// Ensemble engine: owns the device state of `ncol` independent columns and advances them by launching the fused
// per-column kernels (hydro.cuh, friction.cuh, rad_simple.cuh, rad_full.cuh) on one CUDA stream.
// One time step of every still-active column = one launch sequence  k_hydro -> [k_friction] -> [k_rad_*],
// i.e. c_Sim::execute()'s loop body (reference advection.cpp:60-348) with one CTA per column.
#include "engine.h"
#include "hooks.h"
#include "common.cuh"
#include "hydro.cuh"
#include "friction.cuh"
#include "rad_simple.cuh"
#include "rad_full.cuh"
#include "rad_full2.h"
#include "photochem.cuh"
#include "cuda_check.h"
#include "step2.cuh"
#include <cmath>
#include <cstring>
#include <sstream>

namespace aio {


struct Engine::Impl {
    DevParams p{};
    std::vector<void*> allocs;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int NT = 256, CPT = 4;
    size_t smem_hydro = 0, smem_rad = 0, smem_fric = 0;
    int nt_full = 0; size_t smem_full = 0;
    bool has_friction = false;
    bool big = false;               // 1024 threads per column and/or work arrays in HBM (long columns): v1 kernels only
    bool hydro_switch = false;      // deck runs (part of) its time with the hydro module off
    bool v2_unavailable = false;    // the ctor had to switch v2 off (shared memory): option kernels=v2 must not switch it back on
    bool v2 = true;                 // second-generation kernels (step2.cuh); option kernels=v1 selects the line-by-line ones
    size_t smem_rad2 = 0, smem_rad3 = 0, smem_step1 = 0;
    // full FLD: 2 = register-tiled LU, four kernels (rad_full2.cuh; default when n <= 24); 0 = first-generation shared-memory
    // LU, one kernel (rad_full.cuh; option full_lu=smem, and n > 24); 1 = its shuffle-based register variant (full_lu=reg)
    int full_mode = 0;
    bool full2_ok = false; size_t smem_full2 = 0, smem_heat2 = 0; int full_ri = 0;
    bool fused = false;             // single-species persistent step kernel (k_step1)
    int step1_ctas = 3;             // resident CTAs per SM the fused kernel is compiled for (3: 80 registers, 2: 128)
    int steps_per_launch = 16;
    long long launches = 0;
    std::vector<ColState> h_cs;
    double* pinned = nullptr; size_t pinned_bytes = 0;
    double* d_mass = nullptr;
    // double-buffered host->device staging of the conserved state (end-to-end path): copy stream + events
    cudaStream_t copy_stream = nullptr; cudaEvent_t ev_up = nullptr, ev_free = nullptr; double* u_stage = nullptr; bool staged = false;
    bool prof_on = false; cudaEvent_t prof_ev[4] = {nullptr, nullptr, nullptr, nullptr}; double prof_ms[3] = {0, 0, 0};

    template <typename T> T* dalloc(size_t n) {
        void* q = nullptr; CK(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T))); allocs.push_back(q);
        CK(cudaMemsetAsync(q, 0, std::max<size_t>(n, 1) * sizeof(T), stream)); return (T*)q;
    }
    template <typename T> T* upload(const std::vector<T>& v) {
        T* d = dalloc<T>(v.size());
        if (!v.empty()) CK(cudaMemcpyAsync(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
        return d;
    }
    double* pin(size_t bytes) {
        if (bytes > pinned_bytes) { if (pinned) cudaFreeHost(pinned); CK(cudaMallocHost(&pinned, bytes)); pinned_bytes = bytes; }
        return pinned;
    }
};

// phi as the first update_mass_and_pot() call of execute() leaves it (reference source.cpp:50-109): cells 2..N are
// recomputed with star_mass and a cubed semi-major axis, cells 0,1,N+1 keep the values of init_grav_pot + BC edits.
static void phi_after_first_update(const Problem& q, std::vector<double>& phi) {
    const double G = 6.678e-8, au = 1.495978707e13;
    double rs_at_moment = q.rs;   // globalTime(0) < rs_time is rejected at create time
    for (int i = 2; i <= q.num_cells; i++) {
        double r = q.x_i12[i], mass = q.planet_mass, ph;
        if (q.use_linear_gravity) ph = -G * mass * std::abs(q.domain_max - r);
        else if (std::abs(r - q.planet_position) > rs_at_moment) ph = -G * mass / std::abs(r - q.planet_position);
        else ph = -G * mass * (pow(r - q.planet_position, 3.) / pow(rs_at_moment, 4.) - 2. * pow(r - q.planet_position, 2.) / pow(rs_at_moment, 3.) + 2. / rs_at_moment);
        if (q.use_tides == 1) {
            double d3 = q.planet_semimajor * au;
            d3 = d3 * d3 * d3;
            ph -= 0.5 * 3. * G * q.star_mass * q.x_i12[i] * q.x_i12[i] / d3;
        }
        phi[i] = ph;
    }
}

static void check_supported(const Problem& q) {
    auto fail = [](const std::string& m) { throw std::runtime_error("unsupported configuration: " + m); };
    if (q.num_species < 1 || q.num_species > MAXS) fail("num_species outside 1.." + std::to_string(MAXS));
    if (q.num_bands_in > MAXB || q.num_bands_out > MAXB) fail("more than 32 radiation bands");
    if (q.num_cells < 6) fail("fewer than 6 cells");
    if (q.use_self_gravity && q.num_species > 1) fail("PARI_SELF_GRAV_SWITCH 1 with several species");
    if (q.rs_time > 0) fail("PARI_SMOOTHING_TIME > 0 (time-dependent potential)");
    if (q.K_zz_init != 0 && (q.use_self_gravity || q.photochemistry_level != 0)) fail("KZZ_INIT != 0 together with self-gravity or photochemistry");
    if (q.photochemistry_level != 0 && q.photochemistry_level != 1) fail("PHOTOCHEM_LEVEL " + std::to_string(q.photochemistry_level) + " (general chemistry)");
    if (q.photochemistry_level == 1) {
        if (q.num_species != 3) fail("PHOTOCHEM_LEVEL 1 needs exactly the three species H, H+, e- (photochem.cpp:447-476)");
        if (q.use_rad_fluxes != 0) fail("PHOTOCHEM_LEVEL 1 together with PARI_USE_RADIATION != 0 (every shipped deck of that kind stops in the reference itself: unpinned)");
        if (!(q.friction_solver > 0 || q.do_hydrodynamics)) fail("PHOTOCHEM_LEVEL 1 without the drag update");
        if (!q.do_hydrodynamics || q.start_hydro_time > 0 || q.use_self_gravity) fail("PHOTOCHEM_LEVEL 1 with DO_HYDRO 0 / START_HYDRO_TIME / self-gravity");
    }
    if (q.use_drag_predictor_step) fail("PARI_SECONDORDER_DRAG 1");
    if (q.use_total_pressure) fail("USE_TOTAL_PRESSURE 1");
    if (q.use_convective_fluxes) fail("USE_CONVECTION 1");
    if (q.use_planetary_temperature && q.use_rad_fluxes == 1 && q.num_bands_in < q.num_bands_out) fail("USE_PLANET_TEMPERATURE 1 with fewer incoming than outgoing bands");
    if (q.radiation_matter_equilibrium_test != 0) fail("RAD_MATTER_EQUI_TEST != 0");
    if (q.use_rad_fluxes < 0 || q.use_rad_fluxes > 2) fail("PARI_USE_RADIATION outside 0..2");
    if (q.use_rad_fluxes == 2 && q.num_bands_out != 1) fail("simple FLD solver needs NUM_BANDS_OUT 1");
    if (q.use_rad_fluxes == 2 && q.num_species > 1 && q.use_collisional_heating && q.num_species > 8) fail("simple FLD + collisional heating with more than 8 species");
    if (q.use_rad_fluxes == 2 && q.num_species > 1 && q.use_collisional_heating && !q.update_coll_frequently) fail("UPDATE_COLL_FREQUENTLY 0");
    if (q.use_rad_fluxes && std::string("CPTKD").find(q.opacity_model) == std::string::npos) fail(std::string("opacity model ") + q.opacity_model);
    // with hydro off the reference calls compute_drag_update only for friction_solver > 0 (advection.cpp:208-212)
    if (q.num_species > 1 && q.friction_solver == 0 && !q.do_hydrodynamics && q.start_hydro_time > 0) fail("FRICTION_SOLVER 0 together with START_HYDRO_TIME");
    if (q.init_J_factor > 1e10) fail("INIT_J_FACTOR > 1e10");
    for (auto& sp : q.species) {
        if ((sp.boundary_left == 0 || sp.boundary_right == 0) && !hooks_compiled())
            fail("user boundary hooks (PARI_BOUND_TYPE 0) need a library built with a problem plugin: make PROBLEM=<file>, INTEGRATION.md");
    }
}

// Everything the kernels take from column 0 (DevParams scalars, species constants, opacity tables, band layout) must be the
// same in every column: a per-column .par override may only change what the engine keeps per column -- the initial state
// (u, eint_last, phi, Jrad_FLD, dS_band), the irradiation (solar_heating_final, i.e. PARI_UVSTAR / PARI_XSTAR / PARI_LUM...),
// the clock (dt, cflfactor, t_max, timestep_rad2) and T_planet.  Returns the name of the first field that differs.
static std::string ensemble_mismatch(const Problem& a, const Problem& b) {
#define CMP(f) if (!(a.f == b.f)) return #f
    CMP(num_ghosts); CMP(order); CMP(geometry); CMP(problem_number); CMP(type_of_grid);
    CMP(use_rad_fluxes); CMP(use_tides); CMP(use_self_gravity); CMP(use_linear_gravity); CMP(friction_solver);
    CMP(photochemistry_level); CMP(use_collisional_heating); CMP(use_drag_predictor_step); CMP(do_hydrodynamics);
    CMP(closed_radiative_boundaries); CMP(use_planetary_temperature); CMP(couple_J_into_T); CMP(solve_for_j);
    CMP(radiation_matter_equilibrium_test); CMP(use_convective_fluxes); CMP(update_coll_frequently); CMP(use_avg_temperature);
    CMP(use_total_pressure); CMP(collision_model); CMP(opacity_model);
    CMP(max_timestep_change); CMP(dt_min_init); CMP(dt_max); CMP(energy_epsilon); CMP(CFL_break_time); CMP(start_hydro_time);
    CMP(planet_position); CMP(planet_semimajor); CMP(star_mass); CMP(rs); CMP(rs_time); CMP(domain_min); CMP(domain_max); CMP(K_zz_init);
    CMP(alpha_collision); CMP(alpha_collision_ions); CMP(coll_rampup_time); CMP(init_coll_factor);
    CMP(radiation_rampup_time); CMP(init_radiation_factor); CMP(bond_albedo); CMP(xi_rad); CMP(no_rad_trans);
    CMP(photocooling_expansion); CMP(temperature_floor); CMP(max_temperature);
    CMP(const_opacity_solar_factor); CMP(const_opacity_rosseland_factor); CMP(const_opacity_planck_factor);
    CMP(R_star); CMP(T_star); CMP(T_int); CMP(core_cv); CMP(right_extrap_press_multiplier); CMP(max_mdot);
    CMP(ion_precision); CMP(ion_heating_precision); CMP(ion_maxiter); CMP(ion_heating_maxiter);
    CMP(lT_spacing); CMP(avg_temperature_t0); CMP(avg_temperature_t1); CMP(num_plancks);
    CMP(x_i12); CMP(dx); CMP(vol); CMP(surf); CMP(l_i_in); CMP(l_i_out); CMP(photon_energies); CMP(planck_matrix);
    CMP(highenergy_switch); CMP(friction_coeff_mask); CMP(inv_totmasses); CMP(resonant_pair_matrix);
    bool mass_matters = a.use_self_gravity != 0;
    for (size_t s = 0; s < a.species.size(); s++) {
        const SpeciesData &x = a.species[s], &y = b.species[s];
#define CMPS(f) if (!(x.f == y.f)) return "species[" + std::to_string(s) + "]." #f
        CMPS(mass_amu); CMPS(static_charge); CMPS(is_dust_like); CMPS(cv); CMPS(gamma_adiabat); CMPS(degrees_of_freedom);
        CMPS(const_T_space); CMPS(const_opacity); CMPS(boundary_left); CMPS(boundary_right); CMPS(USE_WAVE); CMPS(WAVE_AMPLITUDE);
        CMPS(WAVE_PERIOD); CMPS(initial_fraction); CMPS(density_excess);
        CMPS(opacity_avg_solar); CMPS(opacity_avg_planck); CMPS(opacity_avg_rosseland);
        CMPS(opa_pgrid); CMPS(opa_tgrid); CMPS(opa_grid_solar_log); CMPS(opa_grid_planck_log); CMPS(opa_grid_rosseland_log);
        for (int k = 0; k < 3; k++) if (x.SHOCK_TUBE_UL[k] != y.SHOCK_TUBE_UL[k] || x.SHOCK_TUBE_UR[k] != y.SHOCK_TUBE_UR[k] || x.BACKGROUND_U[k] != y.BACKGROUND_U[k])
            return "species[" + std::to_string(s) + "] shock-tube / background state";
#undef CMPS
        if (x.boundary_left == 5 || x.boundary_right == 5) mass_matters = true;      // giantplanet BC reads planet_mass in the kernel
    }
    if (mass_matters) CMP(planet_mass);     // otherwise the mass only shapes phi and the initial state, both kept per column
#undef CMP
    return "";
}

// A constructor that throws (unsupported deck, factor storage that does not fit, ...) must not leave device memory, streams or
// events behind: the caller may retry in the same process with fewer columns.
Engine::Engine(const std::vector<Problem>& cols, int device) : im(new Impl) {
    try { construct(cols, device); }
    catch (...) { destroy(); throw; }
}

void Engine::construct(const std::vector<Problem>& cols, int device) {
    if (cols.empty()) throw std::runtime_error("empty ensemble");
    const Problem& q0 = cols[0];
    check_supported(q0);
    for (auto& q : cols) {
        if (q.num_cells != q0.num_cells || q.num_species != q0.num_species || q.num_bands_in != q0.num_bands_in ||
            q.num_bands_out != q0.num_bands_out || q.order != q0.order || q.x_i != q0.x_i)
            throw std::runtime_error("ensemble columns must share grid, species count and band layout");
        if (&q != &q0) {
            const std::string f = ensemble_mismatch(q0, q);
            if (!f.empty())
                throw std::runtime_error("unsupported configuration: a per-column override changes '" + f + "', which the engine reads from "
                                         "column 0 for the whole ensemble (per-column: initial state, irradiation, dt / t_max, T_planet)");
        }
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        throw std::runtime_error("no CUDA device: aiolos_b200 has no CPU fallback");
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&im->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&im->ev0)); CK(cudaEventCreate(&im->ev1));
    prob0 = q0;
    DevParams& p = im->p;
    const int N = q0.num_cells, N2 = N + 2, NP = (N2 + 31) / 32 * 32, S = q0.num_species, ncol = (int)cols.size();
    p.N = N; p.N2 = N2; p.NP = NP; p.S = S; p.Bin = q0.num_bands_in; p.Bout = q0.num_bands_out; p.ng = q0.num_ghosts; p.ncol = ncol;
    p.order = q0.order; p.geometry = q0.geometry; p.problem_number = q0.problem_number;
    p.use_rad_fluxes = q0.use_rad_fluxes; p.friction_solver = q0.friction_solver;
    p.do_hydrodynamics = q0.do_hydrodynamics; p.start_hydro_time = q0.start_hydro_time;
    p.use_self_gravity = q0.use_self_gravity; p.use_linear_gravity = q0.use_linear_gravity; p.use_tides = q0.use_tides;
    p.planet_position = q0.planet_position; p.rs = q0.rs;
    p.use_collisional_heating = q0.use_collisional_heating; p.photochemistry_level = q0.photochemistry_level;
    p.closed_radiative_boundaries = q0.closed_radiative_boundaries; p.use_planetary_temperature = q0.use_planetary_temperature;
    p.couple_J_into_T = q0.couple_J_into_T; p.solve_for_j = q0.solve_for_j;
    p.radiation_matter_equilibrium_test = q0.radiation_matter_equilibrium_test; p.use_total_pressure = q0.use_total_pressure;
    p.collision_model = q0.collision_model; p.opacity_model = q0.opacity_model;
    p.use_avg_temperature = q0.use_avg_temperature; p.update_coll_frequently = q0.update_coll_frequently;
    p.avg_temperature_t0 = q0.avg_temperature_t0; p.avg_temperature_t1 = q0.avg_temperature_t1;
    // compute_drag_update (source.cpp:233-243): numeric solver whenever S > 1; the analytic one (0) only while hydro is on
    im->has_friction = S > 1 && (q0.friction_solver > 0 || (q0.friction_solver == 0 && q0.do_hydrodynamics));
    p.last_kernel = (q0.photochemistry_level == 1) ? 3 : q0.use_rad_fluxes ? 2 : (im->has_friction ? 1 : 0);
    p.num_he_bands = 0;
    if (q0.photochemistry_level > 0) for (int b = 0; b < q0.num_bands_in; b++) if (q0.l_i_in[b + 1] < 0.09116) p.num_he_bands++;   // init_and_bounds.cpp:390-398
    p.ion_precision = q0.ion_precision; p.ion_heating_precision = q0.ion_heating_precision;
    p.ion_maxiter = q0.ion_maxiter; p.ion_heating_maxiter = q0.ion_heating_maxiter;
    p.max_timestep_change = q0.max_timestep_change; p.dt_min_init = q0.dt_min_init; p.dt_max = q0.dt_max;
    p.energy_epsilon = q0.energy_epsilon; p.CFL_break_time = q0.CFL_break_time;
    p.right_extrap_press_multiplier = q0.right_extrap_press_multiplier; p.max_mdot = q0.max_mdot; p.planet_mass = q0.planet_mass;
    p.alpha_collision = q0.alpha_collision; p.alpha_collision_ions = q0.alpha_collision_ions;
    p.coll_rampup_time = q0.coll_rampup_time; p.init_coll_factor = q0.init_coll_factor;
    p.radiation_rampup_time = q0.radiation_rampup_time; p.init_radiation_factor = q0.init_radiation_factor;
    p.bond_albedo = q0.bond_albedo; p.xi_rad = q0.xi_rad; p.no_rad_trans = q0.no_rad_trans;
    p.photocooling_expansion = q0.photocooling_expansion; p.temperature_floor = q0.temperature_floor; p.max_temperature = q0.max_temperature;
    p.kS_factor = q0.const_opacity_solar_factor; p.kR_factor = q0.const_opacity_rosseland_factor; p.kP_factor = q0.const_opacity_planck_factor;
    p.T_int = q0.T_int; p.core_cv = q0.core_cv; p.lT_spacing = q0.lT_spacing; p.num_plancks = q0.num_plancks;
    p.photon_energy0 = q0.photon_energies.empty() ? 1 : q0.photon_energies[0];
    p.step_limit = 0;

    auto padded = [&](const std::vector<double>& v) { std::vector<double> w(NP, 0.0); for (size_t i = 0; i < v.size() && i < (size_t)NP; i++) w[i] = v[i]; return w; };
    p.x_i = im->upload(padded(q0.x_i)); p.x_i12 = im->upload(padded(q0.x_i12)); p.x_iVC = im->upload(padded(q0.x_iVC));
    p.dx = im->upload(padded(q0.dx)); p.surf = im->upload(padded(q0.surf)); p.vol = im->upload(padded(q0.vol));
    p.omegaplus = im->upload(padded(q0.omegaplus)); p.omegaminus = im->upload(padded(q0.omegaminus));
    p.spp_left = im->upload(padded(q0.source_pressure_prefactor_left)); p.spp_right = im->upload(padded(q0.source_pressure_prefactor_right));
    {   // grid-only factors of the slope limiter and the flux divergence (reconstruct.cpp:74-79, advection.cpp:503)
        std::vector<double> cF(NP, 0.), cB(NP, 0.), ixF(NP, 0.), ixB(NP, 0.), wl(NP, 0.), wr(NP, 0.), iv(NP, 0.);
        for (int i = 1; i <= N; i++) {
            cF[i] = (q0.x_iVC[i + 1] - q0.x_iVC[i]) / (q0.x_i[i] - q0.x_iVC[i]);
            cB[i] = (q0.x_iVC[i] - q0.x_iVC[i - 1]) / (q0.x_iVC[i] - q0.x_i[i - 1]);
            ixF[i] = 1. / (q0.x_iVC[i + 1] - q0.x_iVC[i]);
            ixB[i] = 1. / (q0.x_iVC[i] - q0.x_iVC[i - 1]);
            wl[i] = q0.x_i[i - 1] - q0.x_iVC[i];
            wr[i] = q0.x_i[i] - q0.x_iVC[i];
            iv[i] = 1. / q0.vol[i];
        }
        std::vector<double> ix(NP, 1.);
        for (int i = 0; i < N2; i++) ix[i] = 1. / q0.dx[i];
        p.g_idx = im->upload(ix);
        p.g_cF = im->upload(cF); p.g_cB = im->upload(cB); p.g_idxF = im->upload(ixF); p.g_idxB = im->upload(ixB);
        p.g_wl = im->upload(wl); p.g_wr = im->upload(wr); p.g_ivol = im->upload(iv);
    }
    p.planck_matrix = im->upload(q0.planck_matrix); p.l_i_out = im->upload(q0.l_i_out);
    p.highenergy_switch = im->upload(q0.highenergy_switch);
    p.fric_mask = im->upload(q0.friction_coeff_mask); p.inv_totmasses = im->upload(q0.inv_totmasses);
    p.resonant_pair = im->upload(q0.resonant_pair_matrix);

    // species constants + band-mean opacities with the model factors folded in (update_opacities, opacities.cpp:110-134,328-347)
    std::vector<SpeciesConst> sc(S);
    const int KA = p.Bin + 2 * p.Bout;
    std::vector<double> kap((size_t)S * KA, 0.0);
    for (int s = 0; s < S; s++) {
        const SpeciesData& d = q0.species[s];
        SpeciesConst& c = sc[s];
        memset(&c, 0, sizeof c);
        c.gm1 = 2. / d.degrees_of_freedom; c.cv = d.cv; c.mass = d.mass_amu * 1.66054e-24; c.mass_amu = d.mass_amu; c.inv_mass = d.inv_mass;
        c.gamma_adiabat = d.gamma_adiabat; c.static_charge = d.static_charge; c.const_T_space = d.const_T_space; c.const_opacity = d.const_opacity;
        c.wave_amplitude = d.WAVE_AMPLITUDE; c.wave_period = d.WAVE_PERIOD; c.use_wave = d.USE_WAVE;
        for (int k = 0; k < 3; k++) { c.shock_ul[k] = d.SHOCK_TUBE_UL[k]; c.shock_ur[k] = d.SHOCK_TUBE_UR[k]; c.background_u[k] = d.BACKGROUND_U[k]; }
        c.is_dust_like = d.is_dust_like; c.bc_left = d.boundary_left; c.bc_right = d.boundary_right;
        c.prim2_speed = d.prim2_speed_init.empty() ? 0 : d.prim2_speed_init[0];
        c.initial_fraction = d.initial_fraction; c.density_excess = d.density_excess; c.degrees_of_freedom = d.degrees_of_freedom;
        for (int b = 0; b < p.Bin; b++) {
            double a = b < (int)d.opacity_avg_solar.size() ? d.opacity_avg_solar[b] : 0;
            kap[(size_t)s * KA + b] = (q0.opacity_model == 'P' && d.is_dust_like) ? a : q0.const_opacity_solar_factor * a;
        }
        for (int b = 0; b < p.Bout; b++) {
            double kp, kr;
            if (q0.opacity_model == 'P') {
                double ap = b < (int)d.opacity_avg_planck.size() ? d.opacity_avg_planck[b] : 0;
                double ar = b < (int)d.opacity_avg_rosseland.size() ? d.opacity_avg_rosseland[b] : 0;
                kp = d.is_dust_like ? ap : q0.const_opacity_planck_factor * ap;
                kr = d.is_dust_like ? ar : q0.const_opacity_rosseland_factor * ar;
            } else {
                kp = q0.const_opacity_planck_factor * d.const_opacity;
                kr = q0.const_opacity_rosseland_factor * d.const_opacity;
            }
            kap[(size_t)s * KA + p.Bin + b] = kp;
            kap[(size_t)s * KA + p.Bin + p.Bout + b] = kr;
        }
    }
    p.sp = im->upload(sc);
    p.kappa_avg = im->upload(kap);
    // state-dependent opacity models: tables (T, K) and the per-cell opacity array filled by k_opacity every step
    p.kap_cell = nullptr; p.opa_data = nullptr; p.opa_tab = nullptr; p.T_star = q0.T_star;
    p.R_star = q0.R_star; p.planet_semimajor = q0.planet_semimajor; p.star_mass = q0.star_mass; p.UV_star = q0.UV_star;
    p.X_star = q0.X_star; p.domain_min = q0.domain_min; p.domain_max = q0.domain_max;
    if (q0.use_rad_fluxes && (q0.opacity_model == 'T' || q0.opacity_model == 'K' || q0.opacity_model == 'D')) {
        std::vector<double> flat;
        std::vector<OpaTable> tabs(S);
        for (int s = 0; s < S; s++) {
            const SpeciesData& d = q0.species[s];
            OpaTable& t = tabs[s];
            memset(&t, 0, sizeof t);
            t.np = d.opa_pgrid_size; t.nT = d.opa_tgrid_size; t.inv_mass = d.inv_mass; t.mass_amu = d.mass_amu; t.is_dust_like = d.is_dust_like;
            if (q0.opacity_model != 'D') {
                if (t.np < 2 || t.nT < 2 || (int)d.opa_grid_planck_log.size() != t.np * t.nT * p.Bout || (int)d.opa_grid_solar_log.size() != t.np * t.nT * p.Bin)
                    throw std::runtime_error("unsupported configuration: opacity table of species " + d.speciesname + " is missing or malformed");
                auto put = [&](const std::vector<double>& v) { int off = (int)flat.size(); flat.insert(flat.end(), v.begin(), v.end()); return off; };
                t.off_p = put(d.opa_pgrid); t.off_t = put(d.opa_tgrid); t.off_plog = put(d.opa_pgrid_log);
                t.off_solar = put(d.opa_grid_solar_log); t.off_planck = put(d.opa_grid_planck_log); t.off_ross = put(d.opa_grid_rosseland_log);
            }
        }
        if (flat.empty()) flat.push_back(0.);
        p.opa_data = im->upload(flat);
        p.opa_tab = im->upload(tabs);
        p.kap_cell = im->dalloc<double>((size_t)ncol * S * KA * NP);
    }

    // per-column state
    std::vector<double> hu((size_t)3 * ncol * S * NP, 0.0), he((size_t)ncol * S * NP, 0.0), hphi((size_t)ncol * NP, 0.0);
    std::vector<double> hphis((size_t)ncol * S * NP, 0.0), hJ((size_t)ncol * p.Bout * NP, 0.0), hsol((size_t)ncol * p.Bin, 0.0);
    std::vector<double> hp2((size_t)ncol * S, 0.0), hphiN(ncol, 0.0);
    im->h_cs.assign(ncol, ColState{});
    for (int c = 0; c < ncol; c++) {
        const Problem& q = cols[c];
        std::vector<double> phi = q.phi;
        phi_after_first_update(q, phi);
        for (int j = 0; j < N2; j++) hphi[idx_c(p, c, j)] = phi[j];
        hphiN[c] = phi[N];
        for (int s = 0; s < S; s++) {
            const SpeciesData& d = q.species[s];
            for (int j = 0; j < N2; j++) {
                hu[idx_u(p, 0, c, s, j)] = d.u1[j]; hu[idx_u(p, 1, c, s, j)] = d.u2[j]; hu[idx_u(p, 2, c, s, j)] = d.u3[j];
                he[idx_s(p, c, s, j)] = d.eint_last[j];
                hphis[idx_s(p, c, s, j)] = phi[j];
            }
            for (int j = N2; j < NP; j++) { hu[idx_u(p, 0, c, s, j)] = 1; hu[idx_u(p, 2, c, s, j)] = 1; }
            hp2[(size_t)c * S + s] = d.prim2_speed_init.empty() ? 0 : d.prim2_speed_init[0];
        }
        for (int b = 0; b < p.Bout; b++) for (int j = 0; j < N2; j++) hJ[idx_J(p, c, b, j)] = q.Jrad_FLD[(size_t)j * p.Bout + b];
        for (int b = 0; b < p.Bin; b++) hsol[(size_t)c * p.Bin + b] = q.solar_heating_final[b];
        ColState& st = im->h_cs[c];
        st.t = 0; st.dt = q.dt; st.cflfactor = q.cflfactor; st.timestep_rad2 = q.timestep_rad2; st.t_max = q.t_max;
        st.T_planet = q.T_planet; st.steps = 0; st.status = COL_OK;
    }
    p.u = im->upload(hu); p.eint = im->upload(he); p.phi = im->upload(hphi); p.phis = im->upload(hphis);
    p.J = im->upload(hJ); p.solar_heating_final = im->upload(hsol); p.prim2v = im->upload(hp2); p.phiN = im->upload(hphiN);
    p.dS = im->dalloc<double>((size_t)ncol * S * NP);
    p.kzz = q0.K_zz_init; p.kzzf = (q0.K_zz_init != 0) ? im->dalloc<double>((size_t)ncol * S * NP) : nullptr;
    p.dpm = im->dalloc<double>((size_t)ncol * S * NP); p.dpp = im->dalloc<double>((size_t)ncol * S * NP);
    p.gsrc = im->dalloc<double>((size_t)ncol * S * NP);
    {   // dS_band(N, b) as the constructor left it: the first update_dS copies it into cell N+1 (radiation.cpp:263-265)
        std::vector<double> top((size_t)ncol * p.Bin, 0.0);
        for (int c = 0; c < ncol; c++)
            for (int b = 0; b < p.Bin; b++)
                if (cols[c].dS_band.size() == (size_t)N2 * p.Bin) top[(size_t)c * p.Bin + b] = cols[c].dS_band[(size_t)N * p.Bin + b];
        p.dSband_top = im->upload(top);
    }
    p.Tsave = nullptr; p.cssave = nullptr;
    if (im->has_friction && (p.use_rad_fluxes || p.photochemistry_level == 1)) {
        p.Tsave = im->dalloc<double>((size_t)ncol * S * NP);
        p.cssave = im->dalloc<double>((size_t)ncol * S * NP);
    }
    p.cs = im->upload(im->h_cs);
    full_setup();

    // launch geometry: NT threads per column, CPT cells per thread
    int NT = 256;
    if (N2 <= 32) NT = 32; else if (N2 <= 64) NT = 64; else if (N2 <= 128) NT = 128;
    int CPT = (N2 + NT - 1) / NT;
    if (CPT > 8) { NT = (N2 <= 4096) ? 512 : 1024; CPT = (N2 + NT - 1) / NT; }   // long columns: 512 / 1024 threads, v1 kernels
    if (CPT > 8) throw std::runtime_error("unsupported configuration: more than 8190 cells per column");
    if (CPT == 3) CPT = 4; if (CPT > 4 || NT > 256) CPT = 8;
    im->NT = NT; im->CPT = CPT;
    im->smem_hydro = ((size_t)9 * N2 + 32 + 12 * p.ng) * sizeof(double);
    im->smem_rad = ((size_t)(5 * S + 7) * N2 + 32) * sizeof(double);
    im->smem_fric = 64 * sizeof(double);
    im->smem_rad2 = ((size_t)(5 * S + 7) * N2 + 32 + 4 * NT) * sizeof(double);
    im->smem_rad3 = ((size_t)7 * NT * CPT + 32 + 4 * NT) * sizeof(double);
    im->smem_step1 = ((size_t)9 * NT * CPT + 32 + 12 * p.ng + 64) * sizeof(double);
    im->fused = (S == 1) && !im->has_friction && !p.kap_cell && (p.use_rad_fluxes == 0 || p.use_rad_fluxes == 2);
    for (auto& sp : q0.species) if (sp.boundary_left == 0 || sp.boundary_right == 0) im->fused = false;   // hooks flag errors through p.cs
    im->hydro_switch = (!q0.do_hydrodynamics || q0.start_hydro_time > 0);
    if (im->hydro_switch || q0.use_self_gravity || q0.K_zz_init != 0) { im->v2 = false; im->fused = false; }   // (K_zz: phi_s changes every step)   // DO_HYDRO 0 / START_HYDRO_TIME: only the v1 kernels carry the switch
    int maxsm = 0; CK(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    // work arrays that do not fit in one CTA's shared memory (long columns, or many species in the simple FLD kernel) go to HBM
    p.scratch = nullptr; p.scratch_stride = 0; p.scratch_hydro = p.scratch_rad = 0;
    {
        size_t need = 0;
        if (im->smem_hydro > (size_t)maxsm) { p.scratch_hydro = 1; need = std::max(need, im->smem_hydro / sizeof(double)); im->smem_hydro = 0; }
        if (p.use_rad_fluxes == 2 && im->smem_rad > (size_t)maxsm) { p.scratch_rad = 1; need = std::max(need, im->smem_rad / sizeof(double)); im->smem_rad = 0; }
        if (need) {
            p.scratch_stride = (need + 31) / 32 * 32;
            p.scratch = im->dalloc<double>((size_t)ncol * p.scratch_stride);
        }
    }
    im->big = (NT > 256) || p.scratch != nullptr;
    if (im->big) { im->v2 = false; im->fused = false; }            // only the v1 kernels take 1024 threads / HBM work arrays
#define SETSM(K, B) CK(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(B)))
    SETSM(k_hydro<1>, im->smem_hydro); SETSM(k_hydro<2>, im->smem_hydro); SETSM(k_hydro<4>, im->smem_hydro); SETSM(k_hydro<8>, im->smem_hydro);
    { auto k5 = k_hydro<8, 512>; auto k8 = k_hydro<8, 1024>; SETSM(k5, im->smem_hydro); SETSM(k8, im->smem_hydro); }
    if (!im->big) {
    { auto kf = k_step1<4, 256>; SETSM(kf, im->smem_step1); }
    { auto kf = k_step1<4, 256, 2>; SETSM(kf, im->smem_step1); }
    SETSM(k_step1<1>, im->smem_step1); SETSM(k_step1<2>, im->smem_step1); SETSM(k_step1<4>, im->smem_step1); SETSM(k_step1<8>, im->smem_step1);
    SETSM(k_hydro2<1>, im->smem_hydro); SETSM(k_hydro2<2>, im->smem_hydro); SETSM(k_hydro2<4>, im->smem_hydro); SETSM(k_hydro2<8>, im->smem_hydro);
    if (p.use_rad_fluxes == 2) {
        if (im->smem_rad2 > (size_t)maxsm) { im->v2 = false; im->fused = false; im->v2_unavailable = true; }
        else {
            SETSM(k_rad3<1>, im->smem_rad3); SETSM(k_rad3<2>, im->smem_rad3); SETSM(k_rad3<4>, im->smem_rad3); SETSM(k_rad3<8>, im->smem_rad3);
            SETSM(k_rad2<1>, im->smem_rad2); SETSM(k_rad2<2>, im->smem_rad2); SETSM(k_rad2<4>, im->smem_rad2); SETSM(k_rad2<8>, im->smem_rad2);
        }
    }
    }
    if (p.use_rad_fluxes == 2) {
        SETSM(k_rad_simple<1>, im->smem_rad); SETSM(k_rad_simple<2>, im->smem_rad); SETSM(k_rad_simple<4>, im->smem_rad); SETSM(k_rad_simple<8>, im->smem_rad);
        { auto k5 = k_rad_simple<8, 512>; auto k8 = k_rad_simple<8, 1024>; SETSM(k5, im->smem_rad); SETSM(k8, im->smem_rad); }
    }
#undef SETSM
    AIO_LAUNCH(k_prime, ncol, 256, 0, im->stream, p);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(im->stream));
}

Engine::~Engine() { destroy(); }

void Engine::destroy() {
    if (!im) return;
    for (void* q : im->allocs) cudaFree(q);
    if (im->pinned) cudaFreeHost(im->pinned);
    if (im->ev0) cudaEventDestroy(im->ev0);
    if (im->ev1) cudaEventDestroy(im->ev1);
    for (int k = 0; k < 4; k++) if (im->prof_ev[k]) cudaEventDestroy(im->prof_ev[k]);
    if (im->copy_stream) { cudaStreamDestroy(im->copy_stream); cudaEventDestroy(im->ev_up); cudaEventDestroy(im->ev_free); }
    if (im->stream) cudaStreamDestroy(im->stream);
    delete im;
    im = nullptr;
}

int Engine::ncol() const { return im->p.ncol; }
int Engine::num_cells() const { return im->p.N; }
int Engine::num_species() const { return im->p.S; }
int Engine::num_bands_out() const { return im->p.Bout; }
int Engine::num_bands_in() const { return im->p.Bin; }
long long Engine::kernel_launches() const { return im->launches; }
void* Engine::stream() const { return (void*)im->stream; }

void Engine::launch_step() {
    DevParams& p = im->p;
    const int ncol = p.ncol, NT = im->NT;
    const bool prof = im->prof_on;
    auto mark = [&](int k) { if (prof) CK(cudaEventRecord(im->prof_ev[k], im->stream)); };
    mark(0);
    if (im->v2) {
        switch (im->CPT) {
        case 1: AIO_LAUNCH(k_hydro2<1>, ncol, NT, im->smem_hydro, im->stream, p); break;
        case 2: AIO_LAUNCH(k_hydro2<2>, ncol, NT, im->smem_hydro, im->stream, p); break;
        case 4: AIO_LAUNCH(k_hydro2<4>, ncol, NT, im->smem_hydro, im->stream, p); break;
        default: AIO_LAUNCH(k_hydro2<8>, ncol, NT, im->smem_hydro, im->stream, p); break;
        }
    } else if (NT > 256) {               // long columns: CPT is 8 by construction (N2 > 8 * 256)
        if (NT == 512) { auto k = k_hydro<8, 512>; AIO_LAUNCH(k, ncol, NT, im->smem_hydro, im->stream, p); }
        else { auto k = k_hydro<8, 1024>; AIO_LAUNCH(k, ncol, NT, im->smem_hydro, im->stream, p); }
    } else
    switch (im->CPT) {
    case 1: AIO_LAUNCH(k_hydro<1>, ncol, NT, im->smem_hydro, im->stream, p); break;
    case 2: AIO_LAUNCH(k_hydro<2>, ncol, NT, im->smem_hydro, im->stream, p); break;
    case 4: AIO_LAUNCH(k_hydro<4>, ncol, NT, im->smem_hydro, im->stream, p); break;
    default: AIO_LAUNCH(k_hydro<8>, ncol, NT, im->smem_hydro, im->stream, p); break;
    }
    im->launches++;
    mark(1);
    if (im->has_friction) { launch_friction(); im->launches++; }
    mark(2);
    if (p.kap_cell) { AIO_LAUNCH(k_opacity, ncol, 256, 0, im->stream, p); im->launches++; }
    if (p.use_rad_fluxes == 2 && im->v2 && p.S == 1 && !p.Tsave && !p.kap_cell) {
        switch (im->CPT) {
        case 1: AIO_LAUNCH(k_rad3<1>, ncol, NT, im->smem_rad3, im->stream, p); break;
        case 2: AIO_LAUNCH(k_rad3<2>, ncol, NT, im->smem_rad3, im->stream, p); break;
        case 4: AIO_LAUNCH(k_rad3<4>, ncol, NT, im->smem_rad3, im->stream, p); break;
        default: AIO_LAUNCH(k_rad3<8>, ncol, NT, im->smem_rad3, im->stream, p); break;
        }
        im->launches++;
    } else if (p.use_rad_fluxes == 2 && im->v2 && !(p.use_collisional_heating && p.S > 1) && !p.kap_cell) {
        switch (im->CPT) {
        case 1: AIO_LAUNCH(k_rad2<1>, ncol, NT, im->smem_rad2, im->stream, p); break;
        case 2: AIO_LAUNCH(k_rad2<2>, ncol, NT, im->smem_rad2, im->stream, p); break;
        case 4: AIO_LAUNCH(k_rad2<4>, ncol, NT, im->smem_rad2, im->stream, p); break;
        default: AIO_LAUNCH(k_rad2<8>, ncol, NT, im->smem_rad2, im->stream, p); break;
        }
        im->launches++;
    } else if (p.use_rad_fluxes == 2 && NT > 256) {
        if (NT == 512) { auto k = k_rad_simple<8, 512>; AIO_LAUNCH(k, ncol, NT, im->smem_rad, im->stream, p); }
        else { auto k = k_rad_simple<8, 1024>; AIO_LAUNCH(k, ncol, NT, im->smem_rad, im->stream, p); }
        im->launches++;
    } else if (p.use_rad_fluxes == 2) {
        switch (im->CPT) {
        case 1: AIO_LAUNCH(k_rad_simple<1>, ncol, NT, im->smem_rad, im->stream, p); break;
        case 2: AIO_LAUNCH(k_rad_simple<2>, ncol, NT, im->smem_rad, im->stream, p); break;
        case 4: AIO_LAUNCH(k_rad_simple<4>, ncol, NT, im->smem_rad, im->stream, p); break;
        default: AIO_LAUNCH(k_rad_simple<8>, ncol, NT, im->smem_rad, im->stream, p); break;
        }
        im->launches++;
    } else if (p.use_rad_fluxes == 1) { launch_full(); im->launches++; }
    if (p.photochemistry_level == 1) { AIO_LAUNCH(k_photochem, (ncol + 4 * PHOTOCHEM_LANES - 1) / (4 * PHOTOCHEM_LANES), 128, 0, im->stream, p); im->launches++; }
    mark(3);
    if (prof) {
        CK(cudaEventSynchronize(im->prof_ev[3]));
        for (int k = 0; k < 3; k++) { float ms = 0; CK(cudaEventElapsedTime(&ms, im->prof_ev[k], im->prof_ev[k + 1])); im->prof_ms[k] += ms; }
    }
}

// Per-kernel device time, measured live with CUDA events on the ensemble's stream: runs `nsteps` ordinary steps with an
// event between the launches and returns the average milliseconds of {hydro, friction, radiation} per step.
void Engine::kernel_times(long long nsteps, double* out3) {
    if (!im->prof_ev[0]) for (int k = 0; k < 4; k++) CK(cudaEventCreate(&im->prof_ev[k]));
    for (int k = 0; k < 3; k++) im->prof_ms[k] = 0;
    im->p.step_limit += nsteps;
    if (im->fused && im->v2) {          // one kernel does the whole step: report it in slot 0
        CK(cudaEventRecord(im->prof_ev[0], im->stream));
        launch_fused(nsteps);
        CK(cudaEventRecord(im->prof_ev[1], im->stream));
        CK(cudaEventSynchronize(im->prof_ev[1]));
        float ms = 0; CK(cudaEventElapsedTime(&ms, im->prof_ev[0], im->prof_ev[1]));
        out3[0] = ms / (double)nsteps; out3[1] = 0; out3[2] = 0;
        return;
    }
    im->prof_on = true;
    for (long long k = 0; k < nsteps; k++) launch_step();
    im->prof_on = false;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(im->stream));
    for (int k = 0; k < 3; k++) out3[k] = nsteps > 0 ? im->prof_ms[k] / (double)nsteps : 0.0;
}

// Single-species decks: one launch of the persistent kernel advances every column by up to steps_per_launch steps.
void Engine::launch_fused(long long nsteps) {
    DevParams& p = im->p;
    const int ncol = p.ncol, NT = im->NT;
    long long left = nsteps;
    while (left > 0) {
        p.steps_per_launch = (int)std::min<long long>(left, im->steps_per_launch);
        switch (im->CPT) {
        case 1: AIO_LAUNCH(k_step1<1>, ncol, NT, im->smem_step1, im->stream, p); break;
        case 2: AIO_LAUNCH(k_step1<2>, ncol, NT, im->smem_step1, im->stream, p); break;
        case 4:
            if (NT == 256 && im->step1_ctas == 2) { auto kf = k_step1<4, 256, 2>; AIO_LAUNCH(kf, ncol, NT, im->smem_step1, im->stream, p); }
            else if (NT == 256) { auto kf = k_step1<4, 256>; AIO_LAUNCH(kf, ncol, NT, im->smem_step1, im->stream, p); }
            else AIO_LAUNCH(k_step1<4>, ncol, NT, im->smem_step1, im->stream, p);
            break;
        default: AIO_LAUNCH(k_step1<8>, ncol, NT, im->smem_step1, im->stream, p); break;
        }
        im->launches++;
        left -= p.steps_per_launch;
    }
}

// Advance every column by up to `nsteps` further steps (columns that reach their t_max or crash stop earlier).
double Engine::advance(long long nsteps, bool timed) {
    DevParams& p = im->p;
    p.step_limit += nsteps;     // columns are bounded by their own step counters: min(steps so far) + nsteps would
                                // desynchronise the ensemble; a global ceiling keeps "advance(k)" = k steps for all
    // step_limit is an absolute ceiling on ColState.steps
    if (timed) CK(cudaEventRecord(im->ev0, im->stream));
    if (im->fused && im->v2) launch_fused(nsteps);
    else for (long long k = 0; k < nsteps; k++) launch_step();
    CK(cudaGetLastError());
    float ms = 0;
    if (timed) { CK(cudaEventRecord(im->ev1, im->stream)); CK(cudaEventSynchronize(im->ev1)); CK(cudaEventElapsedTime(&ms, im->ev0, im->ev1)); }
    else CK(cudaStreamSynchronize(im->stream));
    return ms;
}

// Run until every column has finished (t >= t_max or crashed), polling the per-column clocks every `chunk` steps.
// Returns the steps actually taken: the largest per-column step count since entry (the reference's `steps` of execute()).
long long Engine::run_to_tmax(long long chunk, long long max_total) {
    fetch_states();
    std::vector<long long> steps0(im->h_cs.size());
    for (size_t c = 0; c < im->h_cs.size(); c++) steps0[c] = im->h_cs[c].steps;
    long long done = 0, taken = 0;
    while (done < max_total) {
        const long long k = std::min(chunk, max_total - done);     // never past max_total: execute() stops exactly at maxsteps
        advance(k, false);
        done += k;
        fetch_states();
        bool any = false;
        taken = 0;
        for (size_t c = 0; c < im->h_cs.size(); c++) {
            const ColState& cs = im->h_cs[c];
            taken = std::max(taken, (long long)cs.steps - steps0[c]);
            if (cs.status == COL_OK && cs.t < cs.t_max) any = true;
        }
        if (!any) break;
    }
    return taken;
}

void Engine::fetch_states() {
    CK(cudaMemcpyAsync(im->h_cs.data(), im->p.cs, im->h_cs.size() * sizeof(ColState), cudaMemcpyDeviceToHost, im->stream));
    CK(cudaStreamSynchronize(im->stream));
}
void Engine::push_states() {
    CK(cudaMemcpyAsync(im->p.cs, im->h_cs.data(), im->h_cs.size() * sizeof(ColState), cudaMemcpyHostToDevice, im->stream));
    CK(cudaStreamSynchronize(im->stream));
}
void Engine::set_tmax(int col, double t_max) {
    fetch_states();
    for (int c = 0; c < im->p.ncol; c++) if (col < 0 || c == col) im->h_cs[c].t_max = t_max;
    push_states();
}
void Engine::clamp_tmax(double t_stop) {
    fetch_states();
    for (auto& c : im->h_cs) if (t_stop < c.t_max) c.t_max = t_stop;
    push_states();
}
void Engine::set_option(const std::string& key, const std::string& value) {
    if (key == "tridiag") {
        if (value == "thomas") im->p.tridiag_serial = 1;
        else if (value == "pcr") im->p.tridiag_serial = 0;
        else throw std::invalid_argument("option tridiag: expected 'pcr' or 'thomas'");
    } else if (key == "steps_per_launch") {
        const int k = atoi(value.c_str());
        if (k < 1 || k > 100000) throw std::invalid_argument("option steps_per_launch: expected 1..100000");
        im->steps_per_launch = k;
    } else if (key == "full_lu") {
        if (value == "smem") im->full_mode = 0;
        else if (value == "reg") im->full_mode = 1;
        else if (value == "tile") { if (!im->full2_ok) throw std::invalid_argument("option full_lu=tile: needs PARI_USE_RADIATION 1 with at most 24 unknowns per cell"); im->full_mode = 2; }
        else throw std::invalid_argument("option full_lu: expected 'tile', 'reg' or 'smem'");
    } else if (key == "full_heat") {
        if (value == "column") im->smem_heat2 = 0;           // k_full_heat (warp per column) instead of k_full_heat2
        else throw std::invalid_argument("option full_heat: expected 'column'");
    } else if (key == "step1_ctas") {
        if (value != "2" && value != "3") throw std::invalid_argument("option step1_ctas: expected 2 or 3");
        im->step1_ctas = std::stoi(value);
    } else if (key == "carveout") {          // shared-memory share of the SM's L1/shared storage for the fused kernel, in percent
        const int pct = std::stoi(value);
        if (pct < 0 || pct > 100) throw std::invalid_argument("option carveout: expected 0..100");
        { auto kf = k_step1<4, 256>; CK(cudaFuncSetAttribute(kf, cudaFuncAttributePreferredSharedMemoryCarveout, pct)); }
        { auto kf = k_step1<4, 256, 2>; CK(cudaFuncSetAttribute(kf, cudaFuncAttributePreferredSharedMemoryCarveout, pct)); }
    } else if (key == "fused") {             // 0: per-phase kernels instead of the persistent step kernel; 1 is the default where it applies
        if (value == "0") im->fused = false;
        else if (value == "1") { if (!im->fused) throw std::invalid_argument("option fused=1: the persistent step kernel does not apply to this deck (or was switched off)"); }
        else throw std::invalid_argument("option fused: expected 0 or 1");
    } else if (key == "kernels") {
        if (value != "v1" && value != "v2") throw std::invalid_argument("option kernels: expected 'v1' or 'v2'");
        if (total_steps() != 0) throw std::invalid_argument("option kernels must be set before the first step");
        if (value == "v1") im->v2 = false;
        else if (value == "v2") { if (im->v2_unavailable) throw std::invalid_argument("option kernels=v2: the second-generation radiation kernel does not fit this deck's shared memory"); if (im->big) throw std::invalid_argument("option kernels=v2: long columns run the v1 kernels"); if (im->hydro_switch || im->p.use_self_gravity || im->p.kzzf) throw std::invalid_argument("option kernels=v2: decks with DO_HYDRO 0 / START_HYDRO_TIME / self-gravity / KZZ_INIT run the v1 kernels"); im->v2 = true; }
    } else throw std::invalid_argument("unknown option '" + key + "'");
}
long long Engine::total_steps() {
    fetch_states();
    long long n = 0;
    for (auto& c : im->h_cs) n += c.steps;
    return n;
}
void Engine::status(int col, double* t, double* dt, long long* steps, int* state) {
    fetch_states();
    const ColState& c = im->h_cs.at(col);
    if (t) *t = c.t; if (dt) *dt = c.dt; if (steps) *steps = c.steps; if (state) *state = c.status;
}
void Engine::status_all(double* t, double* dt, long long* steps, int* state) {
    fetch_states();
    for (int c = 0; c < im->p.ncol; c++) {
        const ColState& s = im->h_cs[c];
        if (t) t[c] = s.t; if (dt) dt[c] = s.dt; if (steps) steps[c] = s.steps; if (state) state[c] = s.status;
    }
}
double Engine::col_scalar(int col, int which) {
    fetch_states();
    const ColState& c = im->h_cs.at(col);
    switch (which) { case 0: return c.cflfactor; case 1: return c.timestep_rad2; case 2: return c.cfl_step;
                     case 3: return c.minstep; case 4: return c.maxde; case 5: return c.t_max; default: return 0; }
}

// field ids: include/aiolos_b200.h (AIO_F_*); names follow the reference's members
void Engine::get_field(int col, int species, int field, double* dst, int n) {
    DevParams& p = im->p;
    if (col < 0 || col >= p.ncol || n > p.N2) throw std::invalid_argument("get_field: bad column or length");
    const bool per_species = (field <= 3) || (field >= 5 && field <= 9);
    if (per_species && (species < 0 || species >= p.S)) throw std::invalid_argument("get_field: bad species");
    auto fetch = [&](const double* src, double* d) {
        CK(cudaMemcpyAsync(d, src, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, im->stream));
        CK(cudaStreamSynchronize(im->stream));
    };
    if (field >= 0 && field <= 2) fetch(p.u + idx_u(p, field, col, species, 0), dst);
    else if (field == 3) fetch(p.eint + idx_s(p, col, species, 0), dst);
    else if (field == 4) fetch(p.phi + idx_c(p, col, 0), dst);
    else if (field == 5) fetch(p.phis + idx_s(p, col, species, 0), dst);
    else if (field == 6) fetch(p.dS + idx_s(p, col, species, 0), dst);
    else if (field >= 7 && field <= 9) {            // primitives derived from u with the EOS of eos.cpp:19-63
        std::vector<double> a(n), b(n), c(n);
        fetch(p.u + idx_u(p, 0, col, species, 0), a.data());
        fetch(p.u + idx_u(p, 1, col, species, 0), b.data());
        fetch(p.u + idx_u(p, 2, col, species, 0), c.data());
        const SpeciesData& d = prob0.species[species];
        const double gm1 = 2. / d.degrees_of_freedom;
        for (int j = 0; j < n; j++) {
            const double v = b[j] / a[j], pr = gm1 * (c[j] - 0.5 * b[j] * v);
            dst[j] = (field == 9) ? v : (field == 8) ? pr : ((pr / a[j]) / gm1) / d.cv;
        }
    }
    else if (field >= 16 && field < 16 + p.Bout) fetch(p.J + idx_J(p, col, field - 16, 0), dst);
    else throw std::invalid_argument("get_field: unknown field id");
}

// Replace the conserved state of all columns from / copy it to host buffers laid out [3][ncol][S][N+2] (e2e path).
void Engine::upload_u(const double* src) {
    DevParams& p = im->p;
    const size_t rows = (size_t)3 * p.ncol * p.S;
    CK(cudaMemcpy2DAsync(p.u, (size_t)p.NP * sizeof(double), src, (size_t)p.N2 * sizeof(double), (size_t)p.N2 * sizeof(double), rows,
                         cudaMemcpyHostToDevice, im->stream));
}
// Asynchronous variant for pipelined callers: the copy runs on a second stream into a staging buffer while the compute stream
// is still stepping; commit_upload() makes the staged state the current one (pointer swap, ordered by events).
// Pipelined upload.  The host buffer is unpadded ([3][ncol][S][N+2]); a pitched 2-D copy of 12 288 rows of 7 992 bytes runs
// at ~80 % of the PCIe rate, so the copy stream moves the buffer as ONE contiguous transfer into a raw staging area and the
// re-pitching into the padded device layout is a device kernel on the compute stream at commit time (98 MB at HBM speed).
struct RepitchArgs { const double* src; double* dst; int N2, NP; };
__global__ void k_repitch(RepitchArgs a) {
    const size_t row = blockIdx.x;
    const double* s = a.src + row * a.N2;
    double* d = a.dst + row * a.NP;
    for (int j = threadIdx.x; j < a.N2; j += blockDim.x) d[j] = s[j];
}
void Engine::upload_u_async(const double* src) {
    DevParams& p = im->p;
    const size_t rows = (size_t)3 * p.ncol * p.S;
    if (!im->copy_stream) {
        CK(cudaStreamCreateWithFlags(&im->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&im->ev_up)); CK(cudaEventCreate(&im->ev_free));
        im->u_stage = im->dalloc<double>(rows * p.N2);
        CK(cudaStreamSynchronize(im->stream));
        CK(cudaEventRecord(im->ev_free, im->stream));
    }
    CK(cudaStreamWaitEvent(im->copy_stream, im->ev_free, 0));      // the previous commit has finished reading the staging area
    CK(cudaMemcpyAsync(im->u_stage, src, rows * p.N2 * sizeof(double), cudaMemcpyHostToDevice, im->copy_stream));
    CK(cudaEventRecord(im->ev_up, im->copy_stream));
    im->staged = true;
}
void Engine::commit_upload() {
    if (!im->staged) throw std::invalid_argument("commit_upload without a pending upload_u_async");
    DevParams& p = im->p;
    const size_t rows = (size_t)3 * p.ncol * p.S;
    CK(cudaStreamWaitEvent(im->stream, im->ev_up, 0));
    RepitchArgs a{im->u_stage, p.u, p.N2, p.NP};
    AIO_LAUNCH(k_repitch, (unsigned)rows, 256, 0, im->stream, a);
    CK(cudaGetLastError());
    CK(cudaEventRecord(im->ev_free, im->stream));
    im->staged = false;
}
void Engine::download_u(double* dst) {
    DevParams& p = im->p;
    const size_t rows = (size_t)3 * p.ncol * p.S;
    CK(cudaMemcpy2DAsync(dst, (size_t)p.N2 * sizeof(double), p.u, (size_t)p.NP * sizeof(double), (size_t)p.N2 * sizeof(double), rows,
                         cudaMemcpyDeviceToHost, im->stream));
    CK(cudaStreamSynchronize(im->stream));
}
// Mass per column: a small per-step result for the end-to-end path to read back.
__device__ inline void k_column_mass(const DevParams& p, double* out) {
    __shared__ double red[32];
    const int col = blockIdx.x;
    double acc = 0;
    for (int s = 0; s < p.S; s++)
        for (int j = threadIdx.x; j < p.N2; j += blockDim.x) acc += p.u[idx_u(p, 0, col, s, j)] * p.vol[j];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0; for (unsigned w = 0; w < (blockDim.x + 31) / 32; w++) t += red[w]; out[col] = t; }
}
struct MassArgs { DevParams p; double* out; };
__global__ void k_column_mass_w(MassArgs a) { k_column_mass(a.p, a.out); }
void Engine::column_mass(double* m) {
    DevParams& p = im->p;
    if (!im->d_mass) im->d_mass = im->dalloc<double>(p.ncol);
    MassArgs a{p, im->d_mass};
    AIO_LAUNCH(k_column_mass_w, p.ncol, 256, 0, im->stream, a);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(m, im->d_mass, (size_t)p.ncol * sizeof(double), cudaMemcpyDeviceToHost, im->stream));
    CK(cudaStreamSynchronize(im->stream));
}
void Engine::sync() { CK(cudaStreamSynchronize(im->stream)); }

}  // namespace aio

#include "engine_more.inl"
#include "hooks_glue.inl"
