// Device-side data model shared by all kernels of the per-timestep solver core.
//
// Layout in HBM (structure-of-arrays, FP64, cell index fastest so that a warp touches 256 contiguous bytes):
//   u     [3][ncol][S][NP]     conserved rho, rho*v, E            (reference: std::vector<AOS> c_Species::u, aiolos.h:881)
//   eint  [ncol][S][NP]        prim.internal_energy at the end of the previous step (primlast, advection.cpp:88-91)
//   phi   [ncol][NP]           gravitational potential incl. ghost edits by the BCs (aiolos.h:533)
//   phis  [ncol][S][NP]        species potential phi_s = phi*K_zzf snapshot (source.cpp:207-209)
//   J     [ncol][Bout][NP]     mean intensity per outgoing band (Jrad_FLD, aiolos.h:667)
//   dS    [ncol][S][NP]        stellar heating per species (c_Species::dS)
//   geometry arrays [NP]       shared by all columns of the ensemble (same grid), read through the read-only cache
// NP = N+2 rounded up to 32 entries.  One ColState per column carries its own clock: columns of an ensemble
// desynchronise (per-column CFL dt), nothing in the step couples two columns.
#pragma once
#include "cuda_compat.h"
#include <cstdint>
#include <cmath>

namespace aio {

constexpr int MAXS = 16;        // species per column supported by the generic kernels
constexpr int MAXB = 32;        // radiation bands (in or out)

enum ColStatus : int { COL_OK = 0, COL_NEG_T = 1, COL_NEG_J = 2, COL_NAN = 3,
                       COL_HOOK_RANGE = 4 };   // a plugin boundary hook touched a cell outside its boundary window

struct ColState {               // per-column clock and step-control scalars (c_Sim members, aiolos.h:458-476)
    double t;                   // globalTime
    double dt;                  // last used dt
    double cflfactor;           // mutated to 0.9 after CFL_break_time (helpers.cpp:74-75)
    double timestep_rad2;       // energy-change limiter (helpers.cpp:47)
    double cfl_step;
    double minstep;             // max over (s,i) of |v|/dx + cs/dx, produced by the previous step's tail
    double maxde;               // max over (s,i) of |de|/e
    double t_max;
    double T_planet;
    double crash_value;
    long long steps;
    int status;                 // ColStatus
    int crash_species_or_band;
    int in_step;                // set by the step head: this column is being advanced by the current launch sequence
    int pad;
};

struct SpeciesConst {           // per-species constants (c_Species + IdealGas_EOS, eos.h:28-47)
    double gm1;                 // _gamma_m1 = 2/dof
    double cv;
    double mass;                // mass_amu*amu
    double mass_amu;
    double inv_mass;
    double gamma_adiabat;
    double static_charge;
    double const_T_space;
    double const_opacity;
    double wave_amplitude, wave_period;
    double shock_ul[3], shock_ur[3], background_u[3];
    double prim2_speed;         // prim[2].speed as seen by the 'fixed' BC
    double initial_fraction, density_excess, degrees_of_freedom;   // read by plugin hooks only
    int is_dust_like;
    int bc_left, bc_right;      // BoundaryType
    int use_wave;
};

struct OpaTable {               // one species' .aiopa table (io.cpp:421-566): offsets into DevParams::opa_data
    int np, nT;
    int off_p, off_t, off_plog, off_solar, off_planck, off_ross;
    double inv_mass, mass_amu;
    int is_dust_like, pad;
};

struct DevParams {
    // sizes
    int N, N2, NP, S, Bin, Bout, ng, ncol;
    int order, geometry, problem_number;
    // switches
    int use_rad_fluxes, friction_solver, use_collisional_heating, photochemistry_level;
    int closed_radiative_boundaries, use_planetary_temperature, couple_J_into_T, solve_for_j;
    int radiation_matter_equilibrium_test, use_total_pressure, collision_model, opacity_model;
    int use_avg_temperature, update_coll_frequently;
    int use_self_gravity, use_linear_gravity, use_tides;   // update_mass_and_pot() per step (source.cpp:50-109) when self-gravity is on
    double planet_position, rs;
    int do_hydrodynamics;       // DO_HYDRO; 0 = no hydro stages until t > start_hydro_time (advection.cpp:62-64,139,170-200)
    double start_hydro_time;
    double avg_temperature_t0, avg_temperature_t1;
    int last_kernel;            // 0 hydro, 1 friction, 2 radiation, 3 photochemistry: who runs the step tail
    int tridiag_serial;         // 1: scalar J system solved in the reference's Thomas order on one lane (validation mode)
    // scalars
    double max_timestep_change, dt_min_init, dt_max, energy_epsilon, CFL_break_time;
    double right_extrap_press_multiplier, max_mdot, planet_mass;
    double alpha_collision, alpha_collision_ions, coll_rampup_time, init_coll_factor;
    double radiation_rampup_time, init_radiation_factor, bond_albedo, xi_rad, no_rad_trans;
    double photocooling_expansion, temperature_floor, max_temperature;
    double kS_factor, kR_factor, kP_factor;       // const_opacity_{solar,rosseland,planck}_factor
    double T_int, core_cv, lT_spacing, photon_energy0;
    int num_plancks, h2_index;
    // C2Ray photochemistry (photochem.cuh): number of ionising bands, Brent tolerances / iteration caps (ION_PRECISION, ...)
    int num_he_bands, ion_maxiter, ion_heating_maxiter;
    double ion_precision, ion_heating_precision;
    // shared geometry [NP]
    const double *x_i, *x_i12, *x_iVC, *dx, *surf, *vol, *omegaplus, *omegaminus, *spp_left, *spp_right;
    // shared tables
    const double *planck_matrix;        // [num_plancks][2]
    const double *l_i_out;              // [Bout+1]
    const double *kappa_avg;            // [S][Bin + 2*Bout]: solar | planck | rosseland band means
    const double *highenergy_switch;    // [S][Bin]
    const double *fric_mask, *inv_totmasses, *resonant_pair;   // [S][S]
    const SpeciesConst* sp;             // [S]
    // per-column state
    double *u, *eint, *phi, *phis, *J, *dS;
    const double *phiN;                 // [ncol] phi[num_cells] as update_mass_and_pot() recomputes it every step (see phi_snapshot)
    double *dSband_top;                 // [ncol][Bin] dS_band(N,b) of the previous step (radiation.cpp:263-265)
    // v2 kernels (step2.cuh): tabulated slope-limiter factors / inverse volumes [NP], shared by all columns, and the static
    // per-column hydrostatic offsets and gravity source coefficient [ncol][S][NP]
    const double *g_cF, *g_cB, *g_idxF, *g_idxB, *g_wl, *g_wr, *g_ivol, *g_idx;
    double *dpm, *dpp, *gsrc;
    double *kzzf; double kzz;           // KZZ_INIT != 0: K_zzf per species and cell [ncol][S][NP] (source.cpp:174-219), refreshed every step; else null
    // per-cell opacities for the models that depend on the local state (T, K, D): [ncol][S][Bin+2*Bout][NP], filled by k_opacity
    // before the radiation kernel (the reference's update_opacities pass); null for the constant models C and P
    double *kap_cell;
    const double *opa_data;             // tabulated opacities of all species, see OpaTable
    const struct OpaTable *opa_tab;     // [S]
    double T_star;
    double R_star, planet_semimajor, star_mass, UV_star, X_star, domain_min, domain_max;   // read by plugin hooks only
    double *Xstore;                     // full FLD: [ncol][N2][n][Bout+1]  X_i = S_i^-1 u_i and S_i^-1 rhs_i (rad_full.cuh)
    // second-generation full FLD (rad_full2.cuh): per-cell coefficient stream [ncol][N2][NC], raw temperature solution
    // [ncol][S][NP], and the surface cell's D_core [ncol][Bout] / incoming flux [ncol][Bin] for the planetary slab model
    double *fcoef, *Tsol, *dcore, *ssurf;
    double *Tsave, *cssave;             // [ncol][S][NP] prim.temperature / prim.sound_speed as left by the post-hydro
                                        // compute_pressure: friction changes u but not these (source.cpp:485-505), and the
                                        // radiation solvers and the next CFL estimate read the stale values
    double *prim2v;                     // [ncol][S]  prim[2].speed as last computed (fixed BC, init_and_bounds.cpp:1856)
    const double *solar_heating_final;  // [ncol][Bin]
    ColState* cs;                       // [ncol]
    // long columns (or many species): the per-column work arrays of the v1 kernels live in HBM instead of shared memory
    double* scratch; size_t scratch_stride; int scratch_hydro, scratch_rad;
    int steps_per_launch;               // fused persistent kernel (k_step1): time steps one launch may advance a column
    long long step_limit;               // columns stop when steps reaches this
};

__host__ __device__ inline size_t idx_u(const DevParams& p, int comp, int col, int s, int j) {
    return (((size_t)comp * p.ncol + col) * p.S + s) * p.NP + j;
}
__host__ __device__ inline size_t idx_s(const DevParams& p, int col, int s, int j) {
    return ((size_t)col * p.S + s) * p.NP + j;
}
__host__ __device__ inline size_t idx_c(const DevParams& p, int col, int j) { return (size_t)col * p.NP + j; }
__host__ __device__ inline size_t idx_J(const DevParams& p, int col, int b, int j) {
    return ((size_t)col * p.Bout + b) * p.NP + j;
}

// phi as update_kzz_and_gravpot() sees it at the start of a step.  update_mass_and_pot() rewrites phi[2..num_cells] every step
// (source.cpp:60-77); index num_cells is a ghost cell at order 2, so whatever a right-hand boundary condition wrote there during
// the previous step (reflecting / periodic / user hook: phi[ghost] = phi[active]) is undone before phi_s is refreshed, while
// the outermost entry num_cells+1 and the entries 0, 1 keep the boundary condition's value.
__device__ __forceinline__ double phi_snapshot(const DevParams& p, int col, const double* phi, int j) {
    return (j == p.N && !p.use_self_gravity) ? p.phiN[col] : phi[j];     // with self-gravity phi[2..N] is recomputed in the kernel
}

// Is the hydrodynamics module on for the step that starts at time t?  (advection.cpp:62-64: the switch flips, for good, once
// globalTime has passed START_HYDRO_TIME; t is monotonic, so the comparison alone reproduces the sticky flag.)
__device__ __forceinline__ bool hydro_on(const DevParams& p, double t) {
    return p.do_hydrodynamics != 0 || (p.start_hydro_time > 0. && t > p.start_hydro_time);
}

// physical constants, verbatim values of aiolos.h:47-77
constexpr double C_G = 6.678e-8, C_PI = 3.141592653589793, C_CLIGHT = 2.99792458e10, C_AMU = 1.66054e-24;
constexpr double C_AU = 1.495978707e13;
constexpr double C_KB = 1.380649e-16, C_SIGMA_RAD = 5.670374419e-5, C_ELM_CHARGE = 4.80320425e-10;
constexpr double C_K_TO_EV = 8.621738e-5, C_EV_TO_K = 1. / C_K_TO_EV;

// std::min / std::max with the operand order semantics of libstdc++ (matters only for NaN operands)
__device__ __forceinline__ double smin(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double smax(double a, double b) { return (a < b) ? b : a; }

// Arithmetic with a fixed rounding sequence (never re-associated or contracted by the compiler).  Used where the
// reference's dense solves are ill-conditioned and its result depends on exactly which products are fused: the LU of the
// stiff friction matrix amplifies a one-ulp change by ~1e14 (tests/golden/README.md).  The forms below are the ones g++
// -O3 -ffp-contract=fast emits for the oracle build of those loops (a -= b*c  ->  fnmadd).
#ifndef AIO_EMU_BUILD
__device__ __forceinline__ double d_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double d_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double d_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double d_fnma(double a, double b, double c) { return __fma_rn(-a, b, c); }    // c - a*b, one rounding
#else
inline double d_mul(double a, double b) { volatile double r = a * b; return r; }
inline double d_add(double a, double b) { volatile double r = a + b; return r; }
inline double d_sub(double a, double b) { volatile double r = a - b; return r; }
inline double d_fnma(double a, double b, double c) { return std::fma(-a, b, c); }
#endif

// CTA-wide max reduction; every thread gets the result.  `red` must hold >= 32 doubles.
__device__ inline double block_max(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double r = (l < nw) ? red[l] : -1e308;
    for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
    return r;
}
__device__ inline int block_or(int v) { return __syncthreads_or(v); }

}  // namespace aio
