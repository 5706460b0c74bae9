// Glue between the engine and a link-time problem plugin (hooks.h).  Included at the very end of engine.cu so that the
// facade's global names (`using namespace std`, G, pi, km, ... exactly as the reference header has them) cannot leak into the
// engine sources above.  Built only when the library is made with `make PROBLEM=<file>`:
//   AIO_PROBLEM_FILE = the prepared copy of the problem file (aiolos_b200/plugin/prepare_problem.py),
//   AIO_PROBLEM_NAME = the path it was made from (reported by aio_version / hooks_name()).
#ifdef AIO_PROBLEM_FILE

#include "../plugin/aiolos.h"
#include AIO_PROBLEM_FILE

namespace aio {

bool hooks_compiled() { return true; }
const char* hooks_name() { return AIO_PROBLEM_NAME; }

namespace {
struct HostBinding {            // facade objects over one species' host arrays (AoS copy of u for the duration of a call)
    c_Sim sim; EOS_Base eos; c_Species sp; std::vector<AOS> view; int err = 0;
    ::AOS* cells = nullptr; ::AOS_prim* prims = nullptr; long n = 0;
    explicit HostBinding(HookHost& h) {
        n = h.num_cells + 2;
        cells = new ::AOS[n]; prims = new ::AOS_prim[n];
        for (long j = 0; j < n; j++) cells[j] = ::AOS(h.u1[j], h.u2[j], h.u3[j]);
        sim.num_cells = h.num_cells; sim.num_ghosts = h.num_ghosts; sim.num_species = h.num_species;
        sim.num_bands_in = h.num_bands_in; sim.num_bands_out = h.num_bands_out;
        sim.planet_mass = h.planet_mass; sim.planet_semimajor = h.planet_semimajor; sim.star_mass = h.star_mass;
        sim.T_star = h.T_star; sim.R_star = h.R_star; sim.UV_star = h.UV_star; sim.X_star = h.X_star;
        sim.domain_min = h.domain_min; sim.domain_max = h.domain_max;
        sim.phi = {h.phi, n}; sim.x_i = {h.x_i, n}; sim.x_i12 = {h.x_i12, n}; sim.x_iVC = {h.x_iVC, n};
        sim.dx = {h.dx, n}; sim.surf = {h.surf, n}; sim.vol = {h.vol, n};
        eos._gamma_m1 = h.gm1; eos._cv = h.cv; eos._mass = h.mass_amu * amu;
        sp.base = &sim; sp.eos = &eos; sp.num_cells = h.num_cells; sp.this_species_index = h.species_index;
        sp.is_dust_like = h.is_dust_like; sp.mass_amu = h.mass_amu; sp.inv_mass = h.inv_mass;
        sp.initial_fraction = h.initial_fraction; sp.density_excess = h.density_excess; sp.static_charge = h.static_charge;
        sp.cv = h.cv; sp.gamma_adiabat = h.gamma_adiabat; sp.degrees_of_freedom = h.degrees_of_freedom;
        sp.const_T_space = h.const_T_space; sp.const_opacity = h.const_opacity; sp.speciesname = h.speciesname;
        view.p = cells; view.nlo = n; view.hi0 = n; view.shift = 0; view.n = n; view.err = &err;     // the whole array
        sp.u = view;
        sp.prim.p = prims; sp.prim.nlo = n; sp.prim.hi0 = n; sp.prim.shift = 0; sp.prim.n = n; sp.prim.err = &err;
    }
    void commit(HookHost& h) {
        if (err) throw std::runtime_error("problem plugin: a hook indexed a cell outside 0..num_cells+1");
        for (long j = 0; j < n; j++) { h.u1[j] = cells[j].u1; h.u2[j] = cells[j].u2; h.u3[j] = cells[j].u3; }
    }
    ~HostBinding() { delete[] cells; delete[] prims; }
};
}  // namespace

void hook_initial_conditions(HookHost& h) {
    HostBinding b(h);
    b.sp.user_initial_conditions();
    b.commit(h);
}

void hook_boundary_host(HookHost& h, int side) {
    HostBinding b(h);
    if (side == 0) b.sp.user_boundary_left(b.view); else b.sp.user_boundary_right(b.view);
    b.commit(h);
}

// Device side: called by the one thread that fills a column's ghost cells (apply_boundary_left/right, hydro.cuh) with the
// shared boundary window: cells 0..2ng-1 (left) or N2-2ng..N2-1 (right).  The two sides are filled concurrently by different
// threads: each call copies, exposes and writes back only its own half.
__device__ __noinline__ void hook_boundary(HookArgs a) {
    const int ng = a.ng, nw = 4 * ng, half = 2 * ng, off = a.side ? half : 0;
    ::AOS cells[4]; ::AOS_prim prims[4];
    int err = 0;
    double* w = a.win;
    for (int k = 0; k < half; k++) cells[k] = ::AOS(w[0 * nw + off + k], w[1 * nw + off + k], w[2 * nw + off + k]);
    c_Sim sim; EOS_Base eos; c_Species sp;
    const long n = a.N2;
    sim.num_cells = a.N; sim.num_ghosts = ng; sim.num_species = a.S; sim.num_bands_in = a.Bin; sim.num_bands_out = a.Bout;
    sim.planet_mass = a.planet_mass; sim.planet_semimajor = a.planet_semimajor; sim.star_mass = a.star_mass;
    sim.T_star = a.T_star; sim.R_star = a.R_star; sim.UV_star = a.UV_star; sim.X_star = a.X_star;
    sim.domain_min = a.domain_min; sim.domain_max = a.domain_max;
    sim.globalTime = a.cs->t; sim.dt = a.cs->dt; sim.steps = a.cs->steps;
    sim.phi = {a.phi, n};
    sim.x_i = {const_cast<double*>(a.x_i), n}; sim.x_i12 = {const_cast<double*>(a.x_i12), n};
    sim.x_iVC = {const_cast<double*>(a.x_iVC), n}; sim.dx = {const_cast<double*>(a.dx), n};
    sim.surf = {const_cast<double*>(a.surf), n}; sim.vol = {const_cast<double*>(a.vol), n};
    eos._gamma_m1 = a.gm1; eos._cv = a.cv; eos._mass = a.mass;
    sp.base = &sim; sp.eos = &eos; sp.num_cells = a.N; sp.this_species_index = a.s; sp.is_dust_like = a.is_dust_like;
    sp.mass_amu = a.mass_amu; sp.inv_mass = a.inv_mass; sp.initial_fraction = a.initial_fraction;
    sp.density_excess = a.density_excess; sp.static_charge = a.static_charge; sp.cv = a.cv;
    sp.gamma_adiabat = a.gamma_adiabat; sp.degrees_of_freedom = a.degrees_of_freedom;
    sp.const_T_space = a.const_T_space; sp.const_opacity = a.const_opacity;
    std::vector<AOS> view;
    view.p = cells; view.n = n; view.err = &err;
    if (a.side == 0) { view.nlo = half; view.hi0 = n; view.shift = 0; }           // cells 0 .. 2ng-1
    else { view.nlo = 0; view.hi0 = n - half; view.shift = n - half; }            // cells N2-2ng .. N2-1
    sp.u = view;
    sp.prim.p = prims; sp.prim.n = n; sp.prim.err = &err; sp.prim.nlo = view.nlo; sp.prim.hi0 = view.hi0; sp.prim.shift = view.shift;
    for (int k = 0; k < half; k++) { eos.compute_primitive(&cells[k], &prims[k], 1); eos.compute_auxillary(&prims[k], 1); }
    // user_boundary_right of the shipped plugins ignores its argument and writes the member `u` (cross_over.cpp:23-29):
    // both name the same window here
    if (a.side == 0) sp.user_boundary_left(view); else sp.user_boundary_right(sp.u);
    for (int k = 0; k < half; k++) { w[0 * nw + off + k] = cells[k].u1; w[1 * nw + off + k] = cells[k].u2; w[2 * nw + off + k] = cells[k].u3; }
    if (err) { a.cs->status = COL_HOOK_RANGE; a.cs->t = 1.1 * a.cs->t_max; }
}

}  // namespace aio

#else   // stock library: no plugin

namespace aio {
bool hooks_compiled() { return false; }
const char* hooks_name() { return ""; }
void hook_initial_conditions(HookHost&) { throw std::runtime_error("unsupported configuration: PARI_PROBLEM_NUMBER 0 needs a library built with a problem plugin (make PROBLEM=<file>)"); }
void hook_boundary_host(HookHost&, int) { throw std::runtime_error("unsupported configuration: user boundary hooks need a library built with a problem plugin (make PROBLEM=<file>)"); }
}  // namespace aio

#endif
