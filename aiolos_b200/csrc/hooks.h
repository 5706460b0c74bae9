// Link-time problem plugin (reference makefile:22,28-29: `make PROBLEM=name` links problems/name.cpp, which defines the five
// c_Species::user_* hooks of aiolos.h:1050-1055).  This is the plain-data interface between the host front end
// (frontend.cpp, compiled by g++) and the glue that owns the facade types (hooks_glue.inl, compiled into engine.cu's
// translation unit together with the problem file so that the boundary hooks can run inside the hydro kernels).
#pragma once

namespace aio {

struct HookHost {                 // one species of one column, on the host; arrays have num_cells+2 entries
    int species_index, num_cells, num_ghosts, num_species, num_bands_in, num_bands_out, is_dust_like;
    double mass_amu, inv_mass, initial_fraction, density_excess, static_charge, cv, gamma_adiabat, degrees_of_freedom;
    double const_T_space, const_opacity, gm1;
    const char* speciesname;
    double planet_mass, planet_semimajor, star_mass, T_star, R_star, UV_star, X_star, domain_min, domain_max;
    double *phi, *x_i, *x_i12, *x_iVC, *dx, *surf, *vol;
    double *u1, *u2, *u3;
};

bool hooks_compiled();            // false in the stock library: decks that need a hook are rejected at create time
const char* hooks_name();         // path of the problem file this library was built with ("" if none)
void hook_initial_conditions(HookHost& h);       // c_Species::user_initial_conditions(), init_and_bounds.cpp:1256-1257
void hook_boundary_host(HookHost& h, int side);  // side 0: user_boundary_left(u), 1: user_boundary_right(u); ctor call :1275-1276

}  // namespace aio
