// aiolos.h FACADE for link-time problem plugins (reference: makefile:22,28-29 `make PROBLEM=name` links problems/name.cpp,
// which defines the five c_Species::user_* members declared in the reference's aiolos.h:1050-1055).
//
// A problem file written for the reference says `#include "aiolos.h"`; built with `make -C aiolos_b200/csrc PROBLEM=...` this
// header is what it gets.  c_Species / c_Sim here are thin VIEWS over the engine's structure-of-arrays state that expose the
// members the shipped plugins use: u[j] (.u1/.u2/.u3), prim[j], eos->compute_conserved/primitive/auxillary,
// compute_pressure(u), initial_fraction, mass_amu, num_cells, base->num_ghosts, base->phi[j] (writable), base->x_i12[j],
// base->planet_mass, base->T_star, base->R_star, base->planet_semimajor, the constants, ...  Anything else fails to compile,
// which is the intended diagnostic.  user_boundary_left/right run INSIDE the hydro kernel (one thread per column side, at
// the start of every Heun stage, advection.cpp:422-423) and once on the host when the column is created
// (init_and_bounds.cpp:1275-1276); user_initial_conditions runs on the host once per species (init_and_bounds.cpp:1256-1257).
//
// On the device `u` is a window onto the 2*num_ghosts cells next to the hook's own boundary (ghosts + the same number of
// active cells); an access outside it freezes the column with status COL_HOOK_RANGE instead of reading another thread's cells.
#pragma once
#include <algorithm>
#include <cmath>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#if defined(__CUDACC__)
#define AIO_HD __host__ __device__
#else
#define AIO_HD
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#ifndef __managed__
#define __managed__
#endif
#endif

using namespace std;   // the reference header does this (aiolos.h:40); plugins rely on unqualified cout/min/max/pow

// physical constants: names and values of aiolos.h:47-77
const double G = 6.678e-8, pi = 3.141592653589793, c_light = 2.99792458e10, navo = 6.02214e23, amu = 1.66054e-24;
const double h_planck = 6.62607015e-27, Rgas = 8.31446261815324e7, kb = 1.380649e-16, km = 1e5, mearth = 5.98e27;
const double msolar = 2.0e33, au = 1.495978707e13, year = 365 * 24 * 3600, myear = 1e6 * year, gyear = 1e9 * year;
const double rearth = 6370e5, rjupiter = 69911 * km, rsolar = 695510 * km, pc = 3.08567758e18, kpc = 1e3 * pc, mpc = 1e6 * pc;
const double angstroem = 1e-4, ergcm2_to_wattperm2 = 1e-3, sigma_rad = 5.670374419e-5, K_to_eV = 8.621738e-5;
const double ev_to_K = 1. / K_to_eV, elm_charge = 4.80320425e-10;

struct AOS {                    // conserved variables of one cell (aiolos.h:194-238)
    double u1, u2, u3;
    AIO_HD AOS(double a = 0, double b = 0, double c = 0) : u1(a), u2(b), u3(c) {}
    AIO_HD AOS& operator+=(const AOS& a) { u1 += a.u1; u2 += a.u2; u3 += a.u3; return *this; }
    AIO_HD AOS& operator-=(const AOS& a) { u1 -= a.u1; u2 -= a.u2; u3 -= a.u3; return *this; }
    AIO_HD AOS operator+(const AOS& a) const { return AOS(u1 + a.u1, u2 + a.u2, u3 + a.u3); }
    AIO_HD AOS operator-(const AOS& a) const { return AOS(u1 - a.u1, u2 - a.u2, u3 - a.u3); }
    AIO_HD AOS operator*(const double& a) const { return AOS(u1 * a, u2 * a, u3 * a); }
    AIO_HD AOS operator/(const double& a) const { return AOS(u1 / a, u2 / a, u3 / a); }
};

struct AOS_prim {               // primitive + auxiliary variables of one cell (aiolos.h:240-247)
    AIO_HD AOS_prim(double d = 0, double s = 0, double p = 0, double cs = 0, double u = 0, double T = 0, double n = 0)
        : density(d), speed(s), pres(p), sound_speed(cs), internal_energy(u), temperature(T), number_density(n) {}
    double density, speed, pres;
    double sound_speed, internal_energy, temperature, number_density;
};

namespace aio_plugin {
// Index-checked view used for u / prim: the full array on the host, the two boundary windows on the device.
// Cell j lives at p[j] if j < nlo, at p[j - shift] if j >= hi0; anything between is out of range.
template <typename T> struct CellView {
    T* p = nullptr; long nlo = 0, hi0 = 0, shift = 0, n = 0; int* err = nullptr; T dummy;
    AIO_HD T& at(long j) {
        if (j >= 0 && j < nlo) return p[j];
        if (j >= hi0 && j < n) return p[j - shift];
        if (err) *err = 1;
        dummy = T();
        return dummy;
    }
};
template <typename T> struct Span {     // base->phi, base->x_i12, ...: plain arrays of num_cells+2 entries
    T* p = nullptr; long n = 0;
    AIO_HD T& operator[](long j) const { return p[j]; }
    AIO_HD long size() const { return n; }
};
}  // namespace aio_plugin

// The hooks take `std::vector<AOS>&` (aiolos.h:1050-1051); for these two cell types std::vector IS the view.
namespace std {
template <> class vector<AOS, allocator<AOS>> : public aio_plugin::CellView<AOS> {
public:
    AIO_HD AOS& operator[](long j) { return at(j); }
    AIO_HD long size() const { return n; }
};
template <> class vector<AOS_prim, allocator<AOS_prim>> : public aio_plugin::CellView<AOS_prim> {
public:
    AIO_HD AOS_prim& operator[](long j) { return at(j); }
    AIO_HD long size() const { return n; }
};
}  // namespace std

// Ideal-gas EOS of one species (eos.h:28-47, eos.cpp:19-90); non-virtual here: hooks only call through the pointer.
class EOS_Base {
public:
    double _gamma_m1 = 0, _cv = 0, _mass = 0;    // 2/dof, c_v, mass_amu*amu
    AIO_HD void compute_primitive(const AOS* cons, AOS_prim* prim, int num_cells) const {
        for (int i = 0; i < num_cells; i++) {
            prim[i].density = cons[i].u1;
            prim[i].speed = cons[i].u2 / cons[i].u1;
            prim[i].pres = _gamma_m1 * (cons[i].u3 - 0.5 * cons[i].u2 * prim[i].speed);
        }
    }
    AIO_HD void compute_conserved(const AOS_prim* prim, AOS* cons, int num_cells) const {
        for (int i = 0; i < num_cells; i++) {
            cons[i].u1 = prim[i].density;
            cons[i].u2 = prim[i].density * prim[i].speed;
            cons[i].u3 = prim[i].pres / _gamma_m1 + 0.5 * cons[i].u2 * prim[i].speed;
        }
    }
    AIO_HD void compute_auxillary(AOS_prim* prim, int num_cells) const {
        for (int i = 0; i < num_cells; i++) {
            prim[i].number_density = prim[i].density / _mass;
            prim[i].internal_energy = prim[i].pres / prim[i].density / _gamma_m1;
            prim[i].sound_speed = sqrt((_gamma_m1 + 1) * _gamma_m1 * prim[i].internal_energy);
            prim[i].temperature = prim[i].internal_energy / _cv;
        }
    }
    AIO_HD void update_p_from_eint(AOS_prim* prim, int num_cells) const {
        for (int i = 0; i < num_cells; i++) prim[i].pres = prim[i].internal_energy * prim[i].density * _gamma_m1;
    }
    AIO_HD void update_eint_from_T(AOS_prim* prim, int num_cells) const {
        for (int i = 0; i < num_cells; i++) prim[i].internal_energy = prim[i].temperature * _cv;
    }
    AIO_HD double get_gamma_adiabatic() const { return _gamma_m1 + 1; }
    AIO_HD double get_cv() const { return _cv; }
};

class c_Sim {                   // what hooks reach through `base->` (aiolos.h:361-828)
public:
    int num_cells = 0, num_ghosts = 0, num_species = 0, num_bands_in = 0, num_bands_out = 0;
    double planet_mass = 0, planet_semimajor = 0, star_mass = 0, T_star = 0, R_star = 0, UV_star = 0, X_star = 0;
    double domain_min = 0, domain_max = 0, globalTime = 0, dt = 0;
    long long steps = 0;
    aio_plugin::Span<double> phi, x_i, x_i12, x_iVC, dx, surf, vol;
};

class c_Species {               // aiolos.h:841-1076, members used by problem files
public:
    c_Sim* base = nullptr;
    EOS_Base* eos = nullptr;
    int num_cells = 0, this_species_index = 0, is_dust_like = 0;
    double mass_amu = 0, inv_mass = 0, initial_fraction = 0, density_excess = 0, static_charge = 0;
    double cv = 0, gamma_adiabat = 0, degrees_of_freedom = 0, const_T_space = 0, const_opacity = 0;
    const char* speciesname = "";
    std::vector<AOS> u;
    std::vector<AOS_prim> prim;

    AIO_HD void compute_pressure(std::vector<AOS>& uu) {          // prim + auxiliaries from `uu` wherever both views are defined
        for (long j = 0; j < uu.size(); j++) {
            if (j >= uu.nlo && j < uu.hi0) { j = uu.hi0 - 1; continue; }
            eos->compute_primitive(&uu[j], &prim[j], 1);
            eos->compute_auxillary(&prim[j], 1);
        }
    }

    // the link-time hooks (aiolos.h:1050-1055); defined by the problem file
    AIO_HD void user_boundary_left(std::vector<AOS>& u);
    AIO_HD void user_boundary_right(std::vector<AOS>& u);
    void user_initial_conditions();
    void user_opacity();
    void user_species_loop_function();
};
