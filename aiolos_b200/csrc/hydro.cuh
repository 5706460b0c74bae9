// Fused hydrodynamics kernel: one CTA advances one column through the hydro part of one time step.
//
//   step head   dt selection                       reference advection.cpp:73-77 + helpers.cpp:23-87
//   stage 1/2   boundary conditions                init_and_bounds.cpp:1779-1987
//               cons -> prim, auxiliaries          eos.cpp:19-63
//               well-balanced reconstruction       reconstruct.cpp:11-117, aiolos.h:1003-1016
//               HLLC / dust Riemann flux           advection.cpp:544-650
//               gravity + geometric sources, dudt  source.cpp:157-164, advection.cpp:492-503
//   Heun combine                                   advection.cpp:154-158,187-195
//   step tail (if no friction/radiation follows)   helpers.cpp:28-71 (next dt inputs), advection.cpp:282-307
//
// Thread t owns cells t, t+NT, t+2NT, ... (CPT of them): u and dudt0 of those cells stay in registers for the
// whole step, so HBM sees one read and one write of u per species per step.  Stencil neighbours are exchanged
// through shared memory (9 arrays of N+2 doubles: P/F, L, R), which also makes every edge state and every
// interface flux computed exactly once.
#pragma once
#include "common.cuh"

namespace aio {

struct Prim3 { double rho, v, p; };

__device__ __forceinline__ Prim3 cons_to_prim(double u1, double u2, double u3, double gm1) {
    Prim3 q;                                    // IdealGas_EOS::compute_primitive, eos.cpp:19-29
    q.rho = u1;
    q.v = u2 / u1;
    q.p = gm1 * (u3 - 0.5 * u2 * q.v);
    return q;
}
__device__ __forceinline__ void prim_to_cons(const Prim3& q, double gm1, double& u1, double& u2, double& u3) {
    u1 = q.rho;                                 // IdealGas_EOS::compute_conserved, eos.cpp:38-47
    u2 = q.rho * q.v;
    u3 = q.p / gm1 + 0.5 * u2 * q.v;
}
__device__ __forceinline__ double eint_of(double p, double rho, double gm1) { return (p / rho) / gm1; }   // eos.cpp:58
__device__ __forceinline__ double cs_of(double eint, double gm1) { return sqrt((gm1 + 1) * gm1 * eint); } // eos.cpp:59

// reconstruct.cpp:11-29
__device__ __forceinline__ double mc_slope(double ql, double qm, double qr, double cF, double cB, double dxF, double dxB) {
    double dF = (qr - qm) / dxF;
    double dB = (qm - ql) / dxB;
    if (dF * dB < 0) return 0;
    double a = cF * dF, b = cB * dB;
    double m = (fabs(a) < fabs(b)) ? a : b;
    double c = 0.5 * (dF + dB);
    return (fabs(c) < fabs(m)) ? c : m;
}

// aiolos.h:1003-1016 (phi_s passed as the three neighbouring values)
__device__ __forceinline__ double dp_hydro_minus(const DevParams& p, int i, double phis_im, double phis_i) {
    return -__ldg(p.dx + i) * __ldg(p.omegaplus + i) * (phis_im - phis_i) / (__ldg(p.dx + i - 1) + __ldg(p.dx + i));
}
__device__ __forceinline__ double dp_hydro_plus(const DevParams& p, int i, double phis_ip, double phis_i) {
    return -__ldg(p.dx + i) * __ldg(p.omegaminus + i) * (phis_ip - phis_i) / (__ldg(p.dx + i + 1) + __ldg(p.dx + i));
}

// advection.cpp:544-615.  Edge auxiliaries (reconstruct.cpp:114-115) are evaluated here, once per edge state.
__device__ __forceinline__ void hllc_flux(double dl, double ul, double pl, double dr, double ur, double pr, double gm1,
                                          double& f1, double& f2, double& f3) {
    const double el = eint_of(pl, dl, gm1), er = eint_of(pr, dr, gm1);
    const double csl = cs_of(el, gm1), csr = cs_of(er, gm1);
    const double mom_l = dl * ul, mom_r = dr * ur;
    const double El = dl * el + 0.5 * mom_l * ul;
    const double Er = dr * er + 0.5 * mom_r * ur;
    const double SL = ul - csl, SR = ur + csr;
    const double SS = (pr - pl + mom_l * (SL - ul) - mom_r * (SR - ur)) / (dl * (SL - ul) - dr * (SR - ur));
    f1 = 0; f2 = 0; f3 = 0;
    if ((SL <= 0) && (SS >= 0)) {
        const double F1 = mom_l, F2 = mom_l * ul + pl, F3 = ul * (El + pl);
        const double comp3 = El / dl + (SS - ul) * (SS + pl / (dl * (SL - ul)));
        const double a = SL - ul, b = SL - SS;
        const double U1 = ((1.0 * dl) * a) / b, U2 = ((SS * dl) * a) / b, U3 = ((comp3 * dl) * a) / b;
        f1 = F1 + (U1 - dl) * SL; f2 = F2 + (U2 - mom_l) * SL; f3 = F3 + (U3 - El) * SL;
    } else if ((SS <= 0) && (SR >= 0)) {
        const double F1 = mom_r, F2 = mom_r * ur + pr, F3 = ur * (Er + pr);
        const double comp3 = Er / dr + (SS - ur) * (SS + pr / (dr * (SR - ur)));
        const double a = SR - ur, b = SR - SS;
        const double U1 = ((1.0 * dr) * a) / b, U2 = ((SS * dr) * a) / b, U3 = ((comp3 * dr) * a) / b;
        f1 = F1 + (U1 - dr) * SR; f2 = F2 + (U2 - mom_r) * SR; f3 = F3 + (U3 - Er) * SR;
    } else if (SL >= 0) {
        f1 = mom_l; f2 = mom_l * ul + pl; f3 = ul * (El + pl);
    } else if (SR <= 0) {
        f1 = mom_r; f2 = mom_r * ur + pr; f3 = ur * (Er + pr);
    }
}

// advection.cpp:623-650
__device__ __forceinline__ void dust_flux(double dl, double ul, double pl, double dr, double ur, double pr, double gm1,
                                          double& f1, double& f2, double& f3) {
    if (ul < 0 && ur > 0) { f1 = 0; f2 = 0; f3 = 0; return; }
    const double el = eint_of(pl, dl, gm1), er = eint_of(pr, dr, gm1);
    const double R = sqrt(dr / dl);
    const double SS = (ul + R * ur) / (1 + R);
    const double ml = dl * ul, enl = dl * (el + 0.5 * ul * ul);
    const double mr = dr * ur, enr = dr * (er + 0.5 * ur * ur);
    const double a1 = ml, a2 = ul * ml + pl, a3 = ul * (enl + pl);
    const double b1 = mr, b2 = ur * mr + pr, b3 = ur * (enr + pr);
    if (SS > 0) { f1 = a1; f2 = a2; f3 = a3; }
    else if (SS < 0) { f1 = b1; f2 = b2; f3 = b3; }
    else { f1 = (a1 + b1) * 0.5; f2 = (a2 + b2) * 0.5; f3 = (a3 + b3) * 0.5; }
}

// Ghost-cell fill, run by one thread per side on the small shared window `w` that holds the 2*ng cells next to each
// boundary (w[c][k]: k<2ng -> cell k ; k>=2ng -> cell N2-4ng+k).  init_and_bounds.cpp:1779-1974.
struct BndWin {
    double* w; int ng, N2;
    __device__ double& at(int c, int cell) { int k = (cell < 2 * ng) ? cell : cell - (N2 - 4 * ng); return w[c * 4 * ng + k]; }
};

#ifdef AIO_PROBLEM_FILE
// What a boundary hook can see, passed BY VALUE to the out-of-line glue (hooks_glue.inl, end of engine.cu): taking the address
// of the kernel parameter block would force a local copy of all of DevParams in every hydro kernel.
struct HookArgs {
    int N, N2, ng, S, Bin, Bout, col, s, side, is_dust_like;
    double planet_mass, planet_semimajor, star_mass, T_star, R_star, UV_star, X_star, domain_min, domain_max;
    double gm1, cv, mass, mass_amu, inv_mass, initial_fraction, density_excess, static_charge, gamma_adiabat, degrees_of_freedom;
    double const_T_space, const_opacity;
    double* phi; const double *x_i, *x_i12, *x_iVC, *dx, *surf, *vol;
    ColState* cs;
    double* win;                 // this side's half of the boundary window: [3][4*ng], see BndWin
};
__device__ void hook_boundary(HookArgs a);
__device__ __forceinline__ void call_hook(const DevParams& p, const SpeciesConst& sc, int col, int s, BndWin& W, int side) {
    HookArgs a;
    a.N = p.N; a.N2 = p.N2; a.ng = p.ng; a.S = p.S; a.Bin = p.Bin; a.Bout = p.Bout; a.col = col; a.s = s; a.side = side;
    a.is_dust_like = sc.is_dust_like;
    a.planet_mass = p.planet_mass; a.planet_semimajor = p.planet_semimajor; a.star_mass = p.star_mass; a.T_star = p.T_star;
    a.R_star = p.R_star; a.UV_star = p.UV_star; a.X_star = p.X_star; a.domain_min = p.domain_min; a.domain_max = p.domain_max;
    a.gm1 = sc.gm1; a.cv = sc.cv; a.mass = sc.mass; a.mass_amu = sc.mass_amu; a.inv_mass = sc.inv_mass;
    a.initial_fraction = sc.initial_fraction; a.density_excess = sc.density_excess; a.static_charge = sc.static_charge;
    a.gamma_adiabat = sc.gamma_adiabat; a.degrees_of_freedom = sc.degrees_of_freedom;
    a.const_T_space = sc.const_T_space; a.const_opacity = sc.const_opacity;
    a.phi = p.phi + idx_c(p, col, 0); a.x_i = p.x_i; a.x_i12 = p.x_i12; a.x_iVC = p.x_iVC; a.dx = p.dx; a.surf = p.surf; a.vol = p.vol;
    a.cs = p.cs + col; a.win = W.w;
    hook_boundary(a);
}
#endif

// KZZ: the caller may run with KZZ_INIT != 0 (v1 kernels); the second-generation kernels never do and compile the K_zzf lookup out
template <bool KZZ = true>
__device__ inline void apply_boundary_left(const DevParams& p, const SpeciesConst& sc, int col, int s, BndWin& W, double prim2_speed) {
    const int ng = p.ng;
    double* phi = p.phi + idx_c(p, col, 0);
    switch (sc.bc_left) {
    case 1:  // open
        for (int i = ng; i > 0; i--) {
            Prim3 q = cons_to_prim(W.at(0, i), W.at(1, i), W.at(2, i), sc.gm1);
            const double* kz = (KZZ && p.kzzf) ? p.kzzf + idx_s(p, col, s, 0) : nullptr;      // K_zzf (1 without KZZ_INIT)
            double dphi = (phi[i] * (kz ? kz[i] : 1.0) - phi[i - 1] * (kz ? kz[i - 1] : 1.0)) / (p.dx[i - 1] + p.dx[i]);
            dphi *= (p.omegaplus[i] * p.dx[i] + p.omegaminus[i - 1] * p.dx[i - 1]);
            q.p = q.p + q.rho * dphi;
            q.p = smax(q.p, 0.0);
            q.v = W.at(1, i) / W.at(0, i);
            prim_to_cons(q, sc.gm1, W.at(0, i - 1), W.at(1, i - 1), W.at(2, i - 1));
        }
        break;
    case 2: case 5:  // reflecting, giantplanet (identical on the left)
        for (int i = 0; i < ng; i++) {
            int igh = ng - 1 - i, iact = ng + i;
            W.at(0, igh) = W.at(0, iact); W.at(1, igh) = W.at(1, iact) * -1; W.at(2, igh) = W.at(2, iact);
            phi[igh] = phi[iact];
        }
        break;
    case 3:  // fixed
        for (int i = 0; i < ng; i++) {
            Prim3 q; q.rho = 0; q.v = 0; double T = 0;
            if (p.problem_number != 1) {          // problem 1 leaves the state zero-initialised (reference bug)
                q.rho = sc.background_u[0];
                q.v = (prim2_speed > 0) ? prim2_speed : 0.;
                T = sc.const_T_space;
            }
            double e = T * sc.cv;
            q.p = e * q.rho * sc.gm1;
            prim_to_cons(q, sc.gm1, W.at(0, i), W.at(1, i), W.at(2, i));
        }
        break;
    case 4: {  // periodic
        const int Ncell = p.N - 2 * (ng - 1);
        for (int i = 0; i < ng; i++) {
            int iact = Ncell + i;
            W.at(0, i) = W.at(0, iact); W.at(1, i) = W.at(1, iact); W.at(2, i) = W.at(2, iact);
            phi[i] = phi[iact];
        }
        break; }
    default:         // 0 = user hook (problem plugin); the stock library rejects such decks at create time
#ifdef AIO_PROBLEM_FILE
        if (sc.bc_left == 0) call_hook(p, sc, col, s, W, 0);
#endif
        break;
    }
}

template <bool KZZ = true>
__device__ inline void apply_boundary_right(const DevParams& p, const SpeciesConst& sc, int col, int s, BndWin& W) {
    const int ng = p.ng;
    const int Ncell = p.N - 2 * (ng - 1);
    double* phi = p.phi + idx_c(p, col, 0);
    switch (sc.bc_right) {
    case 1:  // open
        for (int i = Ncell + ng; i < Ncell + 2 * ng; i++) {
            Prim3 q = cons_to_prim(W.at(0, i - 1), W.at(1, i - 1), W.at(2, i - 1), sc.gm1);
            const double* kz = (KZZ && p.kzzf) ? p.kzzf + idx_s(p, col, s, 0) : nullptr;      // K_zzf (1 without KZZ_INIT)
            double dphi = (phi[i] * (kz ? kz[i] : 1.0) - phi[i - 1] * (kz ? kz[i - 1] : 1.0)) / (p.dx[i - 1] + p.dx[i]);
            dphi *= (p.omegaplus[i] * p.dx[i] + p.omegaminus[i - 1] * p.dx[i - 1]);
            q.p = q.p - p.right_extrap_press_multiplier * q.rho * dphi;
            q.p = smax(q.p, 1e-1 * q.p);
            prim_to_cons(q, sc.gm1, W.at(0, i), W.at(1, i), W.at(2, i));
        }
        break;
    case 2:  // reflecting
        for (int i = 0; i < ng; i++) {
            int iact = Ncell + ng - 1 - i, igh = Ncell + ng + i;
            W.at(0, igh) = W.at(0, iact); W.at(1, igh) = W.at(1, iact) * -1; W.at(2, igh) = W.at(2, iact);
            phi[igh] = phi[iact];
        }
        break;
    case 3:  // fixed
        for (int i = Ncell + ng; i < Ncell + 2 * ng; i++) {
            W.at(0, i) = sc.shock_ur[0]; W.at(1, i) = sc.shock_ur[1]; W.at(2, i) = sc.shock_ur[2];
        }
        break;
    case 4:  // periodic
        for (int i = Ncell + ng; i < Ncell + 2 * ng; i++) {
            int iact = i - Ncell;
            W.at(0, i) = W.at(0, iact); W.at(1, i) = W.at(1, iact); W.at(2, i) = W.at(2, iact);
            phi[i] = phi[iact];
        }
        break;
    case 5:  // giantplanet
        for (int i = Ncell + ng; i < Ncell + 2 * ng; i++) {
            Prim3 q = cons_to_prim(W.at(0, i - 1), W.at(1, i - 1), W.at(2, i - 1), sc.gm1);
            double r = p.x_i12[i];
            double mdot = 4. * 3.141592 * q.rho * q.v * r * r;
            double freefallv = -sqrt(2. * C_G * p.planet_mass / r);
            double mdotlimit = p.max_mdot;
            if (sc.is_dust_like) mdotlimit = 1e-4 * mdotlimit;
            if (mdot < mdotlimit) { q.v = freefallv; q.rho = mdotlimit / freefallv / r / r; }
            else { q.p = smin(q.p * 2., 1.); }
            prim_to_cons(q, sc.gm1, W.at(0, i), W.at(1, i), W.at(2, i));
        }
        break;
    default:
#ifdef AIO_PROBLEM_FILE
        if (sc.bc_right == 0) call_hook(p, sc, col, s, W, 1);
#endif
        break;
    }
}

// Step head: the dt chain of advection.cpp:73-77 with get_cfl_timestep() (helpers.cpp:47,74-85) folded in.
// minstep / maxde come from the previous step's tail (or from k_prime for the first step).
__device__ inline double step_head_dt(const DevParams& p, ColState& c) {
    c.timestep_rad2 = c.dt / c.maxde * p.energy_epsilon;
    if (c.t > p.CFL_break_time) c.cflfactor = 0.9;
    c.cfl_step = c.cflfactor / c.minstep;
    double dt;
    if (hydro_on(p, c.t)) dt = smin(c.cfl_step, c.dt * p.max_timestep_change);
    else dt = smin(smin(c.timestep_rad2, c.dt * p.max_timestep_change), p.dt_max);     // helpers.cpp:80-85
    dt = smin(dt, c.timestep_rad2);
    dt = smin(dt, c.t_max - c.t);
    if (c.steps == 0) dt = smin(dt, p.dt_min_init);
    c.dt = dt;
    return dt;
}

// Per-cell contribution to the quantities the next step's dt needs + crash scan + primlast update.
// Loop bounds follow helpers.cpp:31,59 and advection.cpp:89,283.
struct TailAcc { double minstep, maxde; int bad; double badval; };
__device__ __forceinline__ void tail_cell(const DevParams& p, int col, int s, int j, double v, double cs, double eint, double T,
                                          TailAcc& a) {
    const size_t k = idx_s(p, col, s, j);
    if (j >= 1 && j <= p.N - 1) {
        double de = fabs(p.eint[k] - eint) / eint;
        a.maxde = fmax(a.maxde, de);
    }
    if (j >= 1 && j <= p.N) {
        double dxj = __ldg(p.dx + j);
        a.minstep = fmax(a.minstep, fabs(v / dxj) + cs / dxj);
    }
    if (j <= p.N - 1 && (T < 0 || isnan(T))) { a.bad = 1; a.badval = T; }
    if (j <= p.N) p.eint[k] = eint;
}

__device__ inline void tail_commit_on(ColState& c, TailAcc a, double* red, bool negJ = false, double Jval = 0) {
    double ms = block_max(a.minstep, red);
    double md = block_max(a.maxde, red);
    int bad = block_or(a.bad);
    if (threadIdx.x == 0) {
        c.minstep = ms; c.maxde = md;
        c.steps += 1;
        if (bad) { c.status = COL_NEG_T; c.t = 1.1 * c.t_max; }
        else if (negJ) { c.status = COL_NEG_J; c.crash_value = Jval; c.t = 1.1 * c.t_max; }
    }
}
__device__ inline void tail_commit(const DevParams& p, int col, TailAcc a, double* red, bool negJ = false, double Jval = 0) {
    tail_commit_on(p.cs[col], a, red, negJ, Jval);
}

// update_mass_and_pot() with PARI_SELF_GRAV_SWITCH 1 (source.cpp:50-109): enclosed mass by the reference's sequential sum (one
// lane, so that the rounding sequence is the reference's), then phi[2..N] = get_phi_grav(x_i12, M_enc) (+ tides).  RHO holds the
// summed species density per cell; W and EM are work arrays of N+2 doubles.
__device__ inline void self_gravity_update(const DevParams& p, const int col, double* W, const double* RHO, double* EM) {
    const int tid = threadIdx.x, NT = blockDim.x, N = p.N;
    for (int j = tid + 2; j <= N; j += NT) {
        double xn3 = p.x_i[j]; xn3 = xn3 * xn3 * xn3;
        double xl3 = p.x_i[j - 1]; xl3 = xl3 * xl3 * xl3;
        W[j] = 4. * 3.141592 * (xn3 - xl3) / 3.;
    }
    __syncthreads();
    if (tid == 0) {
        double em = p.planet_mass;
        EM[0] = em; EM[1] = em;
        for (int j = 2; j <= N; j++) { em = em + W[j] * RHO[j]; EM[j] = em; }
    }
    __syncthreads();
    double* phi = p.phi + idx_c(p, col, 0);
    for (int j = tid + 2; j <= N; j += NT) {
        const double r = p.x_i12[j], mass = EM[j];
        double ph;
        if (p.use_linear_gravity) ph = -C_G * mass * fabs(p.domain_max - r);
        else if (fabs(r - p.planet_position) > p.rs) ph = -C_G * mass / fabs(r - p.planet_position);
        else ph = -C_G * mass * (pow(r - p.planet_position, 3.) / pow(p.rs, 4.) - 2. * pow(r - p.planet_position, 2.) / pow(p.rs, 3.) + 2. / p.rs);
        if (p.use_tides == 1) {
            double d3 = p.planet_semimajor * C_AU;
            d3 = d3 * d3 * d3;
            ph -= 0.5 * 3. * C_G * p.star_mass * r * r / d3;
        }
        phi[j] = ph;
    }
    __syncthreads();
}

// One column, one time step of hydrodynamics for all species.  `sm` = dynamic shared memory of the CTA
// (9*N2 + 32 + 12*ng doubles).  `snap` != 0 refreshes the phi_s snapshot first.
template <int CPT>
__device__ __forceinline__ void hydro_phase(const DevParams& p, const int col, double* sm, const double dt, const double t_old,
                                            const int snap, TailAcc& acc) {
    const int tid = threadIdx.x, NT = blockDim.x;
    const int N = p.N, N2 = p.N2, ng = p.ng;
    double* Pr = sm;            double* Pv = Pr + N2;  double* Pp = Pv + N2;   // cell prims; reused for the fluxes
    double* Lr = Pp + N2;       double* Lv = Lr + N2;  double* Lp = Lv + N2;   // left edge states  prim_l
    double* Rr = Lp + N2;       double* Rv = Rr + N2;  double* Rp = Rv + N2;   // right edge states prim_r
    double* red = Rp + N2;                                                     // 32 doubles
    double* win = red + 32;                                                    // 3*4*ng doubles
    BndWin W; W.w = win; W.ng = ng; W.N2 = N2;

    // phi_s snapshot for all species before any BC edits phi (update_kzz_and_gravpot, advection.cpp:101-102 /
    // source.cpp:207-209 with K_zzf = 1).  phi is static after the first step (engine.cu), so steps >= 2 skip it.
    if (p.use_self_gravity) {           // advection.cpp:99 (single species: checked at create time)
        for (int j = tid; j < N2; j += NT) Pv[j] = p.u[idx_u(p, 0, col, 0, j)];
        __syncthreads();
        self_gravity_update(p, col, Pr, Pv, Pp);
    }
    if (snap || p.use_self_gravity || p.kzzf) {
        const double* phi = p.phi + idx_c(p, col, 0);
        for (int s = 0; s < p.S; s++) {
            double* phis = p.phis + idx_s(p, col, s, 0);
            if (!p.kzzf) {
                for (int j = tid; j < N2; j += NT) phis[j] = phi_snapshot(p, col, phi, j) * 1.0;
            } else {
                // update_kzz_and_gravpot (source.cpp:174-219): below the homopause of the dominant species (species 0) a heavier
                // species feels the potential scaled by mu / m_i, shifted so that phi_s is continuous at the homopause
                double* kz = p.kzzf + idx_s(p, col, s, 0);
                const double mu = p.sp[0].mass_amu, mi = p.sp[s].mass_amu;
                const bool scaled = (s > 0 && mi > 1e-3);
                const double one = 0.99999;
                auto kzzf_at = [&](int i) -> double {
                    if (!scaled) return 1.;
                    const double n = p.u[idx_u(p, 0, col, 0, i)] / p.sp[0].mass;      // species[0].prim[i].number_density
                    return (p.kzz * n < 1.) ? 1. : mu / mi;
                };
                double hmax = 0;
                for (int j = tid; j < N2; j += NT) {
                    const double k = kzzf_at(j);
                    kz[j] = k;
                    if (scaled && j >= 1 && k > one && kzzf_at(j - 1) < one) hmax = (double)j;     // the last such cell wins
                }
                hmax = block_max(hmax, red);
                __syncthreads();
                const int h = (int)hmax;
                const double corr = (h == 0) ? 0. : phi_snapshot(p, col, phi, h) * (1. - mu / mi);
                for (int j = tid; j < N2; j += NT) {
                    double v = phi_snapshot(p, col, phi, j) * kz[j];
                    if (j < h) v += corr;
                    phis[j] = v;
                }
            }
        }
        __syncthreads();
    }
    for (int s = 0; s < p.S; s++) {
        const SpeciesConst sc = p.sp[s];
        const double gm1 = sc.gm1;
        const bool is_gas = !sc.is_dust_like;
        const double* phis = p.phis + idx_s(p, col, s, 0);
        double u1[CPT], u2[CPT], u3[CPT], d1[CPT], d2[CPT], d3[CPT];
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            int j = tid + k * NT;
            if (j < N2) {
                u1[k] = p.u[idx_u(p, 0, col, s, j)]; u2[k] = p.u[idx_u(p, 1, col, s, j)]; u3[k] = p.u[idx_u(p, 2, col, s, j)];
            } else { u1[k] = 1; u2[k] = 0; u3[k] = 1; }
            d1[k] = d2[k] = d3[k] = 0;
        }
        double prim2_speed = p.prim2v[(size_t)col * p.S + s];
        // hydro off (DO_HYDRO 0, or START_HYDRO_TIME not reached): no stages; at second order the boundary conditions are still
        // applied once (advection.cpp:196-200), at first order nothing happens before the final compute_pressure
        const bool hyd = hydro_on(p, t_old);
        const int nstage = hyd ? ((p.order == 2) ? 2 : 1) : ((p.order == 2) ? 1 : 0);
        for (int stage = 0; stage < nstage; stage++) {
            // ---- boundary conditions -------------------------------------------------------------------
#pragma unroll
            for (int k = 0; k < CPT; k++) {
                int j = tid + k * NT;
                if (j < N2 && (j < 2 * ng || j >= N2 - 2 * ng)) { W.at(0, j) = u1[k]; W.at(1, j) = u2[k]; W.at(2, j) = u3[k]; }
            }
            __syncthreads();
            if (tid == 0) {
                apply_boundary_left(p, sc, col, s, W, prim2_speed);
                if (sc.use_wave == 1 && hyd) {                                  // init_and_bounds.cpp:1982-1987
                    double tt = (stage == 0) ? t_old : t_old + dt;
                    W.at(1, 0) = W.at(0, 0) * sc.wave_amplitude * sin(12. * M_PI * tt / sc.wave_period);
                    W.at(1, 1) = W.at(1, 0);
                }
            }
            if (tid == ((NT > 32) ? 32 : 0)) apply_boundary_right(p, sc, col, s, W);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < CPT; k++) {
                int j = tid + k * NT;
                if (j < N2 && (j < 2 * ng || j >= N2 - 2 * ng)) { u1[k] = W.at(0, j); u2[k] = W.at(1, j); u3[k] = W.at(2, j); }
            }
            if (!hyd) break;
            // ---- cons -> prim ----------------------------------------------------------------------------
#pragma unroll
            for (int k = 0; k < CPT; k++) {
                int j = tid + k * NT;
                if (j < N2) { Prim3 q = cons_to_prim(u1[k], u2[k], u3[k], gm1); Pr[j] = q.rho; Pv[j] = q.v; Pp[j] = q.p; }
            }
            __syncthreads();
            prim2_speed = Pv[2 < N2 ? 2 : 0];
            // ---- reconstruction ---------------------------------------------------------------------------
#pragma unroll
            for (int k = 0; k < CPT; k++) {
                int i = tid + k * NT;
                if (i < N2) {
                    const double rho = Pr[i], v = Pv[i], pr0 = Pp[i];
                    double lr = rho, lv = v, lp = pr0, rr = rho, rv = v, rp = pr0;
                    double dpm = 0, dpp = 0;
                    const double ph = phis[i];
                    double phm = 0, php = 0;
                    if (i > 0) phm = phis[i - 1];
                    if (i <= N) php = phis[i + 1];
                    if (is_gas) {
                        if (i > 0) { dpm = dp_hydro_minus(p, i, phm, ph); lp = pr0 + rho * dpm; }
                        if (i <= N) { dpp = dp_hydro_plus(p, i, php, ph); rp = pr0 + rho * dpp; }
                        if (lp < 0 || rp < 0) { lp = pr0; rp = pr0; }
                    }
                    if (p.order == 2 && i >= 1 && i <= N) {
                        double dp_l = 0, dp_r = 0;
                        const double rhom = Pr[i - 1], rhop = Pr[i + 1];
                        if (is_gas) {
                            dp_l = rho * dpm - rhom * dp_hydro_plus(p, i - 1, ph, phm);
                            dp_r = rho * dpp - rhop * dp_hydro_minus(p, i + 1, ph, php);
                        }
                        const double xvc = __ldg(p.x_iVC + i), xvm = __ldg(p.x_iVC + i - 1), xvp = __ldg(p.x_iVC + i + 1);
                        const double xi = __ldg(p.x_i + i), xim = __ldg(p.x_i + i - 1);
                        const double cF = (xvp - xvc) / (xi - xvc), cB = (xvc - xvm) / (xvc - xim);
                        const double dxF = (xvp - xvc), dxB = (xvc - xvm);
                        const double wl = (xim - xvc), wr = (xi - xvc);
                        double slope = mc_slope(Pp[i - 1] - dp_l, pr0, Pp[i + 1] - dp_r, cF, cB, dxF, dxB);
                        lp += slope * wl; rp += slope * wr;
                        slope = mc_slope(rhom, rho, rhop, cF, cB, dxF, dxB);
                        lr += slope * wl; rr += slope * wr;
                        slope = mc_slope(Pv[i - 1], v, Pv[i + 1], cF, cB, dxF, dxB);
                        lv += slope * wl; rv += slope * wr;
                        if ((lp < 0) || (rp < 0) || (lr < 0) || (rr < 0)) { lr = rr = rho; lv = rv = v; lp = rp = pr0; }
                    }
                    Lr[i] = lr; Lv[i] = lv; Lp[i] = lp; Rr[i] = rr; Rv[i] = rv; Rp[i] = rp;
                }
            }
            __syncthreads();
            // ---- interface fluxes, stored over the (now dead) cell prims ----------------------------------------
#pragma unroll
            for (int k = 0; k < CPT; k++) {
                int j = tid + k * NT;
                if (j <= N) {
                    double f1, f2, f3;
                    if (is_gas) hllc_flux(Rr[j], Rv[j], Rp[j], Lr[j + 1], Lv[j + 1], Lp[j + 1], gm1, f1, f2, f3);
                    else dust_flux(Rr[j], Rv[j], Rp[j], Lr[j + 1], Lv[j + 1], Lp[j + 1], gm1, f1, f2, f3);
                    Pr[j] = f1; Pv[j] = f2; Pp[j] = f3;
                }
            }
            __syncthreads();
            // ---- sources, dudt, update ------------------------------------------------------------------------
#pragma unroll
            for (int k = 0; k < CPT; k++) {
                int j = tid + k * NT;
                double t1 = 0, t2 = 0, t3 = 0;
                if (j >= 1 && j <= N) {
                    const double dxj = __ldg(p.dx + j);
                    const double dphidr_p = (phis[j] - phis[j - 1]) / (dxj + __ldg(p.dx + j - 1));
                    const double dphidr_m = (phis[j + 1] - phis[j]) / (__ldg(p.dx + j + 1) + dxj);
                    const double g = __ldg(p.omegaplus + j) * dphidr_p + __ldg(p.omegaminus + j) * dphidr_m;
                    const double s1 = (0. * (-1.)) * g, s2 = (u1[k] * (-1.)) * g, s3 = (u2[k] * (-1.)) * g;
                    const double sp2 = -(__ldg(p.spp_left + j) * Lp[j] - __ldg(p.spp_right + j) * Rp[j]);
                    const double sl = __ldg(p.surf + j - 1), sr = __ldg(p.surf + j), vj = __ldg(p.vol + j);
                    t1 = (Pr[j - 1] * sl - Pr[j] * sr) / vj + (s1 + 0.);
                    t2 = (Pv[j - 1] * sl - Pv[j] * sr) / vj + (s2 + sp2);
                    t3 = (Pp[j - 1] * sl - Pp[j] * sr) / vj + (s3 + 0.);
                }
                if (stage == 0) {
                    d1[k] = t1; d2[k] = t2; d3[k] = t3;
                    u1[k] = u1[k] + t1 * dt; u2[k] = u2[k] + t2 * dt; u3[k] = u3[k] + t3 * dt;
                } else {
                    u1[k] += (t1 - d1[k]) * dt / 2; u2[k] += (t2 - d2[k]) * dt / 2; u3[k] += (t3 - d3[k]) * dt / 2;
                }
            }
            if (p.use_self_gravity && stage == 0 && nstage == 2) {      // second update_mass_and_pot() at t + dt (advection.cpp:168-169):
                __syncthreads();                                         // changes base->phi (the stage-2 boundary conditions read it), not phi_s
#pragma unroll
                for (int k = 0; k < CPT; k++) { const int j = tid + k * NT; if (j < N2) Pv[j] = u1[k]; }
                __syncthreads();
                self_gravity_update(p, col, Pr, Pv, Pp);
            }
        }
        // ---- write back; final compute_pressure (advection.cpp:204-205) and tail -----------------------------
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            int j = tid + k * NT;
            if (j < N2) {
                p.u[idx_u(p, 0, col, s, j)] = u1[k]; p.u[idx_u(p, 1, col, s, j)] = u2[k]; p.u[idx_u(p, 2, col, s, j)] = u3[k];
                if (j == 2) p.prim2v[(size_t)col * p.S + s] = u2[k] / u1[k];
                if (p.last_kernel == 0) {
                    Prim3 q = cons_to_prim(u1[k], u2[k], u3[k], gm1);
                    double e = eint_of(q.p, q.rho, gm1);
                    tail_cell(p, col, s, j, q.v, cs_of(e, gm1), e, e / sc.cv, acc);
                }
            }
        }
        __syncthreads();
    }
}

// Step head for one column (thread 0 decides, everybody learns through shared memory): is the column still running,
// and with which dt (advection.cpp:60-77).  Returns 0 when the column is finished / crashed / at its step ceiling.
struct StepInfo { double dt, t_old; int active, snap; };
__device__ __forceinline__ StepInfo step_head_on(const DevParams& p, ColState& c) {
    __shared__ double s_dt, s_t;
    __shared__ int s_active, s_snap;
    __syncthreads();
    if (threadIdx.x == 0) {
        int active = (c.status == COL_OK) && (c.t < c.t_max) && (c.steps < p.step_limit);
        s_active = active;
        c.in_step = active;
        if (active) {
            s_t = c.t;
            s_dt = step_head_dt(p, c);
            s_snap = (c.steps <= 1);
            c.t = c.t + s_dt;                          // advection.cpp:162,166
        }
    }
    __syncthreads();
    StepInfo si; si.active = s_active; si.dt = si.t_old = 0; si.snap = 0;
    if (si.active) { si.dt = s_dt; si.t_old = s_t; si.snap = s_snap; }
    return si;
}
__device__ __forceinline__ StepInfo step_head(const DevParams& p, const int col) { return step_head_on(p, p.cs[col]); }

// NTB = largest block the instantiation is launched with: 256 (two CTAs per SM) for columns of up to 2046 cells, 1024 for longer
// ones.  When the nine work arrays do not fit in shared memory they are taken from the column's HBM scratch area instead (the
// code below only ever touches them between __syncthreads(), with plain loads and stores).
template <int CPT, int NTB = 256>
#ifndef AIO_NO_SHARED_KERNELS      // defined by the translation units that only borrow the device functions
__global__ void __launch_bounds__(NTB, NTB == 256 ? 2 : 1) k_hydro(DevParams p) {
    AIO_DYN_SMEM(double, sm_dyn);
    const int col = blockIdx.x;
    double* sm = p.scratch_hydro ? p.scratch + (size_t)col * p.scratch_stride : sm_dyn;
    const StepInfo si = step_head(p, col);
    if (!si.active) return;
    TailAcc acc; acc.minstep = 0; acc.maxde = 0; acc.bad = 0; acc.badval = 0;
    hydro_phase<CPT>(p, col, sm, si.dt, si.t_old, si.snap, acc);
    if (p.last_kernel == 0) tail_commit(p, col, acc, sm + 9 * p.N2);
}
#endif

// Initial values of the dt inputs (what c_Sim's ctor + first get_cfl_timestep() see: init_and_bounds.cpp:678-682).
__global__ void k_prime(DevParams p) {
    __shared__ double red[32];
    const int col = blockIdx.x;
    double ms = 0, md = 0;
    for (int s = 0; s < p.S; s++) {
        const SpeciesConst sc = p.sp[s];
        for (int j = threadIdx.x; j < p.N2; j += blockDim.x) {
            Prim3 q = cons_to_prim(p.u[idx_u(p, 0, col, s, j)], p.u[idx_u(p, 1, col, s, j)], p.u[idx_u(p, 2, col, s, j)], sc.gm1);
            double e = eint_of(q.p, q.rho, sc.gm1), cs = cs_of(e, sc.gm1);
            if (j >= 1 && j <= p.N - 1) md = fmax(md, fabs(p.eint[idx_s(p, col, s, j)] - e) / e);
            if (j >= 1 && j <= p.N) { double dxj = p.dx[j]; ms = fmax(ms, fabs(q.v / dxj) + cs / dxj); }
            if (j <= p.N) p.eint[idx_s(p, col, s, j)] = e;      // primlast <- prim at the start of step 0 (advection.cpp:88-91)
        }
    }
    ms = block_max(ms, red); md = block_max(md, red);
    if (threadIdx.x == 0) { p.cs[col].minstep = ms; p.cs[col].maxde = md; }
}

}  // namespace aio
