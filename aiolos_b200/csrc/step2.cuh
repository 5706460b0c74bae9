// Second-generation per-column kernels ("v2"): the same algorithms as hydro.cuh / rad_simple.cuh (whose functions cite
// the reference line by line and stay available as engine option kernels=v1), re-expressed for the B200's issue and
// FP64 pipes after the round-1a ncu capture (profiles/r1a_*: issue-bound at 40 %, a third of the issued instructions
// FP64, a quarter the branchy scaffolding of IEEE division/sqrt, a quarter 64-bit address arithmetic):
//   * divisions and square roots are branch-free MUFU seed + Newton sequences (frcp/fsqrt, <= ~1.5 ulp); the results agree
//     with the reference within the stated tolerance, not bit for bit;
//   * everything that depends only on the grid is tabulated once on the host (slope-limiter factors, 1/vol, ...), and
//     everything that depends only on the (static) potential once per column on the device (hydrostatic pressure
//     offsets dp-, dp+ and the gravity source coefficient);
//   * HLLC evaluates one star-state formula on the side the reference's branch order selects instead of four branches;
//   * each interface diffusion coefficient of the FLD system is computed once and shared through shared memory;
//   * the scalar tridiagonal system is solved by substructuring: every thread eliminates the interior of its chunk of
//     consecutive rows in registers (two-sweep Thomas), the 1-row-per-thread separator system goes through cyclic
//     reduction, and the interiors are back-substituted in registers.
#pragma once
#include "common.cuh"
#include "hydro.cuh"
#include "rad_simple.cuh"

#ifndef AIO_STEP1_MINB
#define AIO_STEP1_MINB 3
#endif

namespace aio {

// loops that only touch shared memory: partial unrolling keeps the hot code in the instruction cache
constexpr int kUnrollSmem = 8;       // measured: 2 is 4 % slower than full unrolling (profiles/r1f_notes.md)

// ---- branch-free reciprocal / square root ----------------------------------------------------------------------------------
#ifndef AIO_EMU_BUILD
__device__ __forceinline__ double frcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
__device__ __forceinline__ double fsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    double e = fma(-hx * y, y, 0.5);          // 0.5 - 0.5 x y^2
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    double s = x * y;
    const double r = fma(-s, s, x);
    s = fma(r, 0.5 * y, s);
    return (x > 0.0) ? s : ((x == 0.0) ? 0.0 : s);   // x == 0: y = inf, s = NaN -> 0 ; x < 0: NaN propagates like sqrt
}
#else
inline double frcp(double x) { return 1.0 / x; }
inline double fsqrt(double x) { return std::sqrt(x); }
#endif
__device__ __forceinline__ double fdiv(double a, double b) { return a * frcp(b); }

// Per-cell tail (helpers.cpp:28-71, advection.cpp:88-91,282-307) with branch-free reciprocals and the tabulated 1/dx.
__device__ __forceinline__ void tail_cell2(const DevParams& p, double* __restrict__ eint_row, int j, double v, double cs, double eint,
                                           double T, TailAcc& a) {
    if (j >= 1 && j <= p.N - 1) a.maxde = fmax(a.maxde, fabs(eint_row[j] - eint) * frcp(eint));
    if (j >= 1 && j <= p.N) { const double ix = __ldg(p.g_idx + j); a.minstep = fmax(a.minstep, fabs(v * ix) + cs * ix); }
    if (j <= p.N - 1 && (T < 0 || isnan(T))) { a.bad = 1; a.badval = T; }
    if (j <= p.N) eint_row[j] = eint;
}

// ---- static per-column tables: dp-(i), dp+(i) (aiolos.h:1003-1016) and the gravity source coefficient (source.cpp:157-164) ------
__device__ inline void build_static_tables(const DevParams& p, const int col) {
    const int tid = threadIdx.x, NT = blockDim.x, N = p.N, N2 = p.N2;
    for (int s = 0; s < p.S; s++) {
        const double* phis = p.phis + idx_s(p, col, s, 0);
        double* dpm = p.dpm + idx_s(p, col, s, 0);
        double* dpp = p.dpp + idx_s(p, col, s, 0);
        double* gs = p.gsrc + idx_s(p, col, s, 0);
        for (int i = tid; i < N2; i += NT) {
            double a = 0, b = 0, g = 0;
            if (i > 0) a = -p.dx[i] * p.omegaplus[i] * (phis[i - 1] - phis[i]) / (p.dx[i - 1] + p.dx[i]);
            if (i <= N) b = -p.dx[i] * p.omegaminus[i] * (phis[i + 1] - phis[i]) / (p.dx[i + 1] + p.dx[i]);
            if (i >= 1 && i <= N) {
                const double dphidr_p = (phis[i] - phis[i - 1]) / (p.dx[i] + p.dx[i - 1]);
                const double dphidr_m = (phis[i + 1] - phis[i]) / (p.dx[i + 1] + p.dx[i]);
                g = p.omegaplus[i] * dphidr_p + p.omegaminus[i] * dphidr_m;
            }
            dpm[i] = a; dpp[i] = b; gs[i] = g;
        }
    }
}

// MC slope with tabulated inverse distances (reconstruct.cpp:11-29)
__device__ __forceinline__ double mc_slope2(double ql, double qm, double qr, double cF, double cB, double idxF, double idxB) {
    const double dF = (qr - qm) * idxF;
    const double dB = (qm - ql) * idxB;
    const double a = cF * dF, b = cB * dB;
    const double m = (fabs(a) < fabs(b)) ? a : b;
    const double c = 0.5 * (dF + dB);
    const double r = (fabs(c) < fabs(m)) ? c : m;
    return (dF * dB < 0) ? 0.0 : r;
}

// HLLC (advection.cpp:544-615): the star-state formula is evaluated once, on the side the reference's branch order picks.
__device__ __forceinline__ void hllc_flux2(double dl, double ul, double pl, double dr, double ur, double pr, double gm1, double igm1,
                                           double& f1, double& f2, double& f3) {
    const double rdl = frcp(dl), rdr = frcp(dr);
    const double el = pl * rdl * igm1, er = pr * rdr * igm1;
    const double g2 = (gm1 + 1) * gm1;
    const double csl = fsqrt(g2 * el), csr = fsqrt(g2 * er);
    const double mom_l = dl * ul, mom_r = dr * ur;
    const double El = dl * el + 0.5 * mom_l * ul, Er = dr * er + 0.5 * mom_r * ur;
    const double SL = ul - csl, SR = ur + csr;
    const double al = SL - ul, ar = SR - ur;
    const double SS = (pr - pl + mom_l * al - mom_r * ar) * frcp(dl * al - dr * ar);
    const bool isA = (SL <= 0) && (SS >= 0);
    const bool isB = !isA && (SS <= 0) && (SR >= 0);
    const bool left = isA || (!isB && (SL >= 0));
    // side-K quantities
    const double dK = left ? dl : dr, uK = left ? ul : ur, pK = left ? pl : pr, mK = left ? mom_l : mom_r;
    const double EK = left ? El : Er, SK = left ? SL : SR, aK = left ? al : ar, rdK = left ? rdl : rdr;
    const double F1 = mK, F2 = mK * uK + pK, F3 = uK * (EK + pK);
    double r1 = F1, r2 = F2, r3 = F3;
    if (isA || isB) {
        const double comp3 = EK * rdK + (SS - uK) * (SS + pK * frcp(dK * aK));
        const double w = dK * aK * frcp(SK - SS);
        r1 = F1 + (w - dK) * SK;
        r2 = F2 + (SS * w - mK) * SK;
        r3 = F3 + (comp3 * w - EK) * SK;
    } else if (!(SL >= 0) && !(SR <= 0)) { r1 = 0; r2 = 0; r3 = 0; }      // NaN inputs: the reference returns a zero flux
    f1 = r1; f2 = r2; f3 = r3;
}

// Both Heun stages of one species on register-resident cells (u1,u2,u3 of the cells this thread owns); `sm` as in hydro_phase_v2.
template <int CPT, int NTC = 0>
__device__ __forceinline__ void hydro_stages_v2(const DevParams& p, const int col, const int s, const SpeciesConst& sc, double* sm,
                                                const double dt, const double t_old, double (&u1)[CPT], double (&u2)[CPT],
                                                double (&u3)[CPT], double& prim2_speed) {
    const int tid = threadIdx.x; const int NT = NTC > 0 ? NTC : (int)blockDim.x;
    const int N = p.N, N2 = p.N2, ng = p.ng;
    const int LDH = NTC > 0 ? NTC * CPT : N2;      // array stride: a compile-time constant in the specialised instantiation
    double* Pr = sm;            double* Pv = Pr + LDH;  double* Pp = Pv + LDH;
    double* Lr = Pp + LDH;      double* Lv = Lr + LDH;  double* Lp = Lv + LDH;
    double* Rr = Lp + LDH;      double* Rv = Rr + LDH;  double* Rp = Rv + LDH;
    double* win = Rp + LDH + 32;
    BndWin W; W.w = win; W.ng = ng; W.N2 = N2;
    const bool second = (p.order == 2);
    const double gm1 = sc.gm1, igm1 = 1.0 / sc.gm1;
    const bool is_gas = !sc.is_dust_like;
    // plain (coherent) loads: these tables are rewritten by this CTA at steps 0 and 1, possibly within the same launch, so the
    // non-coherent read-only path (__ldg / const __restrict__) must not be used on them
    const double* DPM = p.dpm + idx_s(p, col, s, 0);
    const double* DPP = p.dpp + idx_s(p, col, s, 0);
    const double* GS = p.gsrc + idx_s(p, col, s, 0);
    double d1[CPT], d2[CPT], d3[CPT];
#pragma unroll
    for (int k = 0; k < CPT; k++) d1[k] = d2[k] = d3[k] = 0;
    const int nstage = second ? 2 : 1;
    for (int stage = 0; stage < nstage; stage++) {
        // ---- boundary conditions (init_and_bounds.cpp:1779-1987) on the 2*ng cells next to each end ------------------------
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int j = tid + k * NT;
            if (j < N2 && (j < 2 * ng || j >= N2 - 2 * ng)) { W.at(0, j) = u1[k]; W.at(1, j) = u2[k]; W.at(2, j) = u3[k]; }
        }
        __syncthreads();
        if (tid == 0) {
            apply_boundary_left<false>(p, sc, col, s, W, prim2_speed);
            if (sc.use_wave == 1) {
                const double tt = (stage == 0) ? t_old : t_old + dt;
                W.at(1, 0) = W.at(0, 0) * sc.wave_amplitude * sin(12. * M_PI * tt / sc.wave_period);
                W.at(1, 1) = W.at(1, 0);
            }
        }
        if (tid == ((NT > 32) ? 32 : 0)) apply_boundary_right<false>(p, sc, col, s, W);
        __syncthreads();
        // ---- cons -> prim (eos.cpp:19-29) --------------------------------------------------------------------------------------
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int j = tid + k * NT;
            if (j < N2) {
                if (j < 2 * ng || j >= N2 - 2 * ng) { u1[k] = W.at(0, j); u2[k] = W.at(1, j); u3[k] = W.at(2, j); }
                const double v = u2[k] * frcp(u1[k]);
                Pr[j] = u1[k]; Pv[j] = v; Pp[j] = gm1 * (u3[k] - 0.5 * u2[k] * v);
            }
        }
        __syncthreads();
        prim2_speed = Pv[2 < N2 ? 2 : 0];
        // ---- well-balanced reconstruction (reconstruct.cpp:34-117) ----------------------------------------------------------------
#pragma unroll kUnrollSmem
        for (int k = 0; k < CPT; k++) {
            const int i = tid + k * NT;
            if (i < N2) {
                const double rho = Pr[i], v = Pv[i], pr0 = Pp[i];
                double lr = rho, lv = v, lp = pr0, rr = rho, rv = v, rp = pr0;
                double dpm = 0, dpp = 0;
                if (is_gas) {
                    dpm = DPM[i]; dpp = DPP[i];
                    if (i > 0) lp = pr0 + rho * dpm;
                    if (i <= N) rp = pr0 + rho * dpp;
                    if (lp < 0 || rp < 0) { lp = pr0; rp = pr0; }
                }
                if (second && i >= 1 && i <= N) {
                    const double rhom = Pr[i - 1], rhop = Pr[i + 1];
                    double dp_l = 0, dp_r = 0;
                    if (is_gas) {
                        dp_l = rho * dpm - rhom * DPP[i - 1];
                        dp_r = rho * dpp - rhop * DPM[i + 1];
                    }
                    const double cF = __ldg(p.g_cF + i), cB = __ldg(p.g_cB + i), ixF = __ldg(p.g_idxF + i), ixB = __ldg(p.g_idxB + i);
                    const double wl = __ldg(p.g_wl + i), wr = __ldg(p.g_wr + i);
                    double slope = mc_slope2(Pp[i - 1] - dp_l, pr0, Pp[i + 1] - dp_r, cF, cB, ixF, ixB);
                    lp += slope * wl; rp += slope * wr;
                    slope = mc_slope2(rhom, rho, rhop, cF, cB, ixF, ixB);
                    lr += slope * wl; rr += slope * wr;
                    slope = mc_slope2(Pv[i - 1], v, Pv[i + 1], cF, cB, ixF, ixB);
                    lv += slope * wl; rv += slope * wr;
                    if ((lp < 0) || (rp < 0) || (lr < 0) || (rr < 0)) { lr = rr = rho; lv = rv = v; lp = rp = pr0; }
                }
                Lr[i] = lr; Lv[i] = lv; Lp[i] = lp; Rr[i] = rr; Rv[i] = rv; Rp[i] = rp;
            }
        }
        __syncthreads();
        // ---- interface fluxes, stored over the (now dead) cell prims ---------------------------------------------------------------
#pragma unroll kUnrollSmem
        for (int k = 0; k < CPT; k++) {
            const int j = tid + k * NT;
            if (j <= N) {
                double f1, f2, f3;
                if (is_gas) hllc_flux2(Rr[j], Rv[j], Rp[j], Lr[j + 1], Lv[j + 1], Lp[j + 1], gm1, igm1, f1, f2, f3);
                else dust_flux(Rr[j], Rv[j], Rp[j], Lr[j + 1], Lv[j + 1], Lp[j + 1], gm1, f1, f2, f3);
                Pr[j] = f1; Pv[j] = f2; Pp[j] = f3;
            }
        }
        __syncthreads();
        // ---- sources, dudt, update (source.cpp:157-164, advection.cpp:492-503, 154-158, 187-195) ---------------------------------------
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int j = tid + k * NT;
            double t1 = 0, t2 = 0, t3 = 0;
            if (j >= 1 && j <= N) {
                const double g = GS[j];
                const double s2 = (u1[k] * (-1.)) * g, s3 = (u2[k] * (-1.)) * g;
                const double sp2 = -(__ldg(p.spp_left + j) * Lp[j] - __ldg(p.spp_right + j) * Rp[j]);
                const double sl = __ldg(p.surf + j - 1), sr = __ldg(p.surf + j), iv = __ldg(p.g_ivol + j);
                t1 = (Pr[j - 1] * sl - Pr[j] * sr) * iv;
                t2 = (Pv[j - 1] * sl - Pv[j] * sr) * iv + (s2 + sp2);
                t3 = (Pp[j - 1] * sl - Pp[j] * sr) * iv + s3;
            }
            if (stage == 0) {
                d1[k] = t1; d2[k] = t2; d3[k] = t3;
                u1[k] = u1[k] + t1 * dt; u2[k] = u2[k] + t2 * dt; u3[k] = u3[k] + t3 * dt;
            } else {
                u1[k] += (t1 - d1[k]) * dt * 0.5; u2[k] += (t2 - d2[k]) * dt * 0.5; u3[k] += (t3 - d3[k]) * dt * 0.5;
            }
        }
    }
}

// phi_s snapshot + static tables (first two steps only, see hydro.cuh)
__device__ inline void hydro_snapshot_v2(const DevParams& p, const int col) {
    const int tid = threadIdx.x, NT = blockDim.x, N2 = p.N2;
    const double* phi = p.phi + idx_c(p, col, 0);
    for (int s = 0; s < p.S; s++) {
        double* phis = p.phis + idx_s(p, col, s, 0);
        for (int j = tid; j < N2; j += NT) phis[j] = phi_snapshot(p, col, phi, j) * 1.0;
    }
    __syncthreads();
    build_static_tables(p, col);
    __syncthreads();
}

template <int CPT>
__device__ __forceinline__ void hydro_phase_v2(const DevParams& p, const int col, double* sm, const double dt, const double t_old,
                                               const int snap, TailAcc& acc) {
    const int tid = threadIdx.x, NT = blockDim.x;
    const int N = p.N, N2 = p.N2, ng = p.ng;
    double* Pr = sm;            double* Pv = Pr + N2;  double* Pp = Pv + N2;   // cell prims; reused for the fluxes
    double* Lr = Pp + N2;       double* Lv = Lr + N2;  double* Lp = Lv + N2;   // left edge states  prim_l
    double* Rr = Lp + N2;       double* Rv = Rr + N2;  double* Rp = Rv + N2;   // right edge states prim_r
    double* win = Rp + N2 + 32;                                                // 3*4*ng doubles (after the reduction scratch)
    BndWin W; W.w = win; W.ng = ng; W.N2 = N2;

    if (snap) {
        const double* phi = p.phi + idx_c(p, col, 0);
        for (int s = 0; s < p.S; s++) {
            double* phis = p.phis + idx_s(p, col, s, 0);
            for (int j = tid; j < N2; j += NT) phis[j] = phi_snapshot(p, col, phi, j) * 1.0;
        }
        __syncthreads();
        build_static_tables(p, col);
        __syncthreads();
    }
    const bool second = (p.order == 2);
    for (int s = 0; s < p.S; s++) {
        const SpeciesConst sc = p.sp[s];
        const double gm1 = sc.gm1, igm1 = 1.0 / sc.gm1;
        const bool is_gas = !sc.is_dust_like;
        double* __restrict__ U1 = p.u + idx_u(p, 0, col, s, 0);
        double* __restrict__ U2 = p.u + idx_u(p, 1, col, s, 0);
        double* __restrict__ U3 = p.u + idx_u(p, 2, col, s, 0);
        double u1[CPT], u2[CPT], u3[CPT];
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int j = tid + k * NT;
            if (j < N2) { u1[k] = U1[j]; u2[k] = U2[j]; u3[k] = U3[j]; }
            else { u1[k] = 1; u2[k] = 0; u3[k] = 1; }
        }
        double prim2_speed = p.prim2v[(size_t)col * p.S + s];
        hydro_stages_v2<CPT>(p, col, s, sc, sm, dt, t_old, u1, u2, u3, prim2_speed);
        // ---- write back; final compute_pressure (advection.cpp:204-205) and tail ---------------------------------------------------------
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int j = tid + k * NT;
            if (j < N2) {
                U1[j] = u1[k]; U2[j] = u2[k]; U3[j] = u3[k];
                if (j == 2) p.prim2v[(size_t)col * p.S + s] = u2[k] / u1[k];
                if (p.last_kernel == 0) {
                    const double rin = frcp(u1[k]);
                    const double v = u2[k] * rin;
                    const double pr = gm1 * (u3[k] - 0.5 * u2[k] * v);
                    const double e = pr * rin * igm1;
                    tail_cell2(p, p.eint + idx_s(p, col, s, 0), j, v, fsqrt((gm1 + 1) * gm1 * e), e, e / sc.cv, acc);
                }
            }
        }
        __syncthreads();
    }
}

#ifndef AIO_HYDRO2_MINB
#define AIO_HYDRO2_MINB 3
#endif
template <int CPT>
__global__ void __launch_bounds__(256, AIO_HYDRO2_MINB) k_hydro2(DevParams p) {
    AIO_DYN_SMEM(double, sm);
    const int col = blockIdx.x;
    const StepInfo si = step_head(p, col);
    if (!si.active) return;
    TailAcc acc; acc.minstep = 0; acc.maxde = 0; acc.bad = 0; acc.badval = 0;
    hydro_phase_v2<CPT>(p, col, sm, si.dt, si.t_old, si.snap, acc);
    if (p.last_kernel == 0) tail_commit(p, col, acc, sm + 9 * p.N2);
}

// ---- chunked shared-memory access: thread t owns elements [t*M, t*M+M); 16-byte accesses halve the bank conflicts of the
// 32-byte-strided pattern (arrays must be 16-byte aligned and hold NT*M elements) ---------------------------------------------
template <int M>
__device__ __forceinline__ void load_chunk(const double* a, int r0, double (&v)[M]) {
    if constexpr (M % 2 == 0) {
        const double2* a2 = reinterpret_cast<const double2*>(a + r0);
#pragma unroll
        for (int k = 0; k < M / 2; k++) { const double2 t = a2[k]; v[2 * k] = t.x; v[2 * k + 1] = t.y; }
    } else {
#pragma unroll
        for (int k = 0; k < M; k++) v[k] = a[r0 + k];
    }
}
template <int M>
__device__ __forceinline__ void store_chunk(double* a, int r0, const double (&v)[M]) {
    if constexpr (M % 2 == 0) {
        double2* a2 = reinterpret_cast<double2*>(a + r0);
#pragma unroll
        for (int k = 0; k < M / 2; k++) a2[k] = make_double2(v[2 * k], v[2 * k + 1]);
    } else {
#pragma unroll
        for (int k = 0; k < M; k++) a[r0 + k] = v[k];
    }
}

// Inclusive scan of a[0..n) (array of NT*M elements, entries >= n ignored), chunk per thread, in place.
template <int M, int NTC = 0>
__device__ inline void block_inclusive_scan_chunked(double* a, const int n, double* red) {
    const int tid = threadIdx.x, r0 = tid * M; const int NT = NTC > 0 ? NTC : (int)blockDim.x;
    double v[M];
    load_chunk<M>(a, r0, v);
    double loc = 0;
#pragma unroll
    for (int k = 0; k < M; k++) { loc += (r0 + k < n) ? v[k] : 0.0; v[k] = loc; }
    double s = loc;
    const int lane = tid & 31, w = tid >> 5;
    for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
    __syncthreads();
    if (lane == 31) red[w] = s;
    __syncthreads();
    double woff = 0;
    for (int k = 0; k < w; k++) woff += red[k];
    const double off = woff + (s - loc);
#pragma unroll
    for (int k = 0; k < M; k++) v[k] += off;
    store_chunk<M>(a, r0, v);
    __syncthreads();
    (void)NT;
}

// ---- tridiagonal solve by substructuring ---------------------------------------------------------------------------------------------
// Rows [r0, r0+M) belong to thread t (r0 = t*M); the last row of each chunk is a separator.  A,B,C,D: sub, diag, super, rhs in
// shared memory (n rows, n <= NT*M); solution left in D.  `red` needs 4*NT doubles of scratch.
template <int M, int NTC = 0>
__device__ inline void tridiag_substructured(double* A, double* B, double* C, double* D, const int n, double* red, const bool chunked = false) {
    const int tid = threadIdx.x; const int NT = NTC > 0 ? NTC : (int)blockDim.x;
    const int r0 = tid * M;
    double a[M], b[M], c[M], d[M];
    if (chunked) { load_chunk<M>(A, r0, a); load_chunk<M>(B, r0, b); load_chunk<M>(C, r0, c); load_chunk<M>(D, r0, d); }
#pragma unroll
    for (int k = 0; k < M; k++) {
        const int i = r0 + k;
        if (i < n) { if (!chunked) { a[k] = A[i]; b[k] = B[i]; c[k] = C[i]; d[k] = D[i]; } }
        else { a[k] = 0; b[k] = 1; c[k] = 0; d[k] = 0; }
    }
    if (r0 + M - 1 >= n - 1 && r0 < n) {     // the chunk holding the last row: its trailing rows are identity rows
        // make sure the true last row does not couple beyond the system
#pragma unroll
        for (int k = 0; k < M; k++) if (r0 + k == n - 1) c[k] = 0;
    }
    // interior rows 0..M-2: two-sweep elimination -> a[k] x_{L(t-1)} + b[k] x_k + c[k] x_{L(t)} = d[k]
#pragma unroll
    for (int k = 1; k < M - 1; k++) {
        const double w = a[k] * frcp(b[k - 1]);
        b[k] = fma(-w, c[k - 1], b[k]);
        d[k] = fma(-w, d[k - 1], d[k]);
        a[k] = -w * a[k - 1];
    }
#pragma unroll
    for (int k = M - 3; k >= 0; k--) {
        const double w = c[k] * frcp(b[k + 1]);
        d[k] = fma(-w, d[k + 1], d[k]);
        a[k] = fma(-w, a[k + 1], a[k]);
        c[k] = -w * c[k + 1];
    }
    // publish the first interior row of this chunk for the left neighbour's separator
    double* Fa = red; double* Fb = Fa + NT; double* Fc = Fb + NT; double* Fd = Fc + NT;
    __syncthreads();                        // everybody has read A..D
    Fa[tid] = a[0]; Fb[tid] = b[0]; Fc[tid] = c[0]; Fd[tid] = d[0];
    __syncthreads();
    // separator row L(t): A x_{int last} + B x_L + C x_{first interior of t+1} = D
    double sa = a[M - 1], sb = b[M - 1], sc_ = c[M - 1], sd = d[M - 1];
    if constexpr (M >= 2) {
        const double wl = sa * frcp(b[M - 2]);                  // last interior row of this chunk
        double na = 0, nb = 1, nc = 0, nd = 0;                  // first interior row of the next chunk
        if (tid + 1 < NT) { na = Fa[tid + 1]; nb = Fb[tid + 1]; nc = Fc[tid + 1]; nd = Fd[tid + 1]; }
        const double wr = sc_ * frcp(nb);
        sa = -wl * a[M - 2];
        sb = sb - wl * c[M - 2] - wr * na;
        sc_ = -wr * nc;
        sd = sd - wl * d[M - 2] - wr * nd;
    }
    // separator system (one row per thread) by parallel cyclic reduction in shared memory
    double* SA = Fa; double* SB = Fb; double* SC = Fc; double* SD = Fd;      // the published rows are consumed: reuse the scratch
    __syncthreads();
    SA[tid] = sa; SB[tid] = sb; SC[tid] = sc_; SD[tid] = sd;
    __syncthreads();
    for (int s = 1; s < NT; s <<= 1) {
        const int im = tid - s, ip = tid + s;
        double ai = SA[tid], bi = SB[tid], ci = SC[tid], di = SD[tid];
        double na = 0, nc = 0;
        if (im >= 0) { const double k1 = ai * frcp(SB[im]); na = -SA[im] * k1; bi = fma(-SC[im], k1, bi); di = fma(-SD[im], k1, di); }
        if (ip < NT) { const double k2 = ci * frcp(SB[ip]); nc = -SC[ip] * k2; bi = fma(-SA[ip], k2, bi); di = fma(-SD[ip], k2, di); }
        __syncthreads();
        SA[tid] = na; SB[tid] = bi; SC[tid] = nc; SD[tid] = di;
        __syncthreads();
    }
    const double xL = SD[tid] * frcp(SB[tid]);
    __syncthreads();
    Fa[tid] = xL;
    __syncthreads();
    const double xLm = (tid > 0) ? Fa[tid - 1] : 0.0;
    // back-substitute the interiors and write the solution
    double x[M];
#pragma unroll
    for (int k = 0; k < M; k++) x[k] = (k == M - 1) ? xL : (d[k] - a[k] * xLm - c[k] * xL) * frcp(b[k]);
    if (chunked) store_chunk<M>(D, r0, x);
    else {
#pragma unroll
        for (int k = 0; k < M; k++) if (r0 + k < n) D[r0 + k] = x[k];
    }
    __syncthreads();
}

template <int CPT>
__device__ __forceinline__ void rad_simple_phase_v2(const DevParams& p, const int col, double* sm, const double ddt, const double tnow) {
    const int tid = threadIdx.x, NT = blockDim.x;
    const int N = p.N, N2 = p.N2, S = p.S, Bin = p.Bin;
    double* RHO = sm;                  // [S][N2]
    double* TT = RHO + S * N2;         // [S][N2] temperature
    double* DS = TT + S * N2;          // [S][N2]
    double* E1 = DS + S * N2;          // [S][N2] eta1
    double* DEN = E1 + S * N2;         // [S][N2]
    double* TO = DEN + S * N2;         // total_opacity (out band 0)
    double* JJ = TO + N2;
    double* TAU = JJ + N2;             // scan workspace, then the interface diffusion coefficients
    double* A = TAU + N2; double* B = A + N2; double* C = B + N2; double* D = C + N2;
    double* red = D + N2;              // 32 + 4*NT
    const int KA = Bin + 2 * p.Bout;

    for (int s = 0; s < S; s++) {
        const SpeciesConst sc = p.sp[s];
        const double igm1 = 1.0 / sc.gm1, icv = 1.0 / sc.cv;
        const double* U1 = p.u + idx_u(p, 0, col, s, 0); const double* U2 = p.u + idx_u(p, 1, col, s, 0); const double* U3 = p.u + idx_u(p, 2, col, s, 0);
        for (int j = tid; j < N2; j += NT) {
            const double a1 = U1[j], a2 = U2[j], a3 = U3[j];
            const double rin = frcp(a1), v = a2 * rin;
            const double pr = sc.gm1 * (a3 - 0.5 * a2 * v);
            RHO[s * N2 + j] = a1;
            TT[s * N2 + j] = p.Tsave ? p.Tsave[idx_s(p, col, s, j)] : pr * rin * igm1 * icv;
            DS[s * N2 + j] = 0.;
        }
    }
    {
        const double* Jg = p.J + idx_J(p, col, 0, 0);
        for (int j = tid; j < N2; j += NT) JJ[j] = Jg[j];
    }
    __syncthreads();

    // ---- stellar heating (radiation.cpp:43-348) -------------------------------------------------------------------------------------
    for (int b = 0; b < Bin; b++) {
        double F = p.solar_heating_final[(size_t)col * Bin + b];
        if (tnow < p.radiation_rampup_time)
            F = F * (p.init_radiation_factor + (1. - p.init_radiation_factor) * tnow / p.radiation_rampup_time);
        const double top_prev = p.dSband_top[(size_t)col * Bin + b];
        for (int q = tid; q < N; q += NT) {
            const int j = N - q;
            double tot = 0;
            for (int s = 0; s < S; s++) tot += p.kappa_avg[s * KA + b] * RHO[s * N2 + j];
            TAU[q] = tot * __ldg(p.dx + j);
        }
        __syncthreads();
        block_inclusive_scan(TAU, N, red);
        for (int j = tid + 1; j <= N + 1; j += NT) {
            const double dxj = __ldg(p.dx + j);
            double tot = 0;
            for (int s = 0; s < S; s++) tot += p.kappa_avg[s * KA + b] * RHO[s * N2 + j];
            const double cell_tau = tot * dxj;
            double dSb;
            if (j < N + 1) {
                const double rad_j = TAU[N - j];
                const double rad_jp = (j == N) ? 0. : TAU[N - j - 1];
                const double Sb = F * exp(-rad_jp);
                dSb = 0.25 * Sb * frcp(dxj) * (1. - p.bond_albedo);
                const double dtau_tot = -(rad_jp - rad_j);
                if (dtau_tot > 1e-3) dSb *= (-expm1(-dtau_tot));
                else dSb *= (-fastexpm1_2(-dtau_tot));
                dSb = smax(dSb, 0.);
            } else {
                dSb = top_prev;
            }
            if (j == N) p.dSband_top[(size_t)col * Bin + b] = dSb;
            const double ict = frcp(cell_tau);
            for (int s = 0; s < S; s++) {
                const double hsw = p.highenergy_switch[s * Bin + b];
                const double frac = hsw * p.kappa_avg[s * KA + b] * RHO[s * N2 + j] * dxj * ict;
                double v = DS[s * N2 + j] + hsw * dSb * frac;
                if (v < 1e-50) v = 0.;
                DS[s * N2 + j] = v;
            }
        }
        __syncthreads();
    }
    // ---- thermal band: total opacity; interface diffusion coefficients, each computed once ------------------------------------------------
    for (int j = tid; j < N2; j += NT) {
        double tot = 0;
        if (j >= 1) for (int s = 0; s < S; s++) tot = fmax(p.kappa_avg[s * KA + Bin + p.Bout + 0] * RHO[s * N2 + j], tot);
        TO[j] = tot;
    }
    __syncthreads();
    double* DF = TAU;                  // DF[j] = D across the right face of cell j, 1 <= j <= N
    for (int jj = tid; jj < N2; jj += NT) {
        double Dv = 0;
        if (jj >= 1 && jj <= N) {
            const double dxc = (__ldg(p.x_i12 + jj + 1) - __ldg(p.x_i12 + jj));
            const double t0 = TO[jj], t1 = TO[jj + 1];
            double rhokr = smax(2. * (t0 * t1) * frcp(t0 + t1), 4. / 3. * frcp(dxc));
            rhokr = smin(0.5 * (t0 + t1), rhokr);
            const double tau_inv = frcp(dxc * rhokr);
            const double dJ = fabs(JJ[jj + 1] - JJ[jj]) * frcp(JJ[jj] + 1e-300);
            const double R = p.xi_rad * tau_inv * dJ;
            const double fl = (R <= 2) ? 2 * frcp(3 + fsqrt(9 + 10 * R * R)) : 10 * frcp(10 * R + 9 + fsqrt(81 + 180 * R));
            Dv = tau_inv * __ldg(p.surf + jj) * fl;
        }
        DF[jj] = Dv;
    }
    __syncthreads();
    // ---- FLD-simple assembly (radiation_simple.cpp:44-233) ----------------------------------------------------------------------------------
    const double icdt = frcp(C_CLIGHT * ddt);
    for (int j = tid; j < N2; j += NT) {
        double l = 0, d = 0, u = 0, r = 0;
        if (j >= 2 && j <= N + 1) { const double Dm = DF[j - 1]; d = Dm; l = -Dm; }
        if (j <= N) {
            const double tv = __ldg(p.vol + j) * icdt;
            d += tv;
            r += tv * JJ[j];
        }
        if (j >= 1 && j <= N) { const double Dj = DF[j]; u = -Dj; d += Dj; }
        if (j < p.ng) { l = 0; u = -d; r = 0; }
        if (p.closed_radiative_boundaries) {
            if (j >= N2 - p.ng) { l = -l; u = 0; r = 0; }
        } else if (j == N) {
            double f = 1. / p.xi_rad * __ldg(p.x_i12 + N) / __ldg(p.x_i12 + N + 1);
            if (p.geometry == 0) f = 1;
            else if (p.geometry == 2) f *= f;
            l = -f * d; u = 0; r = 0;
        }
        if (j <= N) {
            double exd = 0, exr = 0;
            for (int s = 0; s < S; s++) {
                const SpeciesConst sc = p.sp[s];
                const double icv = 1.0 / sc.cv;
                const double Ts = TT[s * N2 + j], Ts3 = Ts * Ts * Ts, rhos = RHO[s * N2 + j];
                const double kappa = p.kappa_avg[s * KA + Bin + 0];
                const double fac = ddt * p.no_rad_trans * kappa * icv * C_SIGMA_RAD * Ts3;
                const double denom = 1. + 16. * fac;
                const double idn = frcp(denom);
                double eta1 = Ts * (1. + 12. * fac);
                eta1 += ddt * DS[s * N2 + j] * frcp(rhos) * icv;
                const double eta2 = 4. * C_PI * ddt * kappa * p.no_rad_trans * icv;
                E1[s * N2 + j] = eta1; DEN[s * N2 + j] = idn;           // DEN holds 1/denom in this version
                // same (cancelling) form as radiation_simple.cpp:210-211: its round-off is part of what the reference computes
                exd += p.no_rad_trans * rhos * kappa * (1 - 4. * C_SIGMA_RAD * Ts3 / C_PI * eta2 * idn);
                exr += p.no_rad_trans * rhos * kappa * C_SIGMA_RAD * Ts3 / C_PI * (4 * eta1 * idn - 3 * Ts);
            }
            d += __ldg(p.vol + j) * exd;
            r += __ldg(p.vol + j) * exr;
        }
        A[j] = l; B[j] = d; C[j] = u; D[j] = r;
    }
    __syncthreads();
    if (p.tridiag_serial) {
        if (tid == 0) {
            for (int i = 1; i < N2; i++) { const double m = A[i] / B[i - 1]; B[i] -= m * C[i - 1]; D[i] -= m * D[i - 1]; }
            D[N2 - 1] /= B[N2 - 1];
            for (int i = N2 - 2; i >= 0; i--) D[i] = (D[i] - C[i] * D[i + 1]) / B[i];
        }
        __syncthreads();
    } else {
        tridiag_substructured<CPT>(A, B, C, D, N2, red + 32);
    }
    // ---- J, T update, e <- T, p <- e, cons <- prim, step tail ------------------------------------------------------------------------------------
    int jneg = 0;
    for (int j = tid; j < N2; j += NT) {
        if (p.solve_for_j) JJ[j] = D[j];
        if (JJ[j] < 0. && p.couple_J_into_T) jneg = 1;
    }
    const int Jswitch = block_or(jneg) || (tnow > 1e15);
    TailAcc acc; acc.minstep = 0; acc.maxde = 0; acc.bad = 0; acc.badval = 0;
    int negJ = 0; double negJval = 0;
    double* Jg = p.J + idx_J(p, col, 0, 0);
    for (int j = tid; j < N2; j += NT) {
        double Jv = JJ[j];
        if (p.couple_J_into_T && j <= N - 1 && Jv < 0) {
            if (Jv > -1e-10) Jv *= -1.;
            else if (!(j < N && p.use_planetary_temperature && JJ[N] > 0)) { negJ = 1; negJval = Jv; }
        }
        Jg[j] = Jv;
        for (int s = 0; s < S; s++) {
            const SpeciesConst sc = p.sp[s];
            double* U1 = p.u + idx_u(p, 0, col, s, 0); double* U2 = p.u + idx_u(p, 1, col, s, 0); double* U3 = p.u + idx_u(p, 2, col, s, 0);
            const double a1 = U1[j], a2 = U2[j], a3 = U3[j];
            const double rin = frcp(a1), v = a2 * rin;
            const double pr = sc.gm1 * (a3 - 0.5 * a2 * v);
            const double e_old = pr * rin / sc.gm1;
            const double cs_old = p.cssave ? p.cssave[idx_s(p, col, s, j)] : fsqrt((sc.gm1 + 1) * sc.gm1 * e_old);
            double T = TT[s * N2 + j];
            if (j <= N) {
                const double kappa = p.kappa_avg[s * KA + Bin + 0];
                const double eta2 = 4. * C_PI * ddt * kappa * p.no_rad_trans / sc.cv;
                double tt = E1[s * N2 + j] * DEN[s * N2 + j];
                if (p.couple_J_into_T) tt += eta2 * DEN[s * N2 + j] * JJ[j];
                if (tt < p.temperature_floor) tt = p.temperature_floor;
                if (Jswitch == 0) T = tt;
            }
            const double e = T * sc.cv;
            const double pn = e * a1 * sc.gm1;
            U2[j] = a1 * v; U3[j] = pn / sc.gm1 + 0.5 * (a1 * v) * v;
            p.dS[idx_s(p, col, s, j)] = DS[s * N2 + j];
            tail_cell(p, col, s, j, v, cs_old, e, T, acc);
        }
    }
    negJ = block_or(negJ);
    tail_commit(p, col, acc, red, negJ != 0, negJval);
}

template <int CPT>
__global__ void __launch_bounds__(256, 2) k_rad2(DevParams p) {
    AIO_DYN_SMEM(double, sm);
    const int col = blockIdx.x;
    const ColState& c = p.cs[col];
    if (!c.in_step) return;
    rad_simple_phase_v2<CPT>(p, col, sm, c.dt, c.t);
}

// ---- single-species radiation phase ("v3"): own-cell quantities live in registers, 7 shared arrays (3 CTAs per SM) ----------------
// Same arithmetic as rad_simple_phase_v2 for S == 1; thread t owns cells t, t+NT, ... in every loop, so rho, T, dS never go
// through shared memory, and eta1 / 1/denom / T are parked in the shared arrays that are dead while the J system is solved.
template <int CPT, int NTC = 0>
__device__ __forceinline__ void rad1_core_v3(const DevParams& p, const int col, double* sm, const double ddt, const double tnow,
                                             double (&u1)[CPT], double (&u2)[CPT], double (&u3)[CPT], TailAcc& acc, int& negJ_out,
                                             double& negJval_out) {
    const int tid = threadIdx.x; const int NT = NTC > 0 ? NTC : (int)blockDim.x;
    const int N = p.N, N2 = p.N2, Bin = p.Bin;
    const int NS = NT * CPT;           // padded array length: every thread can touch its 16-byte aligned chunk of CPT entries
    double* TO = sm;                   // total_opacity (out band 0); parks eta1 during the solve
    double* JJ = TO + NS;              // J; parks T during the solve
    double* TAU = JJ + NS;             // scan workspace, then the interface diffusion coefficients; parks 1/denom during the solve
    double* A = TAU + NS; double* B = A + NS; double* C = B + NS; double* D = C + NS;
    double* red = D + NS;              // 32 + 4*NT
    const int KA = Bin + 2 * p.Bout;
    const SpeciesConst sc = p.sp[0];
    const double gm1 = sc.gm1, igm1 = 1.0 / sc.gm1, icv = 1.0 / sc.cv;
    double* __restrict__ Jg = p.J + idx_J(p, col, 0, 0);
    double* __restrict__ EI = p.eint + idx_s(p, col, 0, 0);
    double rho[CPT], T[CPT], dS[CPT];
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        const int j = tid + k * NT;
        rho[k] = 1; T[k] = 1; dS[k] = 0;
        if (j < N2) {
            const double a1 = u1[k], a2 = u2[k], a3 = u3[k];
            const double rin = frcp(a1), v = a2 * rin;
            rho[k] = a1;
            T[k] = gm1 * (a3 - 0.5 * a2 * v) * rin * igm1 * icv;
            JJ[j] = Jg[j];
        }
    }
    __syncthreads();
    // ---- stellar heating (radiation.cpp:43-348) -------------------------------------------------------------------------------------
    for (int b = 0; b < Bin; b++) {
        double F = p.solar_heating_final[(size_t)col * Bin + b];
        if (tnow < p.radiation_rampup_time)
            F = F * (p.init_radiation_factor + (1. - p.init_radiation_factor) * tnow / p.radiation_rampup_time);
        const double top_prev = p.dSband_top[(size_t)col * Bin + b];
        const double kap = p.kappa_avg[b], hsw = p.highenergy_switch[b];
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int j = tid + k * NT;
            if (j >= 1 && j <= N) TAU[N - j] = (kap * rho[k]) * __ldg(p.dx + j);
        }
        __syncthreads();
        block_inclusive_scan_chunked<CPT, NTC>(TAU, N, red);
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int j = tid + k * NT;
            if (j >= 1 && j <= N + 1) {
                const double dxj = __ldg(p.dx + j);
                const double cell_tau = (kap * rho[k]) * dxj;
                double dSb;
                if (j < N + 1) {
                    const double rad_j = TAU[N - j];
                    const double rad_jp = (j == N) ? 0. : TAU[N - j - 1];
                    const double Sb = F * exp(-rad_jp);
                    dSb = 0.25 * Sb * frcp(dxj) * (1. - p.bond_albedo);
                    const double dtau_tot = -(rad_jp - rad_j);
                    // -expm1(-x) = 1 - exp(-x): above the reference's own series threshold (1e-3) the cancellation costs < 1e-13
                    if (dtau_tot > 1e-3) dSb *= (1.0 - exp(-dtau_tot));
                    else dSb *= (-fastexpm1_2(-dtau_tot));
                    dSb = smax(dSb, 0.);
                } else {
                    dSb = top_prev;
                }
                if (j == N) p.dSband_top[(size_t)col * Bin + b] = dSb;
                const double frac = hsw * kap * rho[k] * dxj * frcp(cell_tau);
                double v = dS[k] + hsw * dSb * frac;
                if (v < 1e-50) v = 0.;
                dS[k] = v;
            }
        }
        __syncthreads();
    }
    // ---- thermal band: total opacity, interface diffusion coefficients ------------------------------------------------------------------
    {
        const double kR = p.kappa_avg[Bin + p.Bout + 0];
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int j = tid + k * NT;
            if (j < N2) TO[j] = (j >= 1) ? fmax(kR * rho[k], 0.) : 0.;
        }
    }
    __syncthreads();
    double* DF = TAU;
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        const int jj = tid + k * NT;
        if (jj < N2) {
            double Dv = 0;
            if (jj >= 1 && jj <= N) {
                const double dxc = (__ldg(p.x_i12 + jj + 1) - __ldg(p.x_i12 + jj));
                const double t0 = TO[jj], t1 = TO[jj + 1];
                double rhokr = smax(2. * (t0 * t1) * frcp(t0 + t1), 4. / 3. * frcp(dxc));
                rhokr = smin(0.5 * (t0 + t1), rhokr);
                const double tau_inv = frcp(dxc * rhokr);
                const double dJ = fabs(JJ[jj + 1] - JJ[jj]) * frcp(JJ[jj] + 1e-300);
                const double R = p.xi_rad * tau_inv * dJ;
                const double fl = (R <= 2) ? 2 * frcp(3 + fsqrt(9 + 10 * R * R)) : 10 * frcp(10 * R + 9 + fsqrt(81 + 180 * R));
                Dv = tau_inv * __ldg(p.surf + jj) * fl;
            }
            DF[jj] = Dv;
        }
    }
    __syncthreads();
    // ---- FLD-simple assembly (radiation_simple.cpp:44-233) ----------------------------------------------------------------------------------
    const double icdt = frcp(C_CLIGHT * ddt);
    const double kP = p.kappa_avg[Bin + 0];
    const double eta2 = 4. * C_PI * ddt * kP * p.no_rad_trans * icv;
    double e1[CPT], idn[CPT];
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        const int j = tid + k * NT;
        e1[k] = 0; idn[k] = 1;
        if (j < N2) {
            double l = 0, d = 0, u = 0, r = 0;
            if (j >= 2 && j <= N + 1) { const double Dm = DF[j - 1]; d = Dm; l = -Dm; }
            if (j <= N) {
                const double tv = __ldg(p.vol + j) * icdt;
                d += tv;
                r += tv * JJ[j];
            }
            if (j >= 1 && j <= N) { const double Dj = DF[j]; u = -Dj; d += Dj; }
            if (j < p.ng) { l = 0; u = -d; r = 0; }
            if (p.closed_radiative_boundaries) {
                if (j >= N2 - p.ng) { l = -l; u = 0; r = 0; }
            } else if (j == N) {
                double f = 1. / p.xi_rad * __ldg(p.x_i12 + N) / __ldg(p.x_i12 + N + 1);
                if (p.geometry == 0) f = 1;
                else if (p.geometry == 2) f *= f;
                l = -f * d; u = 0; r = 0;
            }
            if (j <= N) {
                const double Ts = T[k], Ts3 = Ts * Ts * Ts, rhos = rho[k];
                const double fac = ddt * p.no_rad_trans * kP * icv * C_SIGMA_RAD * Ts3;
                const double in_ = frcp(1. + 16. * fac);
                double eta1 = Ts * (1. + 12. * fac);
                eta1 += ddt * dS[k] * frcp(rhos) * icv;
                e1[k] = eta1; idn[k] = in_;
                // same (cancelling) form as radiation_simple.cpp:210-211: its round-off is part of what the reference computes
                const double sT3 = (C_SIGMA_RAD / C_PI) * Ts3;
                const double exd = p.no_rad_trans * rhos * kP * (1 - 4. * sT3 * eta2 * in_);
                const double exr = p.no_rad_trans * rhos * kP * sT3 * (4 * eta1 * in_ - 3 * Ts);
                d += __ldg(p.vol + j) * exd;
                r += __ldg(p.vol + j) * exr;
            }
            A[j] = l; B[j] = d; C[j] = u; D[j] = r;
            p.dS[idx_s(p, col, 0, j)] = dS[k];
        }
    }
    __syncthreads();
    // park the own-cell values that must survive the solve in arrays that are dead now
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        const int j = tid + k * NT;
        if (j < N2) { TO[j] = e1[k]; TAU[j] = idn[k]; JJ[j] = T[k]; }
    }
    if (p.tridiag_serial) {
        __syncthreads();
        if (tid == 0) {
            for (int i = 1; i < N2; i++) { const double m = A[i] / B[i - 1]; B[i] -= m * C[i - 1]; D[i] -= m * D[i - 1]; }
            D[N2 - 1] /= B[N2 - 1];
            for (int i = N2 - 2; i >= 0; i--) D[i] = (D[i] - C[i] * D[i + 1]) / B[i];
        }
        __syncthreads();
    } else {
        tridiag_substructured<CPT, NTC>(A, B, C, D, N2, red + 32, true);
    }
    // ---- J, T update, e <- T, p <- e, cons <- prim, step tail ------------------------------------------------------------------------------------
    int jneg = 0;
    double Jn[CPT];
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        const int j = tid + k * NT;
        Jn[k] = 0;
        if (j < N2) {
            Jn[k] = p.solve_for_j ? D[j] : Jg[j];
            if (Jn[k] < 0. && p.couple_J_into_T) jneg = 1;
        }
    }
    const int Jswitch = block_or(jneg) || (tnow > 1e15);
    int negJ = 0; double negJval = 0;
    const bool allowJ = p.use_planetary_temperature && (D[N] > 0);
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        const int j = tid + k * NT;
        if (j < N2) {
            double Jv = Jn[k];
            if (p.couple_J_into_T && j <= N - 1 && Jv < 0) {
                if (Jv > -1e-10) Jv *= -1.;
                else if (!(j < N && allowJ)) { negJ = 1; negJval = Jv; }
            }
            Jg[j] = Jv;
            const double a1 = u1[k], a2 = u2[k], a3 = u3[k];
            const double rin = frcp(a1), v = a2 * rin;
            const double pr = gm1 * (a3 - 0.5 * a2 * v);
            const double e_old = pr * rin * igm1;
            const double cs_old = fsqrt((gm1 + 1) * gm1 * e_old);
            double Tn = JJ[j];
            if (j <= N) {
                double tt = TO[j] * TAU[j];
                if (p.couple_J_into_T) tt += eta2 * TAU[j] * Jn[k];
                if (tt < p.temperature_floor) tt = p.temperature_floor;
                if (Jswitch == 0) Tn = tt;
            }
            const double e = Tn * sc.cv;
            const double pn = e * a1 * gm1;
            const double m2 = a1 * v;
            u2[k] = m2; u3[k] = pn * igm1 + 0.5 * m2 * v;
            tail_cell2(p, EI, j, v, cs_old, e, Tn, acc);
        }
    }
    negJ_out = block_or(negJ);
    negJval_out = negJval;
}

// Stand-alone radiation phase of one single-species column: u from / to HBM around rad1_core_v3.
template <int CPT>
__device__ __forceinline__ void rad1_phase_v3(const DevParams& p, const int col, double* sm, const double ddt, const double tnow) {
    const int tid = threadIdx.x, NT = blockDim.x, N2 = p.N2;
    double* __restrict__ U1 = p.u + idx_u(p, 0, col, 0, 0);
    double* __restrict__ U2 = p.u + idx_u(p, 1, col, 0, 0);
    double* __restrict__ U3 = p.u + idx_u(p, 2, col, 0, 0);
    double u1[CPT], u2[CPT], u3[CPT];
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        const int j = tid + k * NT;
        if (j < N2) { u1[k] = U1[j]; u2[k] = U2[j]; u3[k] = U3[j]; } else { u1[k] = 1; u2[k] = 0; u3[k] = 1; }
    }
    TailAcc acc; acc.minstep = 0; acc.maxde = 0; acc.bad = 0; acc.badval = 0;
    int negJ = 0; double negJval = 0;
    rad1_core_v3<CPT>(p, col, sm, ddt, tnow, u1, u2, u3, acc, negJ, negJval);
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        const int j = tid + k * NT;
        if (j < N2) { U2[j] = u2[k]; U3[j] = u3[k]; }
    }
    tail_commit(p, col, acc, sm + 7 * NT * CPT, negJ != 0, negJval);
}


template <int CPT>
__global__ void __launch_bounds__(256, 3) k_rad3(DevParams p) {
    AIO_DYN_SMEM(double, sm);
    const int col = blockIdx.x;
    const ColState& c = p.cs[col];
    if (!c.in_step) return;
    rad1_phase_v3<CPT>(p, col, sm, c.dt, c.t);
}

// ---- fused, persistent step kernel for single-species columns ------------------------------------------------------------------
// One CTA advances its column by up to `p.steps_per_launch` time steps: step head -> both hydro stages -> [radiation] -> step
// tail, with the conserved state in registers and the column's clock (ColState) in shared memory for the whole launch.  HBM sees
// u once per launch instead of four times per step; J, primlast and the static tables are re-read through L2.
template <int CPT, int NTC = 0, int MINB = AIO_STEP1_MINB>
__global__ void __launch_bounds__(256, MINB) k_step1(DevParams p) {
    AIO_DYN_SMEM(double, sm);
    __shared__ ColState s_cs;
    __shared__ double s_p2v;
    const int col = blockIdx.x, tid = threadIdx.x, N2 = p.N2; const int NT = NTC > 0 ? NTC : (int)blockDim.x;
    if (tid == 0) s_cs = p.cs[col];
    const SpeciesConst sc = p.sp[0];
    const double gm1 = sc.gm1, igm1 = 1.0 / sc.gm1, icv = 1.0 / sc.cv;
    double* __restrict__ U1 = p.u + idx_u(p, 0, col, 0, 0);
    double* __restrict__ U2 = p.u + idx_u(p, 1, col, 0, 0);
    double* __restrict__ U3 = p.u + idx_u(p, 2, col, 0, 0);
    double u1[CPT], u2[CPT], u3[CPT];
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        const int j = tid + k * NT;
        if (j < N2) { u1[k] = U1[j]; u2[k] = U2[j]; u3[k] = U3[j]; } else { u1[k] = 1; u2[k] = 0; u3[k] = 1; }
    }
    double prim2_speed = p.prim2v[col];
    const int LDH = NTC > 0 ? NTC * CPT : N2;
    double* red = sm + 9 * LDH;
    bool touched = false;
    for (int it = 0; it < p.steps_per_launch; it++) {
        const StepInfo si = step_head_on(p, s_cs);
        if (!si.active) break;
        touched = true;
        if (si.snap) hydro_snapshot_v2(p, col);
        hydro_stages_v2<CPT, NTC>(p, col, 0, sc, sm, si.dt, si.t_old, u1, u2, u3, prim2_speed);
        TailAcc acc; acc.minstep = 0; acc.maxde = 0; acc.bad = 0; acc.badval = 0;
        int negJ = 0; double negJval = 0;
        __syncthreads();
        if (p.use_rad_fluxes == 2) {
            rad1_core_v3<CPT, NTC>(p, col, sm, si.dt, si.t_old + si.dt, u1, u2, u3, acc, negJ, negJval);
        } else {
            double* EI = p.eint + idx_s(p, col, 0, 0);
#pragma unroll
            for (int k = 0; k < CPT; k++) {
                const int j = tid + k * NT;
                if (j < N2) {
                    const double rin = frcp(u1[k]);
                    const double v = u2[k] * rin;
                    const double pr = gm1 * (u3[k] - 0.5 * u2[k] * v);
                    const double e = pr * rin * igm1;
                    tail_cell2(p, EI, j, v, fsqrt((gm1 + 1) * gm1 * e), e, e * icv, acc);
                }
            }
        }
        // prim[2].speed as the next step's 'fixed' boundary sees it (init_and_bounds.cpp:1856)
#pragma unroll
        for (int k = 0; k < CPT; k++) if (tid + k * NT == 2) s_p2v = u2[k] * frcp(u1[k]);
        __syncthreads();
        tail_commit_on(s_cs, acc, red, negJ != 0, negJval);
        __syncthreads();
        if (N2 > 2) prim2_speed = s_p2v;
    }
    if (touched) {
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int j = tid + k * NT;
            if (j < N2) { U1[j] = u1[k]; U2[j] = u2[k]; U3[j] = u3[k]; if (j == 2) p.prim2v[col] = prim2_speed; }
        }
    }
    __syncthreads();
    if (tid == 0) p.cs[col] = s_cs;
}

}  // namespace aio
