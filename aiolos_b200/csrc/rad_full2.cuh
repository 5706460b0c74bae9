// Full FLD solver (PARI_USE_RADIATION 1), second generation: the coupled J_b / T_s block-tridiagonal system of
// update_fluxes_FLD (reference radiation.cpp:353-840) solved by block Thomas (block_tridiag_solve.h:57-130) with the
// dense n x n LU of every cell held in REGISTERS.  Four launches per step:
//
//   k_full_heat   warp per column, lane = incoming band: reset_dS / update_dS in the reference's sequential top-down order
//                 (radiation.cpp:15-348)
//   k_full_asm    CTA per column, 8 warps over contiguous chunks of cells, lane = outgoing band: every coefficient of the
//                 block rows (radiation.cpp:407-615) -- flux limiter, Planck fractions, matter coupling -- none of which
//                 depends on the solve; written as a compact per-cell coefficient stream (diagonal, l, u, rhs, J-T / T-J
//                 couplings: NC doubles instead of 3 n^2)
//   k_full_lu     ONE WARP per column, the sequential part: per cell  S_i = d_i - l_i X_{i-1},  PartialPivLU(S_i) with the
//                 augmented columns [u_i | rhs_i] riding along, back substitution  X_i = S_i^-1 u_i,  A_i = S_i^-1 (...),
//                 then the back sweep  x_i = A_i - X_i x_{i+1}.
//   k_full_post   CTA per column, thread per cell: negative-J scan, T clamp, e <- T, p <- e, cons <- prim, step tail.
//
// k_full_lu register tiling: the augmented matrix [S | u | rhs] (R x 2R, R = 4*RI >= n) is spread over the warp as a
// 4 (row groups) x 8 (column groups) grid of lanes; lane (rg, cg) owns rows rg + 4 i and columns cg + 8 j (i, j < RI),
// i.e. RI x RI doubles in registers (36 for configs[2], n = 23).  One elimination step k is a rank-1 update
//   a[i][j] -= l[i] * pivotrow[j]
// that costs each lane RI + RI shared-memory words (two broadcast buffers) for RI*RI FMAs: for configs[2] one 128-bit
// shared load per 6 DFMA, against 2 LDS + 1 STS per DFMA in the first-generation kernel (rad_full.cuh).  The pivot
// search runs lane = row on the published column with two REDUX.MAX (the IEEE bit pattern of |x| orders like an
// unsigned integer) and a ballot for Eigen's first-maximum rule.  Finished pivot rows go straight to shared memory (U
// transposed, forward-eliminated right-hand sides by row), where the back substitution -- lane = right-hand-side column,
// the column in registers -- reads them; rows are never swapped back.  Every matrix element sees exactly the sequence of
// roundings of Eigen's PartialPivLU / triangular solves (per element the updates are applied in increasing k, each as one
// fused a - l*p, divisions are true divisions, a zero pivot column is not scaled, back substitution skips exactly-zero
// unknowns: the all-zero T rows of the last ghost cell rely on that).
//
// Data movement: the coefficient stream and, in the back sweep, the X_i blocks are fetched two cells ahead with
// cp.async.bulk (TMA 1-D bulk copy, completion on an mbarrier) so that the sequential sweep never waits on HBM.
#pragma once
#include "rad_full.cuh"
#include "rad_full2.h"

namespace aio {

// ---- bulk-copy (TMA 1-D) + mbarrier helpers --------------------------------------------------------------------------------------
#ifndef AIO_EMU_BUILD
__device__ __forceinline__ unsigned smem_u32(const void* ptr) { return (unsigned)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(double* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// one lane: arm the barrier with the byte count and start the copy global -> shared
__device__ __forceinline__ void bulk_g2s(double* dst, const double* src, unsigned bytes, double* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(double* bar, unsigned parity) {
    unsigned done = 0;
    int spins = 0;
    const unsigned a = smem_u32(bar);
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (!done && ++spins > (1 << 24)) __trap();          // a lost copy must not hang the device
    } while (!done);
}
__device__ __forceinline__ unsigned warp_umax(unsigned v) { return __reduce_max_sync(0xffffffffu, v); }
__device__ __forceinline__ unsigned hi32(double x) { return (unsigned)__double2hiint(x); }
// keeps the compiler from hoisting a select across a division (it would divide the raw operands and select afterwards)
__device__ __forceinline__ void opaque(double& x) { asm volatile("" : "+d"(x)); }
__device__ __forceinline__ unsigned lo32(double x) { return (unsigned)__double2loint(x); }
#else
inline void mbar_init(double*, int) {}
inline void fence_mbar_init() {}
inline void fence_proxy_async() {}
inline void bulk_g2s(double* dst, const double* src, unsigned bytes, double*) { memcpy(dst, src, bytes); }
inline void mbar_wait(double*, unsigned) {}
inline unsigned warp_umax(unsigned v) {
    for (int o = 16; o > 0; o >>= 1) { const int w = __shfl_xor_sync(0xffffffffu, (int)v, o); if ((unsigned)w > v) v = (unsigned)w; }
    return v;
}
inline unsigned hi32(double x) { unsigned long long b; memcpy(&b, &x, 8); return (unsigned)(b >> 32); }
inline void opaque(double&) {}
inline unsigned lo32(double x) { unsigned long long b; memcpy(&b, &x, 8); return (unsigned)b; }
#endif

// ---- stellar heating: lane = incoming band, top-down exactly as radiation.cpp:51-63,154-348 -----------------------------------
__global__ void __launch_bounds__(128) k_full_heat(DevParams p) {
    __shared__ double contrib_all[4][MAXS * MAXB];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int col = blockIdx.x * (blockDim.x >> 5) + wib;
    if (col >= p.ncol) return;
    const ColState& cst = p.cs[col];
    if (!cst.in_step) return;
    const double tnow = cst.t;
    const int N = p.N, S = p.S, Bin = p.Bin;
    double* contrib = contrib_all[wib];
    const int b = lane;
    double F = 0, rad_above = 0, top_prev = 0;
    if (b < Bin) {
        F = p.solar_heating_final[(size_t)col * Bin + b];
        if (tnow < p.radiation_rampup_time)
            F = F * (p.init_radiation_factor + (1. - p.init_radiation_factor) * tnow / p.radiation_rampup_time);
        top_prev = p.dSband_top[(size_t)col * Bin + b];
    }
    for (int j = N + 1; j > 0; j--) {
        const double dxj = __ldg(p.dx + j);
        if (b < Bin) {
            double tot = 0;
            for (int s = 0; s < S; s++) tot += kapv(p, col, s, b, j) * p.u[idx_u(p, 0, col, s, j)];
            const double cell_tau = tot * dxj;
            const double rad_j = (j == N + 1) ? 0. : rad_above + cell_tau;
            double dSb;
            if (j < N + 1) {
                const double Sb = F * exp(-rad_above);
                if (j == p.ng - 1) p.ssurf[(size_t)col * Bin + b] = Sb;
                dSb = 0.25 * Sb / dxj * (1. - p.bond_albedo);
                const double dtau_tot = -(rad_above - rad_j);
                double dS_he = dSb;
                if (dtau_tot > 1e-3) dSb *= (-expm1(-dtau_tot));
                else dSb *= (-fastexpm1_2(-dtau_tot));
                dS_he *= -fastexpm1_2(-0.);
                dSb -= dS_he;
                dSb = smax(dSb, 0.);
            } else {
                dSb = top_prev;
            }
            if (j == N) p.dSband_top[(size_t)col * Bin + b] = dSb;
            for (int s = 0; s < S; s++) {
                const double hsw = p.highenergy_switch[s * Bin + b];
                const double frac = hsw * kapv(p, col, s, b, j) * p.u[idx_u(p, 0, col, s, j)] * dxj / cell_tau;
                contrib[s * Bin + b] = hsw * dSb * frac;
            }
            rad_above = rad_j;
        }
        __syncwarp();
        if (lane < S) {
            double acc = 0.;                              // reset_dS
            for (int bb = 0; bb < Bin; bb++) { acc += contrib[lane * Bin + bb]; if (acc < 1e-50) acc = 0.; }
            p.dS[idx_s(p, col, lane, j)] = acc;
        }
        __syncwarp();
    }
    if (lane < S) p.dS[idx_s(p, col, lane, 0)] = 0.;
}

// ---- stellar heating, parallel over cells (k_full_heat above walks the column with one lane per band: 273 dependent iterations of
// exp / expm1, 0.66 ms at 4096 columns of configs[2]).  Only the optical-depth sum is sequential -- and must stay so, its
// rounding order is the reference's (radiation.cpp:154-169) -- so it is the only thing done sequentially here: phase A, thread =
// cell, cell optical depths into shared memory; phase B, thread = band, the running sum top-down (one dependent add per cell);
// phase C, thread = cell, attenuation, heating and the per-species sums in band order.  Same expressions as k_full_heat; CTA =
// column.  Needs Bin * (N2 + 1) doubles of shared memory; columns too long for that keep k_full_heat.
template <int SMAX>
__global__ void __launch_bounds__(512) k_full_heat2(DevParams p) {
    AIO_DYN_SMEM(double, R);                       // [Bin][LDR] cell optical depth, then optical depth above the cell
    const int col = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const ColState& cst = p.cs[col];
    if (!cst.in_step) return;
    const double tnow = cst.t;
    const int N = p.N, N2 = p.N2, S = p.S, Bin = p.Bin;
    const int LDR = (N2 + 1) | 1;
    double* top_prev = R + (size_t)Bin * LDR;      // [Bin] dS_band(N, b) of the previous step
    for (int b = tid; b < Bin; b += NT) top_prev[b] = p.dSband_top[(size_t)col * Bin + b];
    for (int j = tid + 1; j <= N + 1; j += NT) {
        const double dxj = __ldg(p.dx + j);
        for (int b = 0; b < Bin; b++) {
            double tot = 0;
            for (int s = 0; s < S; s++) tot += kapv(p, col, s, b, j) * p.u[idx_u(p, 0, col, s, j)];
            R[b * LDR + j] = tot * dxj;
        }
    }
    __syncthreads();
    for (int b = tid; b < Bin; b += NT) {
        double rad_above = 0;
        for (int j = N + 1; j > 0; j--) {
            const double cell_tau = R[b * LDR + j];
            R[b * LDR + j] = rad_above;
            rad_above = (j == N + 1) ? 0. : rad_above + cell_tau;
        }
    }
    __syncthreads();
    for (int j = tid + 1; j <= N + 1; j += NT) {
        const double dxj = __ldg(p.dx + j);
        double acc[SMAX], rho[SMAX];
#pragma unroll
        for (int s = 0; s < SMAX; s++) { acc[s] = 0.; rho[s] = (s < S) ? p.u[idx_u(p, 0, col, s, j)] : 0.; }
        for (int b = 0; b < Bin; b++) {
            double F = p.solar_heating_final[(size_t)col * Bin + b];
            if (tnow < p.radiation_rampup_time)
                F = F * (p.init_radiation_factor + (1. - p.init_radiation_factor) * tnow / p.radiation_rampup_time);
            const double rad_above = R[b * LDR + j];
            double tot = 0;
#pragma unroll
            for (int s = 0; s < SMAX; s++) if (s < S) tot += kapv(p, col, s, b, j) * rho[s];
            const double cell_tau = tot * dxj;
            const double rad_j = (j == N + 1) ? 0. : rad_above + cell_tau;
            double dSb;
            if (j < N + 1) {
                const double Sb = F * exp(-rad_above);
                if (j == p.ng - 1) p.ssurf[(size_t)col * Bin + b] = Sb;
                dSb = 0.25 * Sb / dxj * (1. - p.bond_albedo);
                const double dtau_tot = -(rad_above - rad_j);
                double dS_he = dSb;
                if (dtau_tot > 1e-3) dSb *= (-expm1(-dtau_tot));
                else dSb *= (-fastexpm1_2(-dtau_tot));
                dS_he *= -fastexpm1_2(-0.);
                dSb -= dS_he;
                dSb = smax(dSb, 0.);
            } else {
                dSb = top_prev[b];
            }
            if (j == N) p.dSband_top[(size_t)col * Bin + b] = dSb;
#pragma unroll
            for (int s = 0; s < SMAX; s++)
                if (s < S) {
                    const double hsw = p.highenergy_switch[s * Bin + b];
                    const double frac = hsw * kapv(p, col, s, b, j) * rho[s] * dxj / cell_tau;
                    acc[s] += hsw * dSb * frac;               // reset_dS + the band sum, in band order
                    if (acc[s] < 1e-50) acc[s] = 0.;
                }
        }
#pragma unroll
        for (int s = 0; s < SMAX; s++) if (s < S) p.dS[idx_s(p, col, s, j)] = acc[s];
    }
    for (int s = tid; s < S; s += NT) p.dS[idx_s(p, col, s, 0)] = 0.;
}

// one record of the coefficient stream: `len` doubles (a multiple of 4) as 16-byte stores, i.e. whole 32-byte sectors
template <int NMAX>
__device__ __forceinline__ void store_record(double* dst, const double (&rec)[NMAX], int len) {
#pragma unroll
    for (int k = 0; k < NMAX; k += 2)
        if (k < len) *reinterpret_cast<double2*>(dst + k) = make_double2(rec[k], rec[k + 1]);
}

// ---- block-row coefficients of every cell (radiation.cpp:407-615), one THREAD per cell -- flux limiter, Planck fractions, matter
// coupling: nothing here depends on the solve.  (The first version ran one lane per band: 22 of 32 lanes busy, the T rows summed
// by a single lane while its warp waited, 1.7 k warp instructions per cell and 2.9 ms at 4096 columns of configs[2], ncu r2l.)
// With lane = cell the state loads are coalesced, every lane is busy, the band loop runs inside the thread, so the T-row sums
// accumulate in registers in the reference's band order, and no shared memory or warp barrier is needed.  The coefficients
// leave as band / species records (Full2Dims): whole 32-byte sectors per store, no partial-sector fills in L2.
template <int SMAX>
__global__ void __launch_bounds__(128) k_full_asm2(DevParams p) {
    const int N = p.N, N2 = p.N2, S = p.S, Bin = p.Bin, Bout = p.Bout;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;        // (column, cell) flattened: no idle tail lanes
    if (gid >= (long long)p.ncol * N2) return;
    const int col = (int)(gid / N2), i = (int)(gid - (long long)col * N2);
    const ColState& cst = p.cs[col];
    if (!cst.in_step) return;
    const double dt = cst.dt, tnow = cst.t;
    const Full2Dims fd = full2_dims(Bout, S);
    const int n = fd.n;
    double* cf = p.fcoef + ((size_t)col * N2 + i) * fd.NC;
    const double log_spacing = log(p.lT_spacing);
    // densities of the cell and its neighbours (total opacity), primitive temperature of the cell
    double um[SMAX], u0[SMAX], up[SMAX], T_s[SMAX], dT[SMAX], rT[SMAX];
#pragma unroll
    for (int s = 0; s < SMAX; s++) {
        um[s] = u0[s] = up[s] = T_s[s] = dT[s] = rT[s] = 0.;
        if (s < S) {
            if (i >= 2) um[s] = p.u[idx_u(p, 0, col, s, i - 1)];                     // total_opacity(j) is 0 outside 1..N+1
            u0[s] = p.u[idx_u(p, 0, col, s, i)];
            if (i + 1 <= N + 1) up[s] = p.u[idx_u(p, 0, col, s, i + 1)];
            if (i <= N) {
                const SpeciesConst sc = p.sp[s];
                const double u1 = p.u[idx_u(p, 0, col, s, i)], u2 = p.u[idx_u(p, 1, col, s, i)], u3 = p.u[idx_u(p, 2, col, s, i)];
                Prim3 q = cons_to_prim(u1, u2, u3, sc.gm1);
                T_s[s] = p.Tsave ? p.Tsave[idx_s(p, col, s, i)] : eint_of(q.p, q.rho, sc.gm1) / sc.cv;
                dT[s] = 1 / dt;
                rT[s] = T_s[s] / dt;
                rT[s] += (p.dS[idx_s(p, col, s, i)] + 0.) / u1 / sc.cv;
            }
        }
    }
    auto Dcoef = [&](int ii, double Jc, double Jn, double toi, double toip) -> double {
        const double dxc = (__ldg(p.x_i12 + ii + 1) - __ldg(p.x_i12 + ii));
        const double rhokr = sqrt(toi * toip);
        const double tau_inv = 1. / (dxc * rhokr);
        const double R = p.xi_rad * tau_inv * fabs(Jn - Jc) / (Jc + 1e-300);
        return 1. * tau_inv * 1. * __ldg(p.surf + ii) * flux_limiter(R);
    };
    const bool coupled = (i <= N && !(i < p.ng || i >= N + 2 - p.ng));
    constexpr int RECMAX = (4 + 2 * SMAX + 3) & ~3, TRECMAX = (2 + SMAX + 3) & ~3;
    double rec[RECMAX];
#pragma unroll
    for (int k = 0; k < RECMAX; k++) rec[k] = 0.;
    for (int b = 0; b < Bout; b++) {
        // total (Rosseland) opacity of the cell and its neighbours: the largest species term
        double to_m = 0, to_i = 0, to_ip = 0;
#pragma unroll
        for (int s = 0; s < SMAX; s++)
            if (s < S) {
                if (i - 1 >= 1) to_m = fmax(kapv(p, col, s, Bin + Bout + b, i - 1) * um[s], to_m);
                if (i >= 1) to_i = fmax(kapv(p, col, s, Bin + Bout + b, i) * u0[s], to_i);      // i <= N+1 always
                if (i + 1 <= N + 1) to_ip = fmax(kapv(p, col, s, Bin + Bout + b, i + 1) * up[s], to_ip);
            }
        const double Jcur = p.J[idx_J(p, col, b, i)];
        const double Jnext = (i + 1 < N2) ? p.J[idx_J(p, col, b, i + 1)] : 0.;
        double D_prev = 0;
        if (i - 1 >= 1 && i - 1 <= N) D_prev = Dcoef(i - 1, p.J[idx_J(p, col, b, i - 1)], Jcur, to_m, to_i);
        double l_rr = 0, u_rr = 0, rhs = 0, d = 0;
        if (i >= 2 && i <= N + 1) { d = D_prev; l_rr = -D_prev; }
        if (i <= N) {
            const double tv = __ldg(p.vol + i) / (C_CLIGHT * dt);
            d += tv;
            rhs += tv * Jcur;
        }
        if (i >= 1 && i <= N) {
            const double D_i = Dcoef(i, Jcur, Jnext, to_i, to_ip);
            u_rr = -D_i; d += D_i;
        }
        if (i < p.ng) { l_rr = 0; u_rr = -d; rhs = 0; }
        if (i == p.ng - 1 && p.use_planetary_temperature) {      // surface cell radiates at T_planet (radiation.cpp:451-472)
            const double Tp = cst.T_planet;
            double Ssrc = C_SIGMA_RAD * pow(Tp, 4.) * planck_integral3(p, p.l_i_out[b], p.l_i_out[b + 1], fabs(Tp), log_spacing);
            if (Tp < 0) Ssrc *= -1;
            const double dxc = (__ldg(p.x_i12 + i + 1) - __ldg(p.x_i12 + i));
            double rhokr = smax(2. * (to_i * to_ip) / (to_i + to_ip), 4. / 3. / dxc);
            rhokr = smin(0.5 * (to_i + to_ip), rhokr);
            const double tau_inv = 1 / (dxc * rhokr);
            const double R = p.xi_rad * tau_inv * fabs(Jnext - Jcur) / (Jnext + 1e-300);
            const double Dc = flux_limiter(R) * tau_inv;
            p.dcore[(size_t)col * Bout + b] = Dc;
            const double Chi = 0.25;
            const double sj = __ldg(p.surf + i);
            l_rr = 0; d = 0.5 * (Dc + Chi) * sj; u_rr = -0.5 * (Dc - Chi) * sj; rhs = sj * Ssrc / (4 * C_PI);
        }
        if (p.closed_radiative_boundaries) {
            if (i >= N2 - p.ng) { l_rr = -d; u_rr = 0; rhs = 0; }
        } else if (i == N + 1) {
            double f = __ldg(p.x_i12 + N) / __ldg(p.x_i12 + N + 1);
            if (p.geometry == 0) f = 1;
            else if (p.geometry == 2) f *= f;
            l_rr = -f * d; u_rr = 0; rhs = 0;
        }
        // matter coupling: fac[s][b] = no_rad_trans * kappa_P * sigma T^3 / pi * planck fraction
#pragma unroll
        for (int s = 0; s < SMAX; s++)
            if (s < S) {
                double jt = 0., tj = 0.;
                if (i <= N) {
                    const double cv = p.sp[s].cv;
                    const double Ts = T_s[s];
                    const double kP = kapv(p, col, s, Bin + b, i);
                    const double fac = 1. * p.no_rad_trans * kP * C_SIGMA_RAD * Ts * Ts * Ts / C_PI *
                                       planck_integral3(p, p.l_i_out[b], p.l_i_out[b + 1], Ts, log_spacing);
                    dT[s] += 16 * C_PI * fac / cv;
                    rT[s] += 12 * C_PI * fac * Ts / cv;
                    tj = -4 * C_PI * kP * p.no_rad_trans / cv;
                    if (coupled) {
                        const double rho = u0[s];
                        const double vj = __ldg(p.vol + i);
                        d += vj * rho * kP * p.no_rad_trans;
                        jt = -4 * vj * rho * fac;
                        rhs -= 3 * vj * rho * fac * Ts;
                    }
                }
                rec[4 + s] = jt;
                rec[4 + S + s] = tj;
            }
        rec[0] = d; rec[1] = l_rr; rec[2] = u_rr; rec[3] = rhs;
        store_record<RECMAX>(cf + b * fd.W, rec, fd.W);
    }
    // T rows: collisional T_i - T_j coupling (radiation.cpp:589-615, source.cpp:656-674), then the row itself
#pragma unroll
    for (int s = 0; s < SMAX; s++)
        if (s < S) {
            double trec[TRECMAX];
#pragma unroll
            for (int k = 0; k < TRECMAX; k++) trec[k] = 0.;
            double d = 0., rhs = 0.;
            if (i <= N) {
                d = dT[s]; rhs = rT[s];
                if (p.use_collisional_heating && S > 1) {
                    const SpeciesConst sc = p.sp[s];
                    const double u1 = u0[s], Ts = T_s[s];
                    double tot0 = 0;
                    if (i >= 1) for (int ss = 0; ss < S; ss++) tot0 = fmax(kapv(p, col, ss, Bin + Bout + 0, i) * p.u[idx_u(p, 0, col, ss, i)], tot0);
                    const double tau = tot0 * (__ldg(p.x_i12 + i + 1) - __ldg(p.x_i12 + i));
                    const double damp = (tau < 1e3) ? 1. : 1.e-4;
                    double diag_sum = 0, self = 0;
                    for (int sj = 0; sj < S; sj++) {          // rolled: density and temperature of species sj are re-derived
                        const SpeciesConst cj = p.sp[sj];
                        const double w1 = p.u[idx_u(p, 0, col, sj, i)];
                        const double v2 = p.u[idx_u(p, 1, col, sj, i)], v3 = p.u[idx_u(p, 2, col, sj, i)];
                        const Prim3 qj = cons_to_prim(w1, v2, v3, cj.gm1);
                        const double Tj = p.Tsave ? p.Tsave[idx_s(p, col, sj, i)] : eint_of(qj.p, qj.rho, cj.gm1) / cj.cv;
                        double fc = p.fric_mask[s * S + sj] * alpha_pair(p, s, sj, S, sc, cj, u1, w1, w1 / cj.mass, Ts, Tj, tnow);
                        fc *= 3 * C_KB / (sc.cv * (sc.mass + cj.mass));
                        diag_sum += fc;
                        const double off = 0. - fc * damp;
#pragma unroll
                        for (int k = 0; k < SMAX; k++) if (k == sj && sj != s) trec[2 + k] = off;
                        if (sj == s) self = fc;
                    }
                    d -= (self - diag_sum) * damp;
                }
            }
            trec[0] = d; trec[1] = rhs;
            store_record<TRECMAX>(cf + fd.o_T + s * fd.V, trec, fd.V);
        }
}

// ---- the sequential sweep: one warp per column, LU in registers ---------------------------------------------------------------------
template <int RI>
struct LuCtx {
    static constexpr int R = 4 * RI, LDU = R + 2, LDY = R;      // same numbers as full2_dims(); compile-time here
    double* colb; double* UT; double* Y;
    // this lane's windows into the broadcast buffers
    double* my_col;     // colb + rg            : rows rg + 4 i of the published column
    double* my_l;       // lb + rg * RI         : multipliers of this lane's rows
    double* my_l_out;   // lb slot of row `lane` (lane = row during the pivot search)
    double* my_prow;    // prow + cg * RI       : pivot row entries of this lane's columns
    double* my_swp;     // swp + cg * RI
    double* my_UT;      // UT + cg * LDU        : U(k, cg + 8 j) goes to my_UT[8 j LDU + k]
    double* my_Y;       // Y + cg               : Y(k, cg + 8 j') goes to my_Y[k LDY + 8 j']
    int n, rg, cg, lane;
};

// Elimination steps k = 4Q .. 4Q+3: rows of local index < Q and column blocks < Q/2 are finished (dead registers).
//
// The step is latency-bound (ncu r2e: one warp issues an instruction every ~8 cycles, the chain  publish column -> pivot
// search -> division -> multipliers -> update -> next column  is strictly serial), so the loop is software-pipelined:
//   * the multipliers are computed speculatively with the diagonal entry as pivot while the pivot search runs (the blocks are
//     diagonally dominant almost everywhere; a row exchange redoes the division),
//   * of the rank-1 update of step k only the block column that holds column k+1 is applied at once, so that column k+1 can
//     be published; the rest of the update is deferred into step k+1, where it fills the latency of that step's pivot
//     search and division (same basic block: the compiler interleaves the two).
template <int RI, int Q>
__device__ __forceinline__ void lu_phase(double (&a)[RI][RI], const LuCtx<RI>& c) {
    constexpr int JK = Q / 2;          // local column block that holds column k
    constexpr int LB = RI / 2;         // column blocks of the LU part; blocks LB.. are the right-hand sides
    constexpr int LDU = LuCtx<RI>::LDU, LDY = LuCtx<RI>::LDY;
    const int lane = c.lane, rg = c.rg, cg = c.cg;
    if (4 * Q >= c.n) return;
    // publish column 4Q (live rows only)
    if (cg == 4 * (Q & 1)) {
#pragma unroll
        for (int i = Q; i < RI; i++) c.my_col[4 * i] = a[i][JK];
    }
    bool deferred = false;
#pragma unroll 1
    for (int kk = 0; kk < 4; kk++) {
        const int k = 4 * Q + kk;
        if (k >= c.n) break;
        __syncwarp();
        // 1. column k as published; multipliers and rest of the previous update from the broadcast buffers of step k-1
        const double v = c.colb[lane];
        const double vk = c.colb[k];
        const bool live = (lane > k) && (lane < c.n);
        // speculative multipliers l = a(r,k) / a(k,k).  Zero entries (structural zeros of the blocks are common) and dead lanes
        // must not reach the division: a zero numerator sends the whole warp through the IEEE slow path of the FP64 division
        // (ncu r2b: 22 % of all executed instructions).  0 / pv = 0, so they bypass it.
        const bool dspec = live && (v != 0.) && (vk != 0.);
        double num = dspec ? v : 1., den = dspec ? vk : 1.;
        opaque(num); opaque(den);
        const double qspec = num / den;
        double l = dspec ? qspec : (live ? v : 0.);
        // 2. pivot search, lane = row: first row holding the largest |a(r,k)|, r >= k  (PartialPivLU)
        const bool cand = (lane >= k) && (lane < c.n);
        const double av = fabs(v);
        const unsigned hk = cand ? hi32(av) + 1u : 0u;
        const unsigned mh = warp_umax(hk);
        const bool top = cand && (hk == mh);
        const unsigned lk = top ? lo32(av) : 0u;
        const unsigned ml = warp_umax(lk);
        const unsigned votes = __ballot_sync(0xffffffffu, top && (lk == ml));
        const int piv = __ffs(votes) - 1;
        const bool swap = (piv != k);                     // warp-uniform
        // 3. the deferred part of the update of step k-1 (block columns right of the one that holds column k)
        if (deferred) {
            double lvp[RI], prp[RI];
#pragma unroll
            for (int i = Q; i < RI; i++) lvp[i] = c.my_l[i];
#pragma unroll
            for (int j = JK + 1; j < RI; j++) prp[j] = c.my_prow[j];
#pragma unroll
            for (int i = Q; i < RI; i++) {
#pragma unroll
                for (int j = JK + 1; j < RI; j++) a[i][j] = d_fnma(lvp[i], prp[j], a[i][j]);
            }
        }
        // 4. a row exchange invalidates the speculation: position piv now holds the old row k, the pivot is a(piv,k) != 0
        if (swap) {
            const double pv = c.colb[piv];
            const double mine = (lane == piv) ? vk : v;
            const bool divide = live && (mine != 0.);
            double num2 = divide ? mine : 1., den2 = divide ? pv : 1.;
            opaque(num2); opaque(den2);
            const double q2 = num2 / den2;
            l = divide ? q2 : (live ? mine : 0.);
        }
        __syncwarp();                   // every lane has read the buffers of step k-1
        if (lane < 4 * RI) *c.my_l_out = l;
        // 5. the pivot row: broadcast copy for the update; it is row k of U (stored transposed: the back substitution reads
        //    column k of U contiguously) and of the forward-eliminated right-hand sides.  After an exchange the old row k
        //    (lanes rg == kk, local row Q) moves to position piv.
        const int ip = piv >> 2, rgp = piv & 3;
        if (!swap) {                    // the usual case: the pivot row is row k itself, local row Q of the lanes rg == kk
            if (rg == kk) {
#pragma unroll
                for (int j = JK; j < RI; j++) c.my_prow[j] = a[Q][j];
#pragma unroll
                for (int j = JK; j < LB; j++) c.my_UT[8 * j * LDU + k] = a[Q][j];
#pragma unroll
                for (int j = LB; j < RI; j++) c.my_Y[k * LDY + 8 * (j - LB)] = a[Q][j];
            }
            __syncwarp();
        } else {
            if (rg == rgp) {
#pragma unroll
                for (int i = Q; i < RI; i++)
                    if (i == ip) {      // ip is warp-uniform: a branch, the stores read the registers of row i directly
#pragma unroll
                        for (int j = JK; j < RI; j++) c.my_prow[j] = a[i][j];
#pragma unroll
                        for (int j = JK; j < LB; j++) c.my_UT[8 * j * LDU + k] = a[i][j];
#pragma unroll
                        for (int j = LB; j < RI; j++) c.my_Y[k * LDY + 8 * (j - LB)] = a[i][j];
                    }
            }
            if (rg == kk) {
#pragma unroll
                for (int j = JK; j < RI; j++) c.my_swp[j] = a[Q][j];
            }
            __syncwarp();
            if (rg == rgp) {
#pragma unroll
                for (int i = Q; i < RI; i++)
                    if (i == ip) {
#pragma unroll
                        for (int j = JK; j < RI; j++) a[i][j] = c.my_swp[j];
                    }
            }
        }
        // 6. update the block column that holds column k+1 and publish that column; the rest waits for step k+1
        {
            const double pj = c.my_prow[JK];
#pragma unroll
            for (int i = Q; i < RI; i++) a[i][JK] = d_fnma(c.my_l[i], pj, a[i][JK]);
        }
        deferred = true;
        if (kk < 3 && k + 1 < c.n && cg == 4 * (Q & 1) + kk + 1) {
#pragma unroll
            for (int i = Q; i < RI; i++) c.my_col[4 * i] = a[i][JK];
        }
    }
    // the deferred part of the phase's last step
    if (deferred) {
        double lvp[RI], prp[RI];
#pragma unroll
        for (int i = Q; i < RI; i++) lvp[i] = c.my_l[i];
#pragma unroll
        for (int j = JK + 1; j < RI; j++) prp[j] = c.my_prow[j];
#pragma unroll
        for (int i = Q; i < RI; i++) {
#pragma unroll
            for (int j = JK + 1; j < RI; j++) a[i][j] = d_fnma(lvp[i], prp[j], a[i][j]);
        }
    }
    __syncwarp();
}

template <int RI>
__global__ void __launch_bounds__(32, FULL_LU_CTAS_PER_SM) k_full_lu(DevParams p) {
    AIO_DYN_SMEM(double, sm);
    const int lane = threadIdx.x & 31;
    const int col = blockIdx.x;
    const ColState& cst = p.cs[col];
    if (!cst.in_step) return;
    const int N2 = p.N2, S = p.S, Bout = p.Bout;
    const Full2Dims fd = full2_dims_ri(Bout, S, RI);
    const int n = fd.n;
    constexpr int R = 4 * RI, LB = RI / 2;
    LuCtx<RI> c;
    c.rg = lane >> 3; c.cg = lane & 7; c.lane = lane; c.n = n;
    const int rg = c.rg, cg = c.cg;
    c.colb = sm + fd.o_col; c.UT = sm + fd.o_UT; c.Y = sm + fd.o_Y;
    c.my_col = c.colb + rg; c.my_l = sm + fd.o_l + rg * RI; c.my_l_out = sm + fd.o_l + (lane & 3) * RI + ((lane >> 2) % RI);
    c.my_prow = sm + fd.o_prow + cg * RI; c.my_swp = sm + fd.o_swap + cg * RI;
    c.my_UT = c.UT + cg * c.LDU; c.my_Y = c.Y + cg;
    double* coefs = sm + fd.o_coef;
    double* bars = sm + fd.o_bar;                  // [0,1]: coefficient ring, [2,3]: X ring of the back sweep
    const double* cf_g = p.fcoef + (size_t)col * N2 * fd.NC;
    double* Xg = p.Xstore + (size_t)col * N2 * fd.XB;
    const unsigned coef_bytes = (unsigned)(fd.NC * sizeof(double)), x_bytes = (unsigned)(fd.XB * sizeof(double));

    if (lane == 0) { for (int q = 0; q < 4; q++) mbar_init(bars + q, 1); fence_mbar_init(); }
    for (int q = lane; q < R * fd.LDY; q += 32) c.Y[q] = 0.;
    __syncwarp();
    if (lane == 0) {
        bulk_g2s(coefs, cf_g, coef_bytes, bars + 0);
        if (N2 > 1) bulk_g2s(coefs + fd.NC, cf_g + fd.NC, coef_bytes, bars + 1);
    }

    // ---- forward sweep ------------------------------------------------------------------------------------------------------------
    for (int i = 0; i < N2; i++) {
        const int slot = i & 1;
        mbar_wait(bars + slot, (unsigned)((i >> 1) & 1));
        const double* cf = coefs + slot * fd.NC;
        double a[RI][RI];
        // assemble this lane's RI x RI piece of [S_i | u_i | rhs_i],  S_i = d_i - l_i X_{i-1}   (X_{i-1}, A_{i-1} live in Y)
#pragma unroll
        for (int ii = 0; ii < RI; ii++) {
            const int r = rg + 4 * ii;
            const bool rowJ = r < Bout, rowT = (r >= Bout && r < n);
            // row r of the stream: a band record (J rows), a species record (T rows), nothing (padding rows)
            const double* rr = rowJ ? cf + r * fd.W : cf + fd.o_T + (rowT ? r - Bout : 0) * fd.V;
            const double Dr = (r < n) ? rr[0] : 0., Lr = rowJ ? rr[1] : 0., Ur = rowJ ? rr[2] : 0., Br = rowJ ? rr[3] : (rowT ? rr[1] : 0.);
            const bool useX = rowJ && (i > 0) && (Lr != 0.);
#pragma unroll
            for (int j = 0; j < LB; j++) {
                const int cc = cg + 8 * j;
                double val = 0.;
                if (rowJ) {
                    if (cc < Bout) {
                        val = (cc == r) ? Dr : 0.;
                        if (useX) val = d_sub(val, d_mul(Lr, c.Y[r * c.LDY + cc]));
                    } else if (cc < n) val = rr[4 + (cc - Bout)];
                } else if (rowT) {
                    if (cc < Bout) val = cf[cc * fd.W + 4 + S + (r - Bout)];
                    else if (cc < n) val = (cc == r) ? Dr : rr[2 + (cc - Bout)];
                }
                a[ii][j] = val;
            }
#pragma unroll
            for (int j = LB; j < RI; j++) {
                const int cc = cg + 8 * (j - LB);
                double val = 0.;
                if (r < n) {
                    if (cc < Bout) val = (cc == r) ? Ur : 0.;
                    else if (cc == Bout) {
                        val = Br;
                        if (useX) val = d_sub(val, d_mul(Lr, c.Y[r * c.LDY + Bout]));
                    }
                }
                a[ii][j] = val;
            }
        }
        __syncwarp();
        if (lane == 0 && i + 2 < N2) bulk_g2s(coefs + slot * fd.NC, cf_g + (size_t)(i + 2) * fd.NC, coef_bytes, bars + slot);

        lu_phase<RI, 0>(a, c);
        lu_phase<RI, 1>(a, c);
        if constexpr (RI > 2) { lu_phase<RI, 2>(a, c); lu_phase<RI, 3>(a, c); }
        if constexpr (RI > 4) { lu_phase<RI, 4>(a, c); lu_phase<RI, 5>(a, c); }
        __syncwarp();

        // ---- back substitution with U, lane = right-hand-side column (Bout columns of u_i and the rhs): x stays in registers ----
        if (lane <= Bout) {
            double y[R];
#pragma unroll
            for (int k = 0; k < R; k++) y[k] = (k < n) ? c.Y[k * c.LDY + lane] : 0.;
#pragma unroll
            for (int k = R - 1; k >= 0; k--) {
                if (k < n) {
                    const double* uk = c.UT + k * c.LDU;            // column k of U, contiguous: 16-byte aligned (LDU even)
                    const double2* uk2 = reinterpret_cast<const double2*>(uk);
                    const bool nz = (y[k] != 0.);             // Eigen's triangular solve skips exactly-zero unknowns (no 0/0)
                    double num = nz ? y[k] : 1., den = nz ? uk[k] : 1.;
                    opaque(num); opaque(den);
                    const double zq = num / den;
                    const double z = nz ? zq : y[k];
                    y[k] = z;
#pragma unroll
                    for (int r = 0; r < k; r += 2) {
                        const double2 uu = uk2[r >> 1];
                        y[r] = d_fnma(z, uu.x, y[r]);
                        if (r + 1 < k) y[r + 1] = d_fnma(z, uu.y, y[r + 1]);
                    }
                }
            }
            double* Xi = Xg + (size_t)i * fd.XB;
#pragma unroll
            for (int k = 0; k < R; k++)
                if (k < n) { c.Y[k * c.LDY + lane] = y[k]; Xi[k * fd.LDXg + lane] = y[k]; }
        }
        __syncwarp();
    }

    // ---- back sweep  x_i = A_i - X_i x_{i+1}  (lane = row), solution staged eight cells at a time ------------------------------------
#ifndef AIO_EMU_BUILD
    __threadfence_block();
#endif
    fence_proxy_async();
    __syncwarp();
    double* ring = sm + fd.o_ring;
    double* xs = sm + fd.o_xs;
    double* stage = sm + fd.o_stage;
    if (lane == 0) {
        // the X blocks were written by this warp with ordinary stores: they are visible to the bulk copy after the fence above
        bulk_g2s(ring, Xg + (size_t)(N2 - 1) * fd.XB, x_bytes, bars + 2);
        if (N2 > 1) bulk_g2s(ring + fd.XB, Xg + (size_t)(N2 - 2) * fd.XB, x_bytes, bars + 3);
    }
    for (int q = 0; q < N2; q++) {
        const int i = N2 - 1 - q, slot = q & 1;
        mbar_wait(bars + 2 + slot, (unsigned)((q >> 1) & 1));
        const double* Xi = ring + slot * fd.XB;
        double x = 0;
        if (lane < n) {
            x = Xi[lane * fd.LDXg + Bout];
            if (q > 0) {
                double t = 0;
                for (int cc = 0; cc < Bout; cc++) t += Xi[lane * fd.LDXg + cc] * xs[cc];
                x = x - t;
            }
        }
        __syncwarp();
        if (lane < n) { xs[lane] = x; stage[lane * 8 + (i & 7)] = x; }
        __syncwarp();
        if (lane == 0 && q + 2 < N2) bulk_g2s(ring + slot * fd.XB, Xg + (size_t)(i - 2) * fd.XB, x_bytes, bars + 2 + slot);
        if ((i & 7) == 0 || false) {
            // flush cells i .. min(i+7, N2-1): J rows to Jrad_FLD, T rows to the temperature solution (both raw, see k_full_post)
            const int cnt = min(8, N2 - i);
            for (int e = lane; e < n * 8; e += 32) {
                const int r = e >> 3, o = e & 7;
                if (o < cnt) {
                    const double val = stage[r * 8 + o];
                    if (r < Bout) p.J[idx_J(p, col, r, i + o)] = val;
                    else p.Tsol[idx_s(p, col, r - Bout, i + o)] = val;
                }
            }
            __syncwarp();
        }
    }
}

// ---- negative-J scan, T clamp, e <- T, p <- e, cons <- prim, step tail (radiation.cpp:782-838, advection.cpp:282-346) ----------------
__global__ void __launch_bounds__(256) k_full_post(DevParams p) {
    __shared__ double red[32];
    __shared__ double rmin[32];
    const int col = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    ColState& cst = p.cs[col];
    if (!cst.in_step) return;
    const double dt = cst.dt;
    const int N = p.N, N2 = p.N2, S = p.S, Bout = p.Bout;
    TailAcc acc; acc.minstep = 0; acc.maxde = 0; acc.bad = 0; acc.badval = 0;
    int negJ = 0; double negJval = 1e308;
    for (int i = tid; i < N2; i += NT) {
        for (int b = 0; b < Bout; b++) {
            const double Jv = p.J[idx_J(p, col, b, i)];
            if (p.couple_J_into_T && i <= N - 1 && Jv < 0) {           // advection.cpp:308-338
                if (Jv > -1e-10) p.J[idx_J(p, col, b, i)] = Jv * -1.;
                else if (!(i < N && p.use_planetary_temperature && p.J[idx_J(p, col, b, N)] > 0)) { negJ = 1; negJval = fmin(negJval, Jv); }
            }
        }
        for (int s = 0; s < S; s++) {
            const SpeciesConst sc = p.sp[s];
            double T = p.Tsol[idx_s(p, col, s, i)];
            if (T < p.temperature_floor) T = p.temperature_floor;
            if (T > p.max_temperature) T = p.max_temperature;
            const double u1 = p.u[idx_u(p, 0, col, s, i)], u2 = p.u[idx_u(p, 1, col, s, i)], u3 = p.u[idx_u(p, 2, col, s, i)];
            Prim3 q = cons_to_prim(u1, u2, u3, sc.gm1);
            const double e_old = eint_of(q.p, q.rho, sc.gm1);
            const double cs_old = p.cssave ? p.cssave[idx_s(p, col, s, i)] : cs_of(e_old, sc.gm1);
            const double e = T * sc.cv;
            q.p = e * q.rho * sc.gm1;
            double n1, n2, n3;
            prim_to_cons(q, sc.gm1, n1, n2, n3);
            p.u[idx_u(p, 0, col, s, i)] = n1; p.u[idx_u(p, 1, col, s, i)] = n2; p.u[idx_u(p, 2, col, s, i)] = n3;
            tail_cell(p, col, s, i, q.v, cs_old, e, T, acc);
        }
    }
    const double ms = block_max(acc.minstep, red), md = block_max(acc.maxde, red);
    const int bad = block_or(acc.bad), nj = block_or(negJ);
    const double njv = -block_max(-negJval, rmin);
    __syncthreads();
    if (tid == 0) {
        // slab model of the planet's surface temperature (radiation.cpp:812-828), bands summed in order
        if (p.use_planetary_temperature) {
            double F = 0, S_surf = 0;
            for (int b = 0; b < Bout; b++) {
                F -= 4 * C_PI * p.dcore[(size_t)col * Bout + b] * (p.J[idx_J(p, col, b, p.ng)] - p.J[idx_J(p, col, b, p.ng - 1)]);
                S_surf += 0.25 * p.ssurf[(size_t)col * p.Bin + b];
            }
            const double S_tot = S_surf - F + C_SIGMA_RAD * (p.T_int * p.T_int) * (p.T_int * p.T_int);
            double Tp = cst.T_planet;
            if (p.core_cv > 0) { const double denom = 4 * C_SIGMA_RAD * Tp * Tp * Tp + p.core_cv / dt; Tp += S_tot / denom; }
            else Tp = pow(pow(Tp, 4) + S_tot / C_SIGMA_RAD, 0.25);
            cst.T_planet = Tp;
        }
        cst.minstep = ms; cst.maxde = md;
        cst.steps += 1;
        if (bad) { cst.status = COL_NEG_T; cst.t = 1.1 * cst.t_max; }
        else if (nj) { cst.status = COL_NEG_J; cst.crash_value = njv; cst.t = 1.1 * cst.t_max; }
    }
}

}  // namespace aio
