// State-dependent opacity models (reference opacities.cpp:74-369): the reference's update_opacities() pass as a kernel.
//   T  tabulated kappa_S, kappa_P, kappa_R (p, T) from .aiopa tables      opacities.cpp:268-282
//   K  tabulated kappa_P, kappa_R; band-mean kappa_S from the .opa file   opacities.cpp:283-314
//   D  single-grain-size dust toy model                                    opacities.cpp:215-267
// interpol_tabulated_opacity (opacities.cpp:892-983) is deliberately mixed precision in the reference (float corner values,
// float pressure with a fast float log10, double temperature weights); it is restated operation by operation.
#pragma once
#include "common.cuh"
#include "hydro.cuh"

namespace aio {

__device__ __forceinline__ double kapv(const DevParams& p, int col, int s, int idx, int j) {
    const int KA = p.Bin + 2 * p.Bout;
    return p.kap_cell ? p.kap_cell[((((size_t)col * p.S + s) * KA + idx) * p.NP) + j] : p.kappa_avg[s * KA + idx];
}

#ifdef AIO_EMU_BUILD
inline float int_as_float_(int a) { float r; memcpy(&r, &a, sizeof r); return r; }
inline int float_as_int_(float a) { int r; memcpy(&r, &a, sizeof r); return r; }
#else
__device__ __forceinline__ float int_as_float_(int a) { return __int_as_float(a); }
__device__ __forceinline__ int float_as_int_(float a) { return __float_as_int(a); }
#endif

// aiolos.h:110-131 (njuffa's fast natural log) and njuffas_log10f
__device__ __forceinline__ float njuffas_logf_(float a) {
    float m, r, s, t, i, f;
    int e;
    e = (float_as_int_(a) - 0x3f2aaaab) & 0xff800000;
    m = int_as_float_(float_as_int_(a) - e);
    i = (float)e * 1.19209290e-7f;
    f = m - 1.0f;
    s = f * f;
    r = fmaf(0.230836749f, f, -0.279208571f);
    t = fmaf(0.331826031f, f, -0.498910338f);
    r = fmaf(r, s, t);
    r = fmaf(r, s, f);
    r = fmaf(i, 0.693147182f, r);
    return r;
}
__device__ __forceinline__ float njuffas_log10f_(float a) {
    const float log10ff = (float)2.302585092994046;       // const float log10ff = std::log(10.)
#ifdef AIO_EMU_BUILD
    return njuffas_logf_(a) / log10ff;
#else
    return __fdiv_rn(njuffas_logf_(a), log10ff);
#endif
}

// opacities.cpp:892-983
__device__ inline double interpol_tabulated_opacity(const DevParams& p, const OpaTable& t, int off_array, int band, double T_gas, double pressure) {
    const double* tg = p.opa_data + t.off_t;
    const double* pg = p.opa_data + t.off_p;
    const double* plog = p.opa_data + t.off_plog;
    const double* arr = p.opa_data + off_array;
    double tT = T_gas, denom;
    float tP = (float)pressure;
    int Plow = 0, Phigh = 0, Tlow = 0, Thigh = 0;
    if (tT < tg[0]) tT = tg[0];
    else if (tT > tg[t.nT - 1]) tT = tg[t.nT - 1];
    if (tP < pg[0]) tP = (float)pg[0];
    else if (tP > pg[t.np - 1]) tP = (float)pg[t.np - 1];
    for (int i = t.nT - 2; i >= 0; i--) if (tT >= tg[i]) { Tlow = i; Thigh = i + 1; break; }
    if (Thigh == t.nT) { Thigh = t.nT - 1; Tlow = t.nT - 2; tT = tg[Thigh]; }
    for (int i = t.np - 2; i >= 0; i--) if (tP >= pg[i]) { Plow = i; Phigh = i + 1; break; }
    if (Phigh == t.np) { Phigh = t.np - 1; Plow = t.np - 2; tP = (float)pg[Phigh]; }
    const float k1 = (float)arr[Tlow + Plow * t.nT + band * t.nT * t.np];
    const float k2 = (float)arr[Thigh + Plow * t.nT + band * t.nT * t.np];
    const float k3 = (float)arr[Tlow + Phigh * t.nT + band * t.nT * t.np];
    const float k4 = (float)arr[Thigh + Phigh * t.nT + band * t.nT * t.np];
    const float pmax = (float)plog[Phigh];
    const float pmin = (float)plog[Plow];
    denom = ((tg[Thigh] - tg[Tlow]) * (double)(pmax - pmin));
    denom = 1.0 / denom;
    float ktotal = 0;
    const float muls[3] = {(float)0.99, (float)1.00, (float)1.01};
    for (int i = 0; i < 3; i++) {
        const float tpp = tP * muls[i];
        const float pval = njuffas_log10f_(tpp);
        const double mul1 = (tg[Thigh] - tT) * (double)(pmax - pval);
        const double mul2 = (tT - tg[Tlow]) * (double)(pmax - pval);
        const double mul3 = (tg[Thigh] - tT) * (double)(pval - pmin);
        const double mul4 = (tT - tg[Tlow]) * (double)(pval - pmin);
        ktotal = (float)((double)ktotal + denom * (mul1 * (double)k1 + mul2 * (double)k2 + mul3 * (double)k3 + mul4 * (double)k4));
    }
    return pow(10., (double)ktotal * 0.33333333333333333);
}

// update_opacities for every species and cell of one column: CTA per column, thread per cell
#ifndef AIO_NO_SHARED_KERNELS      // defined by the translation units that only borrow the device functions
__global__ void __launch_bounds__(256) k_opacity(DevParams p) {
    const int col = blockIdx.x;
    const ColState& c = p.cs[col];
    if (!c.in_step) return;
    const int KA = p.Bin + 2 * p.Bout;
    for (int s = 0; s < p.S; s++) {
        const SpeciesConst sc = p.sp[s];
        const OpaTable tb = p.opa_tab[s];
        double* K = p.kap_cell + (((size_t)col * p.S + s) * KA) * p.NP;
        for (int j = threadIdx.x; j < p.N2; j += blockDim.x) {
            const double u1 = p.u[idx_u(p, 0, col, s, j)], u2 = p.u[idx_u(p, 1, col, s, j)], u3 = p.u[idx_u(p, 2, col, s, j)];
            Prim3 q = cons_to_prim(u1, u2, u3, sc.gm1);
            const double T = p.Tsave ? p.Tsave[idx_s(p, col, s, j)] : eint_of(q.p, q.rho, sc.gm1) / sc.cv;
            if (p.opacity_model == 'T' || p.opacity_model == 'K') {
                for (int b = 0; b < p.Bin; b++)
                    K[(size_t)b * p.NP + j] = (p.opacity_model == 'T')
                        ? p.kS_factor * interpol_tabulated_opacity(p, tb, tb.off_solar, b, T, q.p) * tb.inv_mass
                        : p.kappa_avg[s * KA + b];
                for (int b = 0; b < p.Bout; b++) {
                    K[(size_t)(p.Bin + b) * p.NP + j] = p.kP_factor * interpol_tabulated_opacity(p, tb, tb.off_planck, b, T, q.p) * tb.inv_mass;
                    K[(size_t)(p.Bin + p.Bout + b) * p.NP + j] = p.kR_factor * interpol_tabulated_opacity(p, tb, tb.off_ross, b, T, q.p) * tb.inv_mass;
                }
            } else {                                  // 'D'
                double ks = 1e-4, k = 1e-4;
                if (tb.is_dust_like) {
                    const double RHO = 1;
                    const double size = pow(3 / (4 * M_PI) * tb.mass_amu * C_AMU / RHO, 1 / 3.);
                    const double k0 = 3 / (4 * RHO * size);
                    const double T0 = 1 / (2 * M_PI * 3.475 * size);
                    ks = k0 * p.T_star / (T0 + p.T_star);
                    k = k0 * T / (T0 + T);
                }
                for (int b = 0; b < p.Bin; b++) K[(size_t)b * p.NP + j] = ks;
                for (int b = 0; b < p.Bout; b++) { K[(size_t)(p.Bin + b) * p.NP + j] = k; K[(size_t)(p.Bin + p.Bout + b) * p.NP + j] = k; }
            }
        }
    }
}
#endif

}  // namespace aio
