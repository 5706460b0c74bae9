// Host-visible part of the second-generation full FLD solver (rad_full2.cuh / rad_full2.cu): the shared-memory and HBM
// geometry both sides agree on, and the launchers.  The kernels live in their own translation unit so that their code
// generation does not depend on what else the engine's translation unit holds (the 16 k-instruction hydro kernels next to
// it changed ptxas' register allocation of the LU sweep: 7.4 k instead of 4.4 k instructions, ncu r2c).
#pragma once
#include "common.cuh"

namespace aio {

struct Full2Dims {
    int n, R, RI;
    // Coefficient stream of one cell (k_full_asm2 -> k_full_lu): Bout band records of W doubles
    //   [D, L, U, B, JT(s = 0..S-1), TJ(s = 0..S-1), pad]     D: diagonal of d_i, L / U: the (diagonal) l_i / u_i entry, B: rhs,
    //                                                          JT = d_i(J_b, T_s), TJ = d_i(T_s, J_b)
    // followed by S species records of V doubles  [D, B, TT(t = 0..S-1), pad]  (TT: off-diagonal T-T couplings; TT(s, s) = 0).
    // W and V are multiples of 4 doubles, so a thread writes a record as whole 32-byte sectors.
    int W, V, o_T, NC;
    int LDXg, XB;                  // row stride (odd => conflict-free lane = row reads) and block size (even) of X_i in HBM
    int LDU, LDY;
    int sm_warp;                   // doubles of shared memory per column in k_full_lu
    int o_col, o_l, o_prow, o_swap, o_UT, o_Y, o_coef, o_bar;      // forward-sweep carve-up
    int o_ring, o_xs, o_stage;                                      // back-sweep carve-up (aliases the forward one)
};
// RI is a function of n; the kernels pass their template parameter so that everything that depends only on RI folds
__host__ __device__ inline Full2Dims full2_dims_ri(int Bout, int S, int RI) {
    Full2Dims d;
    d.n = Bout + S;
    d.RI = RI;
    d.R = 4 * d.RI;
    d.W = (4 + 2 * S + 3) & ~3; d.V = (2 + S + 3) & ~3;
    d.o_T = Bout * d.W; d.NC = d.o_T + S * d.V;
    d.LDXg = (Bout + 1) | 1;
    d.XB = d.n * d.LDXg; d.XB += d.XB & 1;
    d.LDU = d.R + 2; d.LDY = d.R;
    int o = 0;
    d.o_col = o; o += 32;
    d.o_l = o; o += d.R;
    d.o_prow = o; o += 8 * d.RI;
    d.o_swap = o; o += 8 * d.RI;
    d.o_UT = o; o += d.R * d.LDU;
    d.o_Y = o; o += d.R * d.LDY;
    d.o_coef = o; o += 2 * d.NC;
    d.o_bar = o; o += 4;
    const int fwd = o;
    o = 0;
    d.o_ring = o; o += 2 * d.XB;
    d.o_xs = o; o += 32;
    d.o_stage = o; o += d.n * 8;
    const int back = o;
    // the mbarriers keep their place during the back sweep
    d.sm_warp = (fwd > back + 4) ? fwd : back + 4;
    if (back > d.o_bar) { d.o_bar = back; }
    d.sm_warp += d.sm_warp & 1;
    return d;
}
__host__ __device__ inline Full2Dims full2_dims(int Bout, int S) {
    const int n = Bout + S;
    return full2_dims_ri(Bout, S, (n <= 8) ? 2 : (n <= 16) ? 4 : 6);
}

// cudaFuncSetAttribute for the four kernels; launch of one radiation step on `stream`
#ifndef FULL_LU_CTAS_PER_SM
#define FULL_LU_CTAS_PER_SM 16        // resident single-warp CTAs per SM k_full_lu is compiled for (128 registers)
#endif
void full2_configure(int RI, size_t smem_lu, size_t smem_heat2);
void full2_launch(const DevParams& p, int RI, size_t smem_lu, size_t smem_heat2, void* stream);     // smem_heat2 = 0: k_full_heat

void full2_selftest(int ncols, int N2, int Bout, int S, const double* coef, double* x);

}  // namespace aio
