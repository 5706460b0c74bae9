// Translation unit of the second-generation full FLD kernels (rad_full2.cuh) and their launchers.
#define AIO_NO_SHARED_KERNELS
#include "rad_full2.cuh"
#include "cuda_check.h"
#include <algorithm>
#include <cstring>
#include <vector>

namespace aio {

void full2_configure(int RI, size_t smem_lu, size_t smem_heat2) {
    if (smem_heat2) {
        CK(cudaFuncSetAttribute(k_full_heat2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_heat2));
        CK(cudaFuncSetAttribute(k_full_heat2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_heat2));
        CK(cudaFuncSetAttribute(k_full_heat2<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_heat2));
        CK(cudaFuncSetAttribute(k_full_heat2<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_heat2));
    }
    if (RI == 2) CK(cudaFuncSetAttribute(k_full_lu<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lu));
    else if (RI == 4) CK(cudaFuncSetAttribute(k_full_lu<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lu));
    else CK(cudaFuncSetAttribute(k_full_lu<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lu));
}

void full2_launch(const DevParams& p, int RI, size_t smem_lu, size_t smem_heat2, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (smem_heat2) {
        const int nt = std::min(512, (p.N2 + 31) / 32 * 32);          // one cell per thread where the column fits (no half-empty second pass)
        if (p.S <= 1) AIO_LAUNCH(k_full_heat2<1>, p.ncol, nt, smem_heat2, stream, p);
        else if (p.S <= 2) AIO_LAUNCH(k_full_heat2<2>, p.ncol, nt, smem_heat2, stream, p);
        else if (p.S <= 4) AIO_LAUNCH(k_full_heat2<4>, p.ncol, nt, smem_heat2, stream, p);
        else AIO_LAUNCH(k_full_heat2<16>, p.ncol, nt, smem_heat2, stream, p);
    } else
    AIO_LAUNCH(k_full_heat, (p.ncol + 3) / 4, 128, 0, stream, p);
    {
        const unsigned grid = (unsigned)(((long long)p.ncol * p.N2 + 127) / 128);
        if (p.S <= 1) AIO_LAUNCH(k_full_asm2<1>, grid, 128, 0, stream, p);
        else if (p.S <= 2) AIO_LAUNCH(k_full_asm2<2>, grid, 128, 0, stream, p);
        else if (p.S <= 4) AIO_LAUNCH(k_full_asm2<4>, grid, 128, 0, stream, p);
        else AIO_LAUNCH(k_full_asm2<16>, grid, 128, 0, stream, p);
    }
    if (RI == 2) AIO_LAUNCH(k_full_lu<2>, p.ncol, 32, smem_lu, stream, p);
    else if (RI == 4) AIO_LAUNCH(k_full_lu<4>, p.ncol, 32, smem_lu, stream, p);
    else AIO_LAUNCH(k_full_lu<6>, p.ncol, 32, smem_lu, stream, p);
    AIO_LAUNCH(k_full_post, p.ncol, 256, 0, stream, p);
}

// Block-Thomas sweep on caller-supplied block rows (aio_selftest_block_thomas): `coef` holds, per column and cell, the
// coefficient stream k_full_asm would have written (band and species records, see Full2Dims in rad_full2.h); x receives the solution [ncols][N2][n] (J rows, then T rows).
void full2_selftest(int ncols, int N2, int Bout, int S, const double* coef, double* x) {
    if (ncols < 1 || N2 < 1 || Bout < 1 || S < 1 || Bout + S > 24 || Bout > MAXB || S > MAXS) throw std::invalid_argument("block_thomas: sizes out of range");
    const Full2Dims fd = full2_dims(Bout, S);
    DevParams p{};
    p.ncol = ncols; p.N2 = N2; p.N = N2 - 2; p.NP = (N2 + 31) / 32 * 32; p.S = S; p.Bout = Bout; p.Bin = 1;
    std::vector<ColState> cs(ncols);
    for (auto& c : cs) { memset(&c, 0, sizeof c); c.in_step = 1; }
    const size_t n_coef = (size_t)ncols * N2 * fd.NC, n_X = (size_t)ncols * N2 * fd.XB;
    const size_t n_J = (size_t)ncols * Bout * p.NP, n_T = (size_t)ncols * S * p.NP;
    std::vector<void*> allocs;
    auto dal = [&](size_t bytes) { void* q = nullptr; CK(cudaMalloc(&q, bytes ? bytes : 8)); allocs.push_back(q); CK(cudaMemsetAsync(q, 0, bytes ? bytes : 8, nullptr)); return q; };
    try {
        p.cs = (ColState*)dal(cs.size() * sizeof(ColState));
        p.fcoef = (double*)dal(n_coef * 8); p.Xstore = (double*)dal(n_X * 8);
        p.J = (double*)dal(n_J * 8); p.Tsol = (double*)dal(n_T * 8);
        CK(cudaMemcpy(p.cs, cs.data(), cs.size() * sizeof(ColState), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(p.fcoef, coef, n_coef * 8, cudaMemcpyHostToDevice));
        const size_t smem_lu = (size_t)fd.sm_warp * sizeof(double);
        full2_configure(fd.RI, smem_lu, 0);
        cudaStream_t stream = nullptr;
        if (fd.RI == 2) AIO_LAUNCH(k_full_lu<2>, ncols, 32, smem_lu, stream, p);
        else if (fd.RI == 4) AIO_LAUNCH(k_full_lu<4>, ncols, 32, smem_lu, stream, p);
        else AIO_LAUNCH(k_full_lu<6>, ncols, 32, smem_lu, stream, p);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        std::vector<double> hJ(n_J), hT(n_T);
        CK(cudaMemcpy(hJ.data(), p.J, n_J * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hT.data(), p.Tsol, n_T * 8, cudaMemcpyDeviceToHost));
        const int n = Bout + S;
        for (int c = 0; c < ncols; c++)
            for (int i = 0; i < N2; i++) {
                double* xi = x + ((size_t)c * N2 + i) * n;
                for (int b = 0; b < Bout; b++) xi[b] = hJ[idx_J(p, c, b, i)];
                for (int s = 0; s < S; s++) xi[Bout + s] = hT[idx_s(p, c, s, i)];
            }
    } catch (...) { for (void* q : allocs) cudaFree(q); throw; }
    for (void* q : allocs) cudaFree(q);
}

}  // namespace aio
