// Measured denominators for the FP64 roofline of the dense-solve kernels (friction LU, block Thomas): the chip's DFMA
// issue rate and its HBM copy bandwidth, timed with CUDA events on the calling device.  MEASURED_PEAKS.json (driver-written)
// carries HBM and bf16 tensor figures only; the FP64 pipe peak is measured here, live, by the same process that reports
// against it.
#include "cuda_check.h"
#include <algorithm>

namespace aio {

// 8 independent DFMA chains per thread, 512 threads per CTA, 4 CTAs per SM: enough warps and ILP to keep the FP64 pipe issuing
// every cycle; 2 flop per DFMA.
__global__ void __launch_bounds__(512) k_dfma_peak(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[0] = s;          // never true; keeps the chains alive
}

__global__ void k_copy_peak(const double2* __restrict__ src, double2* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

void measure_device_peaks(int device, double* fp64_tflops, double* copy_gbs) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw std::runtime_error("CUDA: no device (aiolos_b200 has no CPU path)");
    CK(cudaSetDevice(device));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double* d = nullptr; CK(cudaMalloc(&d, 64));
    const int iters = 4096, grid = sms * 4;
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(e0));
        k_dfma_peak<<<grid, 512>>>(d, iters, 0.999999, 1e-7);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double flop = 2.0 * 64.0 * iters * 512.0 * grid;
        if (rep) best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    if (fp64_tflops) *fp64_tflops = best;
    CK(cudaFree(d));
    if (copy_gbs) {
        const size_t n = (size_t)1 << 26;        // 2 x 1 GiB
        double2 *a = nullptr, *b = nullptr;
        CK(cudaMalloc(&a, n * sizeof(double2))); CK(cudaMalloc(&b, n * sizeof(double2)));
        CK(cudaMemset(a, 0, n * sizeof(double2)));
        double bw = 0;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaEventRecord(e0));
            k_copy_peak<<<sms * 16, 512>>>(a, b, n);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep) bw = std::max(bw, 2.0 * n * sizeof(double2) / (ms * 1e-3) / 1e9);
        }
        *copy_gbs = bw;
        CK(cudaFree(a)); CK(cudaFree(b));
    }
    CK(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

}  // namespace aio
