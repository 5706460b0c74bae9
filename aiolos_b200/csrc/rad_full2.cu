// Translation unit of the second-generation full FLD kernels (rad_full2.cuh) and their launchers.
#define AIO_NO_SHARED_KERNELS
#include "rad_full2.cuh"
#include "cuda_check.h"

namespace aio {

void full2_configure(int RI, size_t smem_lu, size_t smem_asm) {
    if (RI == 2) CK(cudaFuncSetAttribute(k_full_lu<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lu));
    else if (RI == 4) CK(cudaFuncSetAttribute(k_full_lu<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lu));
    else CK(cudaFuncSetAttribute(k_full_lu<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lu));
    CK(cudaFuncSetAttribute(k_full_asm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_asm));
}

void full2_launch(const DevParams& p, int RI, size_t smem_lu, size_t smem_asm, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AIO_LAUNCH(k_full_heat, (p.ncol + 3) / 4, 128, 0, stream, p);
    AIO_LAUNCH(k_full_asm, p.ncol, 32 * FULL_ASM_WARPS, smem_asm, stream, p);
    if (RI == 2) AIO_LAUNCH(k_full_lu<2>, p.ncol, 32, smem_lu, stream, p);
    else if (RI == 4) AIO_LAUNCH(k_full_lu<4>, p.ncol, 32, smem_lu, stream, p);
    else AIO_LAUNCH(k_full_lu<6>, p.ncol, 32, smem_lu, stream, p);
    AIO_LAUNCH(k_full_post, p.ncol, 256, 0, stream, p);
}

}  // namespace aio
