// CUDA front door of the engine sources.  The product build (nvcc, sm_100a) takes the first branch.  The second branch
// exists only for tests/emu (a CTA emulator used to debug kernel logic on the GPU-less dev container); it is selected by
// -DAIO_EMU_BUILD, which the product Makefile never passes.
#pragma once
#ifndef AIO_EMU_BUILD
#include <cuda_runtime.h>
#define AIO_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define AIO_DYN_SMEM(T, name) extern __shared__ __align__(16) unsigned char name##_raw_[]; T* name = reinterpret_cast<T*>(name##_raw_)
#else
#include "cuda_emu.h"
#define AIO_LAUNCH(kern, grid, block, smem, stream, ...) emu::launch(kern, dim3(grid), dim3(block), (smem), __VA_ARGS__)
#define AIO_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(emu::dyn_smem)
#endif
