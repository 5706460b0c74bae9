// Full FLD solver (PARI_USE_RADIATION 1): coupled J_b / T_s block-tridiagonal system.  See rad_full kernels below.
#pragma once
#include "common.cuh"
namespace aio {
}  // namespace aio
