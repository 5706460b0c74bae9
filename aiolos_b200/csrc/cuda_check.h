// CUDA error check shared by the host translation units of the engine.
#pragma once
#include "cuda_compat.h"
#include <sstream>
#include <stdexcept>
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    std::ostringstream os_; os_ << "CUDA error " << cudaGetErrorString(e_) << " at " << __FILE__ << ":" << __LINE__; \
    throw std::runtime_error(os_.str()); } } while (0)
