// warp_kernels.cuh -- warp-per-env bit-plane kernels (boards up to 32x32).
#pragma once
#include <cuda_runtime.h>
#include "cells.cuh"
#include "../../include/safelife_b200.h"

namespace sl {

constexpr int WARP_MAX_DIM = 32;

inline bool warp_kernel_supports(int H, int W) {
    (void)H; (void)W;
    return false;   // enabled once the bit-plane kernels land
}

inline void launch_advance_warp(const uint16_t *, uint16_t *, int32_t *, int, int, int, uint64_t,
                                int, uint64_t *, cudaStream_t) {}
inline void launch_env_step_warp(const SlEnvBatch &, const uint8_t *, cudaStream_t) {}

}  // namespace sl
