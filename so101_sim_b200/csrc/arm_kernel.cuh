#pragma once
#include <cuda_runtime.h>
#include "arm_dynamics.cuh"
#include "arm_solver.cuh"
#include "env_state.cuh"

namespace so101 {
template <typename T>
void launch_arm_step(const ArmModelT<T> &am, const ArmModelT<double> &am64, const StepCfg &cfg, const EnvState<T> &S, const float *action,
                     const so101_step_out &out, cudaStream_t stream);
template <typename T>
void launch_arm_reset(const StepCfg &cfg, const EnvState<T> &S, const uint8_t *mask, const so101_step_out &out, cudaStream_t stream);
}  // namespace so101
