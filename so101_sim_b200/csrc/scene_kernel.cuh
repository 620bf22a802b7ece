#pragma once
#include <cuda_runtime.h>
#include "arm_dynamics.cuh"
#include "arm_solver.cuh"
#include "env_state.cuh"
#include "scene_collide.cuh"
#include "scene_model.cuh"

namespace so101 {
// Scene-kernel state layout is array-of-rows ([N][nq] etc.): one warp owns one env and reads its row with one
// coalesced request (the arm-only kernel, one THREAD per env, uses [k][N] instead).
template <typename T>
void launch_scene_step(const ArmModelT<T> &am, const SceneModel<T> &sm, const StepCfg &cfg, const EnvState<T> &S, const float *action,
                       const so101_step_out &out, cudaStream_t stream);
template <typename T>
void launch_scene_reset(const StepCfg &cfg, const EnvState<T> &S, const uint8_t *mask, const so101_step_out &out, cudaStream_t stream);
template <typename T>
size_t scene_smem_bytes();
}  // namespace so101
