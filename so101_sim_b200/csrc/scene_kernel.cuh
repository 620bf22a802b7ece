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

// Device scratch that crosses the kernels of one substep (written by one kernel, read by the next; L2-resident).
template <typename T>
struct PipeBuf {
  T *xpos, *xmat;           // [N][NSLOT*3], [N][NSLOT*9]  world poses of the 8 dynamic bodies
  uint2 *work;              // [N*PAIRCAP]  narrow-phase work list: (env, g1 | g2 << 8 | pair index << 16)
  int *nwork;               // [nsub+1][4]  per substep: pairs appended, pair cursor, large-tier envs queued, large-tier cursor
  int *big;                 // [N]  envs deferred to the large solver tier in this substep
  T *con;                   // [N][CONBUF][8]  raw contacts: normal3, pos3, dist
  int *con_key;             // [N][CONBUF]     pair index << 20 | manifold index << 16 | g1 << 8 | g2  (sort key)
  int *ncon_raw;            // [N]
  uint8_t *active, *flags;  // [N]  env steps this call (not being reset) / env diverged during this control step
  int narrow_grid;          // CTAs of the persistent narrow-phase kernel
};

template <typename T>
int launch_scene_step(const ArmModelT<T> &am, const SceneModel<T> &sm, const StepCfg &cfg, const EnvState<T> &S, const PipeBuf<T> &pb,
                      const float *action, const so101_step_out &out, cudaStream_t stream);
template <typename T>
void launch_scene_reset(const StepCfg &cfg, const EnvState<T> &S, const uint8_t *mask, const so101_step_out &out, cudaStream_t stream);
template <typename T>
size_t scene_smem_bytes();
template <typename T>
int scene_narrow_grid();
}  // namespace so101
