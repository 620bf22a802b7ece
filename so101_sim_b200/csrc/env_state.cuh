// Device-resident state of N lockstep environments, structure-of-arrays: field[k][env] so that a warp touching the
// same component of 32 consecutive envs issues one coalesced 128 B (f32) / 256 B (f64) request.
#pragma once
#include <cstdint>
#include "../../include/so101_b200.h"

namespace so101 {

// The integration state (qpos, qvel and the reset pool) is float64 in BOTH precisions.  The float32 product path runs
// collision, the contact rows and the Newton solver in float32 (> 95 % of the arithmetic), but the arm's smooth dynamics
// (FK, CRB, RNE, actuators, M^-1), and the semi-implicit Euler update in float64: the reference's actuators are anti-damped
// (bias +1 * qvel, scene_pbr.xml:11) and amplify a perturbation ~300x over 100 control steps, so that neither a float32 state
// (1.8e-3 relative error after 100 steps) nor float32 smooth dynamics on a float64 state (1.6e-4) meet north_star's 1e-4;
// float64 smooth dynamics + float32 solver give 9e-6 (tools/exp_arm_precision.py).
using TS = double;

template <typename T>
struct EnvState {
  int N, nq, nv;
  TS *qpos, *qvel;                  // [nq][N], [nv][N]
  T *warm;                          // [nv][N]  (qacc_warmstart)
  TS *init_qpos, *init_qvel;        // reset pool [npool][...same layout as qpos / qvel...]: episode e of an env starts from entry e % npool
  int npool;                        // entries in the reset pool (>= 1)
  int *episode;                     // [N] resets this env has gone through (selects the pool entry)
  T *ctrl;                          // [6][N]
  int *step;                        // [N] control steps since reset
  uint8_t *needs_reset;             // [N] previous step was LAST (dm_control auto-reset on next step)
  float *ring_joints;               // [Dj+1][6][N]
  float *ring_phys;                 // [Dp+1][nq+nv][N]
  int *diverged_count;              // [1]
  int *solver_iter;                 // [N] Newton iterations of the last substep (parity/diagnostics)
  int *ncon;                        // [N] contacts of the last substep
  unsigned long long *prof;         // optional [16] stage-profile accumulators (SO101_PROFILE=1), see scene_kernel.inl
  float *dbg_contacts;              // optional [N][1 + 9*NCON] parity probe (null unless requested)
};

struct StepCfg {
  int nsub, last_step, dj, dp, terminate_on_success, max_iter;
  int arm_mode;  // developer switch (SO101_ARM_MODE): which parts of the float32 arm path run in float64, see arm_kernel.cu
  float tol;
  int dbg_env, dbg_step;  // developer probe: device printf of one env's manifold inputs (SO101_DBG_ENV / SO101_DBG_STEP)
  float offsets[6], home[6];
};

}  // namespace so101
