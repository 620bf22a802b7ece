// Control step of the FULL contact scene (BASELINE config 3: SO100HandOverBanana, nq=20 nv=18) as a pipeline of small
// kernels per physics substep, all envs in lockstep (DESIGN.md section 4 has the table):
//
//   scene_begin_kernel       (once per control step, warp per env)   auto-reset from the reset pool, action -> ctrl
//   scene_kindyn_kernel      (thread per env)                        arm forward kinematics, CRB mass matrix, RNE bias, actuation,
//                                                                    prop mass blocks, qacc_smooth, friction / limit rows -> dyn record
//   scene_broad_kernel       (warp per env)                          broad + mid phase -> per-geom work queues; after the last
//                                                                    substep instead: the task layer (delay rings, reward, flags)
//   scene_gjk_kernel         (thread per candidate geom pair)        boolean GJK over the work queues -> hit slots per queue
//   scene_narrow_seq_kernel  (thread per intersecting pair)          EPA -> support-feature clipping manifold -> raw contacts
//   scene_classify_kernel    (thread per env)                        solver tier by contact / Jacobian-block count
//   scene_solve_kernel + scene_solve_tier_kernel x2 (warp per env, concurrent streams)
//                            contact gather, constraint rows, elliptic-cone Newton, semi-implicit Euler
//
// The arm's kinematics and smooth dynamics are long straight-line scalar code: one THREAD per env runs them once per 32 envs
// instead of 32 lanes of a warp running them redundantly, and the warp-per-env solve kernel is left with loop-structured
// code that is a third of its former size (it was bound by instruction fetch: profiles/r01u_ncu_scene_solve_kernel.txt).  The envs of a handle are processed as two
// pipeline groups on separate streams (abi.cu), so the launches below exist once per group.
// Splitting by stage keeps each kernel's code and shared-memory footprint small (more resident warps, less instruction-
// cache thrash) and turns the narrow phase - whose cost varies 10x between pairs - into flat work lists.  What crosses
// kernels (body poses, pair queues, hit list, raw contacts; a few KB per env) is written once and read once.
// Per substep order ([upstream] mj_step; legacy dm_control order is equivalent):
//   arm FK / CRB / RNE            (one thread per env, registers; arm_dynamics.cuh)
//   prop kinematics, M, bias      (free joints: linear dofs world frame, angular dofs body frame)
//   collision                     (scene_collide.cuh, scene_collide_seq.cuh)
//   constraint rows               (lane per contact: parameter mixing, impedance, Jacobian blocks, aref)
//   Newton solve                  (scene_solve.cuh)
//   semi-implicit Euler           (quaternion integration for the free joints)
// Task layer: so100_task.py:266-368, so100_hand_over.py:238-275.
#include <cstdlib>
#include <string>

#include "scene_kernel.cuh"
#include "scene_collide_seq.cuh"
#include "scene_solve.cuh"

namespace so101 {

// envs (warps) per CTA in the tier-0 solve kernel.  One: a two-env CTA keeps the slot of its faster env (or of an env that belongs
// to a larger tier) idle until the slower one is done; measured at 131072 envs late in a random-action rollout, when a third of
// the envs are in the larger tiers: 623 k -> 654 k env-steps/s.  (A persistent variant that takes tier-0 envs from a queue,
// as the larger tiers do, was slower - 634 k: its CTAs hold every solver slot of the SM until the queue is empty, and the
// larger tiers' CTAs no longer run beside them.)
constexpr int WARPS_SOLVE = 1;
// Solver tiers by contact capacity: a resting scene has ~20 contacts, an arm pressed into the table or props 40-100.  Each
// tier is the same code with a larger shared-memory scratch; an env that does not fit tier t is queued for tier t + 1.
constexpr int NC_S = 32, NB_S = 36;            // tier 0: every env that fits; 13.2 KB of scratch = 16 single-warp CTAs per SM
constexpr int NC_M = 64, NB_M = 80;            // tier 1
constexpr int NC_L = CONBUF, NB_L = CONBUF + 32;  // tier 2 (last: excess contacts are dropped and counted)
constexpr int WARPS_M = 1, WARPS_L = 1;  // single-warp CTAs: small enough to co-reside with tier-0 CTAs on an SM

constexpr int GMAX = GMAX_GEOMS;  // geoms the broad phase can hold (model: 83 colliding geoms)
constexpr int CANDCAP = 256;  // geom pairs that may survive the bounding-sphere test per env

template <typename T>
struct BroadScratch {
  unsigned pairq[PAIRCAP];
  unsigned cand[CANDCAP];
  T gcenter[3][GMAX];                  // world bounding-sphere centres of all geoms
  T opos[3][GMAX], omat[9][GMAX];      // world oriented boxes (geom AABB in the geom frame) of all geoms
  T bcenter[3][BMAX_BODIES];           // world bounding-sphere centres of the bodies
};

// per-env scratch of the begin / solve kernels (NC contacts, NB Jacobian blocks)
template <typename T, int NC, int NB>
struct Scratch {
  // the poses are only read while the constraint rows are built, the Hessian only exists afterwards: they share storage
  // (552 bytes less per env let 8 instead of 7 two-env CTAs of the tier-0 kernel fit the 228 KB of an SM)
  union {
    struct { T xpos[NSLOT][3], xmat[NSLOT][9], arm_p[NA][3], arm_a[NA][3]; };
    T H[NH];
  };
  ArmRows<T> arows[NARM];
  T q[NQ], qd[NV], warm[NV];
  T Mprop[NPROP][21];
  T Marm[NARM][21];
  T qacc_s[NV], delta[NV], grad[NV], search[NV], Md[NV], hscale[NV];
  int ncon, dbg, profon;
  long long prof[16];  // developer probe (SO101_PROFILE=1): per-stage clock64 sums and counters of this env
  SolveScratch<T, NC, NB> sol;
};

// per-env scratch of the broad-phase / task-layer kernel
template <typename T>
struct BroadEnv {
  T xpos[NSLOT][3], xmat[NSLOT][9];
  T q[NQ], qd[NV], ctrl[NA];
  int profon;
  long long prof[16];
  BroadScratch<T> broad;
};

// stage profiler: lane 0 accumulates clock64 deltas into Scratch::prof when the handle was created with SO101_PROFILE=1
enum { P_DYN = 0, P_BROAD, P_PLANE, P_GJK, P_EPA, P_MANI, P_ROWS, P_SOLVE, P_INTEG, P_TASK, P_NPQ, P_NCON, P_NEWTON, P_LINE, P_NEPA, P_NSUB };
enum { P_GJKIT = P_PLANE, P_EPAIT = P_INTEG, P_BIGENV = P_TASK };  // counters that re-use the (tiny) plane / integrate / task slots
#define PROF_START(s) long long pt_ = (s).profon ? clock64() : 0
#define PROF_ACC(s, i, lane) do { if ((s).profon) { const long long n_ = clock64(); if ((lane) == 0) (s).prof[i] += n_ - pt_; pt_ = n_; } } while (0)
#define PROF_CNT(s, i, v, lane) do { if ((s).profon && (lane) == 0) (s).prof[i] += (v); } while (0)

// ------------------------------------------------------------------------------------------------ kinematics + smooth dynamics
template <typename T>
__device__ __forceinline__ void prop_rotation(const T *quat, T *R) {
  T w = quat[0], x = quat[1], y = quat[2], z = quat[3];
  const T n = t_sqrt(w * w + x * x + y * y + z * z);
  w /= n; x /= n; y /= n; z /= n;
  R[0] = w * w + x * x - y * y - z * z; R[4] = w * w - x * x + y * y - z * z; R[8] = w * w - x * x - y * y + z * z;
  R[1] = T(2) * (x * y - w * z); R[2] = T(2) * (x * z + w * y); R[3] = T(2) * (x * y + w * z);
  R[5] = T(2) * (y * z - w * x); R[6] = T(2) * (x * z - w * y); R[7] = T(2) * (y * z + w * x);
}

// poses (and, for the solve kernels, the dyn record) published by scene_kindyn_kernel for the current state -> shared memory
template <typename T, typename S>
__device__ __forceinline__ void load_poses(const PipeBuf<T> &pb, S &s, int env, int lane) {
  const T *gx = pb.xpos + (size_t)env * (NSLOT * 3), *gm = pb.xmat + (size_t)env * (NSLOT * 9);
  for (int i = lane; i < NSLOT * 3; i += 32) (&s.xpos[0][0])[i] = gx[i];
  for (int i = lane; i < NSLOT * 9; i += 32) (&s.xmat[0][0])[i] = gm[i];
}
template <typename T, typename S>
__device__ __forceinline__ void load_dyn(const PipeBuf<T> &pb, S &s, int env, int lane) {
  const T *gd = pb.dyn + (size_t)env * DYNW;
  static_assert(sizeof(ArmRows<T>) == 24 * sizeof(T), "ArmRows layout");
  for (int i = lane; i < 3 * NA; i += 32) { (&s.arm_p[0][0])[i] = gd[DYN_P + i]; (&s.arm_a[0][0])[i] = gd[DYN_A + i]; }
  if (lane < NV) s.qacc_s[lane] = gd[DYN_QACC + lane];
  for (int i = lane; i < 21 * NARM; i += 32) (&s.Marm[0][0])[i] = gd[DYN_MARM + i];
  for (int i = lane; i < 21 * NPROP; i += 32) (&s.Mprop[0][0])[i] = gd[DYN_MPROP + i];
  for (int i = lane; i < 24 * NARM; i += 32) reinterpret_cast<T *>(&s.arows[0])[i] = gd[DYN_ROWS + i];
}

// free-joint mass block (packed lower 6x6) and bias force for prop p (uniform; [upstream] mj_crb / mj_rne for a free body)
template <typename T>
__device__ __forceinline__ void prop_dynamics(const SceneModel<T> &sm, const ArmModelT<T> &am, const T *R, const T *qd, int p, T (&M)[21], T (&bias)[6]) {
  const T m = sm.prop_mass[p];
  const T ip[3] = {sm.prop_ipos[p][0], sm.prop_ipos[p][1], sm.prop_ipos[p][2]};
  // M_tt = m I ; M_rt = (-m R [ipos]x)^T ; M_rr = I_origin (body axes, constant)
#pragma unroll
  for (int i = 0; i < 21; i++) M[i] = T(0);
  M[tri(0, 0)] = m; M[tri(1, 1)] = m; M[tri(2, 2)] = m;
  // K = -m R [ip]x  (3x3, rows = translational dof, cols = rotational dof)
  const T sk[9] = {T(0), -ip[2], ip[1], ip[2], T(0), -ip[0], -ip[1], ip[0], T(0)};
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      T v = T(0);
#pragma unroll
      for (int e = 0; e < 3; e++) v += R[3 * r + e] * sk[3 * e + c];
      M[tri(3 + c, r)] = -m * v;
    }
  const T *Io = sm.prop_Iorg[p];
  M[tri(3, 3)] = Io[0]; M[tri(4, 4)] = Io[1]; M[tri(5, 5)] = Io[2]; M[tri(4, 3)] = Io[3]; M[tri(5, 3)] = Io[4]; M[tri(5, 4)] = Io[5];
  // bias: f = m (w x (w x c) - g), tau_com = w x (Ic w)  (world); generalized: [f ; R^T (tau + c x f)]
  const T wl[3] = {qd[3], qd[4], qd[5]};
  T w[3], c[3], t1[3], t2[3], f[3];
  mulmv(w, R, wl); mulmv(c, R, ip);
  cross3(t1, w, c); cross3(t2, w, t1);
#pragma unroll
  for (int e = 0; e < 3; e++) f[e] = m * (t2[e] - am.gravity[e]);
  const T *Ic = sm.prop_Icom[p];
  const T Iw[3] = {Ic[0] * wl[0] + Ic[3] * wl[1] + Ic[4] * wl[2], Ic[3] * wl[0] + Ic[1] * wl[1] + Ic[5] * wl[2], Ic[4] * wl[0] + Ic[5] * wl[1] + Ic[2] * wl[2]};
  T tl[3], cf[3], cfl[3];
  cross3(tl, wl, Iw);  // body axes
  cross3(cf, c, f);
  mulmtv(cfl, R, cf);
  bias[0] = f[0]; bias[1] = f[1]; bias[2] = f[2];
  bias[3] = tl[0] + cfl[0]; bias[4] = tl[1] + cfl[1]; bias[5] = tl[2] + cfl[2];
}

template <typename T>
__device__ __forceinline__ bool sphere_vs_obb(const T *c, T r, const T *bpos, const T *bmat, const T *half) {
  T t[3], l[3];
  sub3(t, c, bpos); mulmtv(l, bmat, t);
  T d2 = T(0);
#pragma unroll
  for (int k = 0; k < 3; k++) { const T e = t_abs(l[k]) - half[k]; if (e > T(0)) d2 += e * e; }
  return d2 <= r * r;
}

// bounding box half sizes of a primitive in its own frame (hulls use their sphere)
template <typename T>
__device__ __forceinline__ void bound_half(int type, const T *size, T rbound, T *half) {
  if (type == G_BOX) { half[0] = size[0]; half[1] = size[1]; half[2] = size[2]; }
  else if (type == G_CYLINDER) { half[0] = half[1] = size[0]; half[2] = size[1]; }
  else if (type == G_CAPSULE) { half[0] = half[1] = size[0]; half[2] = size[0] + size[1]; }
  else { half[0] = half[1] = half[2] = rbound; }
}

// geom pose pieces needed by the mid phase, computed per lane
template <typename T, typename S>
__device__ __forceinline__ void geom_pose(const SceneModel<T> &sm, const S &s, int g, T *pos, T *mat) {
  const int slot = sm.geom_slot[g];
  if (slot < 0) {
    for (int c = 0; c < 3; c++) pos[c] = sm.geom_pos[3 * g + c];
    for (int c = 0; c < 9; c++) mat[c] = sm.geom_mat[9 * g + c];
  } else {
    const T *X = s.xpos[slot], *R = s.xmat[slot];
    const T gp[3] = {sm.geom_pos[3 * g], sm.geom_pos[3 * g + 1], sm.geom_pos[3 * g + 2]};
    T t[3];
    mulmv(t, R, gp);
    for (int c = 0; c < 3; c++) pos[c] = X[c] + t[c];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        T v = T(0);
        for (int k = 0; k < 3; k++) v += R[3 * i + k] * sm.geom_mat[9 * g + 3 * k + j];
        mat[3 * i + j] = v;
      }
  }
}

// oriented-box overlap, 15-axis separating-axis test.  (pa, Ra, ha): centre, rotation (columns = box axes), half sizes.
template <typename T>
__device__ __forceinline__ bool obb_overlap(const T *pa, const T *Ra, const T *ha, const T *pb, const T *Rb, const T *hb) {
  T R[9], AR[9], d[3], t[3];
  sub3(d, pb, pa); mulmtv(t, Ra, d);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const T v = Ra[i] * Rb[j] + Ra[3 + i] * Rb[3 + j] + Ra[6 + i] * Rb[6 + j];
      R[3 * i + j] = v; AR[3 * i + j] = t_abs(v) + T(1e-6);
    }
#pragma unroll
  for (int i = 0; i < 3; i++)
    if (t_abs(t[i]) > ha[i] + hb[0] * AR[3 * i] + hb[1] * AR[3 * i + 1] + hb[2] * AR[3 * i + 2]) return false;
#pragma unroll
  for (int j = 0; j < 3; j++)
    if (t_abs(t[0] * R[j] + t[1] * R[3 + j] + t[2] * R[6 + j]) > ha[0] * AR[j] + ha[1] * AR[3 + j] + ha[2] * AR[6 + j] + hb[j]) return false;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const T ra = ha[i1] * AR[3 * i2 + j] + ha[i2] * AR[3 * i1 + j], rb = hb[j1] * AR[3 * i + j2] + hb[j2] * AR[3 * i + j1];
      if (t_abs(t[i2] * R[3 * i1 + j] - t[i1] * R[3 * i2 + j]) > ra + rb) return false;
    }
  }
  return true;
}

// world oriented bounding box of geom g (local AABB of the geom in its own frame, inflated by 0.1 mm)
template <typename T, typename S>
__device__ __forceinline__ void geom_obb(const SceneModel<T> &sm, const S &s, int g, T *pos, T *mat, T *half) {
  T gp[3];
  geom_pose(sm, s, g, gp, mat);
  const T c[3] = {sm.geom_aabb[6 * g], sm.geom_aabb[6 * g + 1], sm.geom_aabb[6 * g + 2]};
  T t[3];
  mulmv(t, mat, c);
#pragma unroll
  for (int k = 0; k < 3; k++) { pos[k] = gp[k] + t[k]; half[k] = sm.geom_aabb[6 * g + 3 + k] + T(1e-4); }
}

// Broad + mid phase of one env ([upstream] mj_collision before the narrow phase): body-pair bounding spheres, geom
// bounding spheres, then oriented boxes (geom AABBs in the geom frame, as MuJoCo's mid phase uses).  The surviving
// geom pairs are appended, in the oracle's pair order, to the device-wide work list of substep `sub`.
template <typename T, typename S>
__device__ __noinline__ void scene_broadphase(const SceneModel<T> &sm, S &s, const PipeBuf<T> &pb, int env, int sub, int &dropped, int lane) {
  BroadScratch<T> &cs = s.broad;
  PROF_START(s);
  // world bounding spheres and oriented boxes of all geoms (lane per geom)
#pragma unroll 1
  for (int g = lane; g < sm.ngeom; g += 32) {
    const int slot = sm.geom_slot[g];
    const T bc[3] = {sm.geom_bcenter[3 * g], sm.geom_bcenter[3 * g + 1], sm.geom_bcenter[3 * g + 2]};
    if (slot < 0) { cs.gcenter[0][g] = bc[0]; cs.gcenter[1][g] = bc[1]; cs.gcenter[2][g] = bc[2]; }
    else {
      T t[3];
      mulmv(t, s.xmat[slot], bc);
      for (int c = 0; c < 3; c++) cs.gcenter[c][g] = s.xpos[slot][c] + t[c];
    }
    T pos[3], mat[9], half[3];
    geom_obb(sm, s, g, pos, mat, half);
#pragma unroll
    for (int c = 0; c < 3; c++) cs.opos[c][g] = pos[c];
#pragma unroll
    for (int c = 0; c < 9; c++) cs.omat[c][g] = mat[c];
  }
  __syncwarp();
  // phase 1: body-pair spheres (lane per body pair), then geom-pair spheres over the static pair list of the surviving body
  // pairs -> candidate list (order kept: broad-phase order == oracle order)
  for (int bdy = lane; bdy < sm.nbody; bdy += 32) {  // world bounding-sphere centres of the bodies
    const int sl = sm.body_slot[bdy];
    const T bc[3] = {sm.body_bcenter[3 * bdy], sm.body_bcenter[3 * bdy + 1], sm.body_bcenter[3 * bdy + 2]};
    T c[3] = {bc[0], bc[1], bc[2]};
    if (sl >= 0) { T t[3]; mulmv(t, s.xmat[sl], bc); for (int e = 0; e < 3; e++) c[e] = s.xpos[sl][e] + t[e]; }
    cs.bcenter[0][bdy] = c[0]; cs.bcenter[1][bdy] = c[1]; cs.bcenter[2][bdy] = c[2];
  }
  __syncwarp();
  int ncand = 0;
#pragma unroll 1
  for (int p0 = 0; p0 < sm.npair; p0 += 32) {
    const int p = p0 + lane;
    bool live = false;
    if (p < sm.npair) {
      const int b1 = sm.bodypair[2 * p], b2 = sm.bodypair[2 * p + 1];
      live = true;
      if (b1 != 0) {  // (the world body holds only the floor plane: always descend)
        const T t[3] = {cs.bcenter[0][b1] - cs.bcenter[0][b2], cs.bcenter[1][b1] - cs.bcenter[1][b2], cs.bcenter[2][b1] - cs.bcenter[2][b2]};
        const T r = sm.body_rbound[b1] + sm.body_rbound[b2];
        live = !(dot3(t, t) > r * r);
      }
    }
    unsigned livemask = __ballot_sync(FULL, live);
#pragma unroll 1
    while (livemask) {
      const int pp = p0 + __ffs(livemask) - 1;
      livemask &= livemask - 1;
      const int k0 = sm.pair_start[pp], k1 = sm.pair_start[pp + 1];
#pragma unroll 1
      for (int base = k0; base < k1; base += 32) {
        const int k = base + lane;
        bool keep = false;
        unsigned gp = 0;
        if (k < k1) {
          gp = sm.geompair[k];
          const int g1 = (int)(gp & 0xff), g2 = (int)(gp >> 8);
          const T cB[3] = {cs.gcenter[0][g2], cs.gcenter[1][g2], cs.gcenter[2][g2]};
          const T rB = sm.geom_rbound[g2];
          if (sm.geom_type[g1] == G_PLANE) {
            const T n[3] = {sm.geom_mat[9 * g1 + 2], sm.geom_mat[9 * g1 + 5], sm.geom_mat[9 * g1 + 8]};
            const T pp3[3] = {sm.geom_pos[3 * g1], sm.geom_pos[3 * g1 + 1], sm.geom_pos[3 * g1 + 2]};
            keep = !(dot3(n, cB) - dot3(n, pp3) - rB > T(0));
          } else {
            const T t[3] = {cs.gcenter[0][g1] - cB[0], cs.gcenter[1][g1] - cB[1], cs.gcenter[2][g1] - cB[2]};
            const T r = sm.geom_rbound[g1] + rB;
            keep = !(dot3(t, t) > r * r);
          }
        }
        const unsigned m = __ballot_sync(FULL, keep);
        const int idx = ncand + __popc(m & ((1u << lane) - 1));
        if (keep && idx < CANDCAP) cs.cand[idx] = gp;
        ncand += __popc(m);
      }
    }
  }
  if (ncand > CANDCAP) { dropped += ncand - CANDCAP; if (lane == 0) DROPCAT(1, ncand - CANDCAP); ncand = CANDCAP; }
  __syncwarp();
  // phase 2: oriented boxes on the compacted candidates (lane per candidate)
  int npq = 0;
#pragma unroll 1
  for (int base = 0; base < ncand; base += 32) {
    const int k = base + lane;
    bool keep = false;
    int g1 = 0, g2 = 0;
    if (k < ncand) {
      const unsigned cd = cs.cand[k];
      g1 = (int)(cd & 0xff); g2 = (int)(cd >> 8);
      T p2[3], m2[9], h2[3];
#pragma unroll
      for (int c = 0; c < 3; c++) { p2[c] = cs.opos[c][g2]; h2[c] = sm.geom_aabb[6 * g2 + 3 + c] + T(1e-4); }
#pragma unroll
      for (int c = 0; c < 9; c++) m2[c] = cs.omat[c][g2];
      if (sm.geom_type[g1] == G_PLANE) {
        const T n[3] = {sm.geom_mat[9 * g1 + 2], sm.geom_mat[9 * g1 + 5], sm.geom_mat[9 * g1 + 8]};
        const T pp[3] = {sm.geom_pos[3 * g1], sm.geom_pos[3 * g1 + 1], sm.geom_pos[3 * g1 + 2]};
        const T ext = t_abs(n[0] * m2[0] + n[1] * m2[3] + n[2] * m2[6]) * h2[0] + t_abs(n[0] * m2[1] + n[1] * m2[4] + n[2] * m2[7]) * h2[1] +
                      t_abs(n[0] * m2[2] + n[1] * m2[5] + n[2] * m2[8]) * h2[2];
        keep = !(dot3(n, p2) - dot3(n, pp) - ext > T(0));
      } else {
        T p1[3], m1[9], h1[3];
#pragma unroll
        for (int c = 0; c < 3; c++) { p1[c] = cs.opos[c][g1]; h1[c] = sm.geom_aabb[6 * g1 + 3 + c] + T(1e-4); }
#pragma unroll
        for (int c = 0; c < 9; c++) m1[c] = cs.omat[c][g1];
        keep = obb_overlap(p1, m1, h1, p2, m2, h2);
      }
    }
    const unsigned m = __ballot_sync(FULL, keep);
    const int idx = npq + __popc(m & ((1u << lane) - 1));
    if (keep && idx < PAIRCAP) cs.pairq[idx] = (unsigned)g1 | ((unsigned)g2 << 8) | ((unsigned)idx << 16);
    npq += __popc(m);
  }
  if (npq > PAIRCAP) { dropped += npq - PAIRCAP; if (lane == 0) DROPCAT(2, npq - PAIRCAP); npq = PAIRCAP; }
  __syncwarp();
  // append (env, g1 | g2 << 8 | pair index << 16) to the work queue of its second geom - or of its first one when only that
  // one is a hull (arm link vs table): the queue key is the geom whose vertex data the pair's threads will stream
  int qdrop = 0;
  for (int i = lane; i < npq; i += 32) {
    const unsigned pq = cs.pairq[i];
    const int g1 = (int)(pq & 0xff), g2 = (int)((pq >> 8) & 0xff);
    const int q = (sm.geom_type[g2] != G_HULL && sm.geom_type[g1] == G_HULL) ? g1 : g2;
    const int slot = atomicAdd(pb.nwork + WSTRIDE * sub + q, 1);
    if (slot < pb.work_cap) pb.work[(size_t)q * pb.work_cap + slot] = make_uint2((unsigned)env, pq);
    else { qdrop++; DROPCAT(3, 1); }
  }
  dropped += warp_sum(qdrop);
  PROF_ACC(s, P_BROAD, lane);
  PROF_CNT(s, P_NPQ, npq, lane);
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------ reward (so100_hand_over.py:238-275)
template <typename T>
__device__ __forceinline__ void mat2quat(const T *m, T *q) {  // mju_mat2Quat
  if (m[0] + m[4] + m[8] > T(0)) {
    q[0] = T(0.5) * t_sqrt(T(1) + m[0] + m[4] + m[8]);
    q[1] = T(0.25) * (m[7] - m[5]) / q[0]; q[2] = T(0.25) * (m[2] - m[6]) / q[0]; q[3] = T(0.25) * (m[3] - m[1]) / q[0];
  } else if (m[0] > m[4] && m[0] > m[8]) {
    q[1] = T(0.5) * t_sqrt(T(1) + m[0] - m[4] - m[8]);
    q[0] = T(0.25) * (m[7] - m[5]) / q[1]; q[2] = T(0.25) * (m[1] + m[3]) / q[1]; q[3] = T(0.25) * (m[2] + m[6]) / q[1];
  } else if (m[4] > m[8]) {
    q[2] = T(0.5) * t_sqrt(T(1) - m[0] + m[4] - m[8]);
    q[0] = T(0.25) * (m[2] - m[6]) / q[2]; q[1] = T(0.25) * (m[1] + m[3]) / q[2]; q[3] = T(0.25) * (m[5] + m[7]) / q[2];
  } else {
    q[3] = T(0.5) * t_sqrt(T(1) - m[0] - m[4] + m[8]);
    q[0] = T(0.25) * (m[3] - m[1]) / q[3]; q[1] = T(0.25) * (m[2] + m[6]) / q[3]; q[2] = T(0.25) * (m[5] + m[7]) / q[3];
  }
  const T n = t_sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) q[i] /= n;
}
template <typename T>
__device__ __forceinline__ void quat_mul(T *r, const T *a, const T *b) {
  const T t0 = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], t1 = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  const T t2 = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], t3 = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = t0; r[1] = t1; r[2] = t2; r[3] = t3;
}
template <typename T>
__device__ __forceinline__ void quat2mat(const T *q, T *m) {
  const T w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[4] = w * w - x * x + y * y - z * z; m[8] = w * w - x * x - y * y + z * z;
  m[1] = T(2) * (x * y - w * z); m[2] = T(2) * (x * z + w * y); m[3] = T(2) * (x * y + w * z);
  m[5] = T(2) * (y * z - w * x); m[6] = T(2) * (x * z - w * y); m[7] = T(2) * (y * z + w * x);
}
// oobb_utils.py:202-273 — 6-axis SAT on the projected corners, strict comparisons
template <typename T>
__device__ bool overlap_oobb_oobb(const T *p0, const T *q0, const T *h0, const T *p1, const T *q1, const T *h1) {
  const T inv[4] = {q0[0], -q0[1], -q0[2], -q0[3]};
  T dp[3], rp[3], rq[4], Rm[9], Ri[9];
  sub3(dp, p1, p0);
  quat2mat(inv, Ri); mulmv(rp, Ri, dp);
  quat_mul(rq, inv, q1);
  quat2mat(rq, Rm);
  for (int a = 0; a < 6; a++) {
    T ax[3];
    if (a < 3) { ax[0] = a == 0; ax[1] = a == 1; ax[2] = a == 2; }
    else { const T e[3] = {T(a == 3), T(a == 4), T(a == 5)}; mulmv(ax, Rm, e); }
    T amax = -INFINITY, amin = INFINITY, bmax = -INFINITY, bmin = INFINITY;
    for (int i = 0; i < 8; i++) {
      const int iz = i / 4, ixy = i % 4;
      const T t[3] = {T(ixy % 2), T(ixy / 2), T(iz)};
      T va[3], l[3], vb[3];
      for (int c = 0; c < 3; c++) { va[c] = -h0[c] * (T(1) - t[c]) + h0[c] * t[c]; l[c] = -h1[c] * (T(1) - t[c]) + h1[c] * t[c]; }
      mulmv(vb, Rm, l);
      for (int c = 0; c < 3; c++) vb[c] += rp[c];
      const T pa = dot3(va, ax), pb = dot3(vb, ax);
      amax = max(amax, pa); amin = min(amin, pa); bmax = max(bmax, pb); bmin = min(bmin, pb);
    }
    if (amax < bmin || amin > bmax) return false;
  }
  return true;
}
template <typename T, typename S>
__device__ float scene_reward(const SceneModel<T> &sm, const S &s) {
  // success_detector_utils.py:22-28 — linear velocity of either prop >= 1e-3 -> 0
  for (int p = 0; p < NPROP; p++) {
    T mx = T(0);
    for (int c = 0; c < 3; c++) mx = max(mx, t_abs(s.qd[NA + 6 * p + c]));
    if (mx >= T(1e-3)) return 0.f;
  }
  // object OOBB: root BVH box at xipos / ximat (oobb_utils.py:137-148,165-172)
  const T *R0 = s.xmat[NA], *X0 = s.xpos[NA];
  T ximat[9], xipos[3], t[3], q0[4], p0[3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { T v = T(0); for (int k = 0; k < 3; k++) v += R0[3 * i + k] * sm.prop_Riq[0][3 * k + j]; ximat[3 * i + j] = v; }
  mulmv(t, R0, sm.prop_ipos[0]);
  for (int c = 0; c < 3; c++) xipos[c] = X0[c] + t[c];
  mat2quat(ximat, q0);
  T Rq[9];
  quat2mat(q0, Rq);
  mulmv(t, Rq, sm.reward_obj_box);
  for (int c = 0; c < 3; c++) p0[c] = t[c] + xipos[c];
  // container box (oobb_utils.py:175-199) — xquat of the bowl body = normalised qpos quaternion
  const T *qb = s.q + NA + 7 + 3;
  T q1[4] = {qb[0], qb[1], qb[2], qb[3]}, p1[3];
  const T n = t_sqrt(q1[0] * q1[0] + q1[1] * q1[1] + q1[2] * q1[2] + q1[3] * q1[3]);
  for (int i = 0; i < 4; i++) q1[i] /= n;
  for (int k = 0; k < sm.nreward_box; k++) {  // every overlap box must be touched by the object's box (so100_hand_over.py:263-273)
    mulmv(t, s.xmat[NA + 1], sm.reward_box_pos[k]);
    for (int c = 0; c < 3; c++) p1[c] = t[c] + s.xpos[NA + 1][c];
    if (!overlap_oobb_oobb(p0, q0, sm.reward_obj_box + 3, p1, q1, sm.reward_box_half[k])) return 0.f;
  }
  return 1.f;
}

// ------------------------------------------------------------------------------------------------ task layer: observations
template <typename T, typename SC>
__device__ void write_obs_scene(const StepCfg &cfg, const EnvState<T> &S, const so101_step_out &out, const SC &s, int env, int t, float reward,
                                float discount, uint8_t st, int lane) {
  constexpr int SD = NQ + NV;
  const int dj = cfg.dj + 1, dp = cfg.dp + 1;
  const size_t N = S.N;
  float *rj = S.ring_joints + ((size_t)(t % dj) * N + env) * NA;
  float *rp = S.ring_phys + ((size_t)(t % dp) * N + env) * SD;
  const int tj = t - cfg.dj > 0 ? t - cfg.dj : 0, tp = t - cfg.dp > 0 ? t - cfg.dp : 0;
  const float *sj = S.ring_joints + ((size_t)(tj % dj) * N + env) * NA;
  const float *sp = S.ring_phys + ((size_t)(tp % dp) * N + env) * SD;
  for (int i = lane; i < SD; i += 32) {
    const float v = i < NQ ? (float)s.q[i] : (float)s.qd[i - NQ];
    rp[i] = v;
    if (out.physics_state) out.physics_state[(size_t)env * SD + i] = v;
    if (out.delayed_physics_state) out.delayed_physics_state[(size_t)env * SD + i] = tp == t ? v : sp[i];
  }
  if (lane < NA) {
    const float v = (float)s.q[lane];
    rj[lane] = v;
    if (out.undelayed_joints_pos) out.undelayed_joints_pos[(size_t)env * NA + lane] = v;
    if (out.joints_pos) out.joints_pos[(size_t)env * NA + lane] = tj == t ? v : sj[lane];
    if (out.commanded_joints_pos) out.commanded_joints_pos[(size_t)env * NA + lane] = (float)s.ctrl[lane];
  }
  if (lane == 0) {
    if (out.reward) out.reward[env] = reward;
    if (out.discount) out.discount[env] = discount;
    if (out.step_type) out.step_type[env] = st;
  }
}

// ------------------------------------------------------------------------------------------------ placements (initialize_episode)
// Philox4x32-10 (Salmon et al. 2011): counter-based, so a placement is a pure function of (seed, env, episode draw, attempt)
__device__ __forceinline__ void philox4x32_10(unsigned long long key, unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned (&out)[4]) {
  unsigned k0 = (unsigned)key, k1 = (unsigned)(key >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0, hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u01(unsigned x) { return (double)(x >> 8) * (1.0 / 16777216.0); }  // [0, 1)

// pose of prop p for (env, episode draw, attempt): position ~ U(lo, hi), rotation about z by U(yaw)  (so100_hand_over.py:37-55)
template <typename T>
__device__ __forceinline__ void sample_prop_pose(const EnvState<T> &S, int env, int p, unsigned draw, unsigned attempt, TS *qp) {
  unsigned r[4];
  philox4x32_10(S.place.seed, (unsigned)env, draw, attempt, (unsigned)p, r);
  for (int c = 0; c < 3; c++) qp[c] = (TS)S.place.lo[p][c] + u01(r[c]) * ((TS)S.place.hi[p][c] - (TS)S.place.lo[p][c]);
  const TS yaw = (TS)S.place.yaw[p][0] + u01(r[3]) * ((TS)S.place.yaw[p][1] - (TS)S.place.yaw[p][0]);
  TS sn, cn;
  sincos(TS(0.5) * yaw, &sn, &cn);
  qp[3] = cn; qp[4] = TS(0); qp[5] = TS(0); qp[6] = sn;
}

// SETTLE mode, once per control step (scene_begin_kernel): draw a placement if one is due and keep the env stepping with the
// home command.  Returns false if the env idles this step (settled and waiting).
template <typename T>
__device__ bool settle_begin(const StepCfg &cfg, const EnvState<T> &S, const PipeBuf<T> &pb, int env, int lane) {
  const int st = S.sstate[env];
  __syncwarp();
  if (st == SETTLE_DONE || (env >= S.NU && !S.use_ring)) { if (lane == 0) pb.active[env] = 0; return false; }
  if (st == SETTLE_SAMPLE) {
    // arm qpos = 0 (the reference never applies its home pose, so100_task.py:308-313), velocities 0, props at their sampled poses;
    // a rejected container placement (attempt > 0) re-draws the container only (the second PropPlacer, so100_hand_over.py:216-221)
    const unsigned draw = S.draws[env], att = (unsigned)S.attempt[env];
    if (lane < NPROP) sample_prop_pose(S, env, lane, draw, S.place.check_collisions[lane] ? att : 0u, S.qpos + (size_t)env * NQ + NA + 7 * lane);
    if (lane < NA) S.qpos[(size_t)env * NQ + lane] = TS(0);
    if (lane < NV) { S.qvel[(size_t)env * NV + lane] = TS(0); S.warm[(size_t)env * NV + lane] = T(0); }
    if (lane == 0) { S.settle_sub[env] = 0; S.sstate[env] = SETTLE_RUN; }
  }
  if (lane < NA) S.ctrl[(size_t)env * NA + lane] = (T)cfg.home[lane % NJ] + (T)cfg.offsets[lane % NJ];   // so100_task.py:316-317
  if (lane == 0) { pb.active[env] = 1; pb.flags[env] = 0; }
  return true;
}

template <typename T, typename SC>
__device__ void reset_env_scene(const StepCfg &cfg, const EnvState<T> &S, const so101_step_out &out, SC &s, int env, int lane) {
  // initialize_episode (so100_hand_over.py:320-323): each episode starts from the next entry of the env's pool of sampled and
  // settled prop placements
  const int ep = S.episode[env];
  const size_t slot = (size_t)(ep % S.npool) * S.N + env;
  // a fresh settled placement from the nursery's ring if one is available (every placement is consumed once), else the env's
  // own initial state / reset pool entry
  const TS *srcq = S.init_qpos + slot * NQ, *srcv = S.init_qvel + slot * NV;
  if (S.use_ring && ep > 0) {   // (episode 0 starts from the env's own settled state: deterministic for a given seed)
    int t = 0;
    if (lane == 0) {
      t = atomicAdd(S.ring_ctr + RC_TAIL, 1);
      if (t >= *((volatile int *)(S.ring_ctr + RC_CLAIM))) { atomicSub(S.ring_ctr + RC_TAIL, 1); atomicAdd(S.ring_ctr + RC_REUSED, 1); t = -1; }
    }
    t = __shfl_sync(FULL, t, 0);
    if (t >= 0) { srcq = S.ring_q + (size_t)(t % S.ring_cap) * NQ; srcv = S.ring_v + (size_t)(t % S.ring_cap) * NV; }
  }
  __syncwarp();
  for (int i = lane; i < NQ; i += 32) { const TS v = srcq[i]; s.q[i] = (T)v; S.qpos[(size_t)env * NQ + i] = v; }
  for (int i = lane; i < NV; i += 32) { const TS v = srcv[i]; s.qd[i] = (T)v; S.qvel[(size_t)env * NV + i] = v; S.warm[(size_t)env * NV + i] = T(0); }
  if (lane < NA) { s.ctrl[lane] = (T)cfg.home[lane % NJ] + (T)cfg.offsets[lane % NJ]; S.ctrl[(size_t)env * NA + lane] = s.ctrl[lane]; }
  if (lane == 0) { S.step[env] = 0; S.needs_reset[env] = 0; S.episode[env] = ep + 1; }
  __syncwarp();
  write_obs_scene(cfg, S, out, s, env, 0, 0.f, 1.f, SO101_STEP_FIRST, lane);
}

// SETTLE mode, end of a control step (task-layer launch; one lane): a diverged settle starts over; a settled NURSERY env
// publishes its state into the ring (if there is room) and goes back to sampling; a settled user env waits for
// so101_sample_and_settle to finish.  Publishing here and consuming in scene_begin_kernel keeps producers and consumers in
// different kernels.
template <typename T>
__device__ void settle_end_of_step(const EnvState<T> &S, int env, bool diverged) {
  if (diverged) { S.sstate[env] = SETTLE_SAMPLE; S.attempt[env] = 0; S.draws[env] += 1; return; }
  if (S.sstate[env] != SETTLE_DONE) { if (env < S.NU) atomicAdd(S.ring_ctr + RC_PENDING, 1); return; }
  if (env < S.NU || !S.use_ring) return;
  const int c = atomicAdd(S.ring_ctr + RC_CLAIM, 1);
  if (c - *((volatile int *)(S.ring_ctr + RC_TAIL)) >= S.ring_cap) { atomicSub(S.ring_ctr + RC_CLAIM, 1); return; }  // ring full: try again next step
  TS *dq = S.ring_q + (size_t)(c % S.ring_cap) * NQ, *dv = S.ring_v + (size_t)(c % S.ring_cap) * NV;
  for (int i = 0; i < NQ; i++) dq[i] = i < NA ? TS(0) : S.qpos[(size_t)env * NQ + i];
  for (int i = 0; i < NV; i++) dv[i] = i < NA ? TS(0) : S.qvel[(size_t)env * NV + i];
  S.sstate[env] = SETTLE_SAMPLE; S.attempt[env] = 0; S.draws[env] += 1;
}

// ------------------------------------------------------------------------------------------------ kernels
template <typename SC, typename T>
__device__ __forceinline__ void prof_begin(SC &s, const EnvState<T> &S, int lane) {
  if (lane == 0) {
    s.profon = S.prof != nullptr;
    for (int i = 0; i < 16; i++) s.prof[i] = 0;
  }
}
template <typename SC, typename T>
__device__ __forceinline__ void prof_flush(SC &s, const EnvState<T> &S, int lane) {
  if (s.profon && lane == 0)
    for (int i = 0; i < 16; i++)
      if (s.prof[i]) atomicAdd(S.prof + i, (unsigned long long)s.prof[i]);
}

// envs (warps) per CTA in the begin / broad-phase / task-layer kernels (static shared memory: 48 KB per CTA)
template <typename T> struct BroadCfg { static constexpr int WARPS = sizeof(T) == 8 ? 2 : 4; };
#define WARPS_BROAD (BroadCfg<T>::WARPS)
constexpr int KD_THREADS = 32;   // envs (threads) per CTA in the kinematics + smooth-dynamics kernel: one warp, one staging tile
// per-env record the kernel produces: poses of the dynamic bodies, then the dyn record; staged per warp in shared memory with an
// odd row stride (conflict-free thread-per-row writes) and written out as ONE contiguous, coalesced block per array
constexpr int KD_POSE = NSLOT * 12, KD_ROW = KD_POSE + DYNW, KD_STRIDE = KD_ROW | 1;

// Once per control step: dm_control auto-reset (writes the FIRST TimeStep), action + calibration -> ctrl.
template <typename T>
__global__ void __launch_bounds__(WARPS_BROAD * 32) scene_begin_kernel(const __grid_constant__ StepCfg cfg, const EnvState<T> S, const PipeBuf<T> pb,
                                                                      const float *__restrict__ action, const so101_step_out out) {
  __shared__ BroadEnv<T> all[WARPS_BROAD];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int env = pb.env0 + blockIdx.x * WARPS_BROAD + wib;
  if (env >= pb.env0 + pb.nenv) return;
  if (lane == 0) pb.ncon_raw[env] = 0;
  if (S.mode[env]) { settle_begin(cfg, S, pb, env, lane); return; }
  if (S.needs_reset[env]) {  // the step() after a LAST step resets and returns FIRST; no physics this call
    reset_env_scene(cfg, S, out, all[wib], env, lane);
    if (lane == 0) pb.active[env] = 0;
    return;
  }
  if (lane < NA) S.ctrl[(size_t)env * NA + lane] = (T)action[(size_t)env * NA + lane] + (T)cfg.offsets[lane % NJ];  // so100_task.py:266-287
  if (lane == 0) { pb.active[env] = 1; pb.flags[env] = 0; }
}

// ONE THREAD per env: forward kinematics of the arm and the props at the current qpos (poses for the collision kernels and
// the constraint Jacobians) and - unless this is the refresh after the last substep - the smooth dynamics the solve kernels
// start from ([upstream] mj_kinematics, mj_comPos, mj_crb, mj_factorM, mj_rne, mj_fwdActuation, mj_fwdAcceleration and the
// friction-loss / limit rows of mj_makeConstraint).  All of it is straight-line scalar code on registers, the same functions
// the arm-only kernel uses (arm_dynamics.cuh, arm_solver.cuh).
template <typename T>
__global__ void __launch_bounds__(KD_THREADS) scene_kindyn_kernel(const __grid_constant__ ArmSetT<T> am, const __grid_constant__ ArmSetT<double> am64,
                                                                 const __grid_constant__ SceneModel<T> sm, const EnvState<T> S, const PipeBuf<T> pb,
                                                                 int need_dyn) {
  extern __shared__ __align__(16) unsigned char kd_raw[];
  T *tile = reinterpret_cast<T *>(kd_raw);                      // [32 envs][KD_STRIDE]
  const int lane = threadIdx.x;
  const int env0 = pb.env0 + blockIdx.x * KD_THREADS, env = env0 + lane;
  const bool live = env < pb.env0 + pb.nenv && pb.active[env];
  const unsigned livemask = __ballot_sync(0xffffffffu, live);
  if (!livemask) return;
  const bool dyn = live && need_dyn && !pb.flags[env];  // (a diverged env is frozen for the rest of the control step)
  const unsigned dynmask = __ballot_sync(0xffffffffu, dyn);
  // thread-per-env results go to the env's row of the tile (the pointers keep the names of the global arrays they stand for)
  T *gx = tile + (size_t)lane * KD_STRIDE, *gm = gx + NSLOT * 3, *gd = gx + KD_POSE;
  if (live) {
  const TS *gq = S.qpos + (size_t)env * NQ, *gv = S.qvel + (size_t)env * NV;
#pragma unroll 1
  for (int p = 0; p < NPROP; p++) {
    const TS *qp = gq + NA + 7 * p;
    const T quat[4] = {(T)qp[3], (T)qp[4], (T)qp[5], (T)qp[6]};
    T Rp[9];
    prop_rotation(quat, Rp);
#pragma unroll
    for (int c = 0; c < 3; c++) gx[3 * (NA + p) + c] = (T)qp[c];
#pragma unroll
    for (int e = 0; e < 9; e++) gm[9 * (NA + p) + e] = Rp[e];
    if (dyn) {
      T qdp[6], M[21], bias[6], x[6];
#pragma unroll
      for (int i = 0; i < 6; i++) qdp[i] = (T)gv[NA + 6 * p + i];
      prop_dynamics(sm, am[0], Rp, qdp, p, M, bias);
#pragma unroll
      for (int i = 0; i < 21; i++) gd[DYN_MPROP + 21 * p + i] = M[i];
#pragma unroll
      for (int i = 0; i < 6; i++) x[i] = -bias[i];
      chol6(M);
      chol6_solve(M, x);
#pragma unroll
      for (int i = 0; i < 6; i++) gd[DYN_QACC + NA + 6 * p + i] = x[i];
    }
  }
  // each arm's smooth dynamics run in float64 on the float64 state in both precisions (env_state.cuh); what the float32
  // collision and solve kernels read (poses, mass matrix, qacc_smooth, rows) is rounded when it is stored
#pragma unroll 1
  for (int arm = 0; arm < NARM; arm++) {
    const int o = NJ * arm;   // first dof / pose slot / actuator of this arm
    double qa[NJ], qda[NJ];
    T ca[NJ];
#pragma unroll
    for (int i = 0; i < NJ; i++) { qa[i] = gq[o + i]; qda[i] = gv[o + i]; ca[i] = S.ctrl[(size_t)env * NA + o + i]; }
    ArmKin<double> k;
    {
      double R[NJ][9];
      arm_fk<double>(am64[arm], qa, k, R);
#pragma unroll
      for (int i = 0; i < NJ; i++) {
        gx[3 * (o + i)] = (T)k.p[i].x; gx[3 * (o + i) + 1] = (T)k.p[i].y; gx[3 * (o + i) + 2] = (T)k.p[i].z;
        gd[DYN_P + 3 * (o + i)] = (T)k.p[i].x; gd[DYN_P + 3 * (o + i) + 1] = (T)k.p[i].y; gd[DYN_P + 3 * (o + i) + 2] = (T)k.p[i].z;
        gd[DYN_A + 3 * (o + i)] = (T)k.a[i].x; gd[DYN_A + 3 * (o + i) + 1] = (T)k.a[i].y; gd[DYN_A + 3 * (o + i) + 2] = (T)k.a[i].z;
#pragma unroll
        for (int e = 0; e < 9; e++) gm[9 * (o + i) + e] = (T)R[i][e];
      }
    }
    if (!dyn) continue;
    double M[21], bias[NJ], frc[NJ], qs[NJ], L[21];
    arm_crb_rne(am64[arm], k, qda, M, bias);
    arm_actuation_d(am[arm], qa, qda, ca, frc);
    gd[DYN_VELMASK + arm] = (T)arm_actuation_vel_mask(am[arm], frc);
#pragma unroll
    for (int i = 0; i < 21; i++) { L[i] = M[i]; gd[DYN_MARM + 21 * arm + i] = (T)M[i]; }
    chol6(L);
#pragma unroll
    for (int i = 0; i < NJ; i++) qs[i] = frc[i] - bias[i];
    chol6_solve(L, qs);
    T qaT[NJ], qdaT[NJ], qsT[NJ];
#pragma unroll
    for (int i = 0; i < NJ; i++) { qaT[i] = (T)qa[i]; qdaT[i] = (T)qda[i]; qsT[i] = (T)qs[i]; }
    ArmRows<T> arows;
    arm_make_rows(am[arm], qaT, qdaT, qsT, arows);
#pragma unroll
    for (int i = 0; i < NJ; i++) {
      gd[DYN_QACC + o + i] = qsT[i];
      T *gr = gd + DYN_ROWS + 24 * arm;
      gr[i] = arows.jar0_f[i]; gr[6 + i] = arows.jar0_l[i]; gr[12 + i] = arows.D_l[i]; gr[18 + i] = arows.js[i];
    }
  }
  }  // live
  __syncwarp();
  // coalesced write-out: the 32 envs' rows are contiguous in each global array ([N][NSLOT*3], [N][NSLOT*9], [N][DYNW])
  T *ox = pb.xpos + (size_t)env0 * (NSLOT * 3), *om = pb.xmat + (size_t)env0 * (NSLOT * 9), *od = pb.dyn + (size_t)env0 * DYNW;
  for (int i = lane; i < 32 * NSLOT * 3; i += 32) { const int r = i / (NSLOT * 3), k = i - r * (NSLOT * 3); if ((livemask >> r) & 1u) ox[i] = tile[(size_t)r * KD_STRIDE + k]; }
  for (int i = lane; i < 32 * NSLOT * 9; i += 32) { const int r = i / (NSLOT * 9), k = i - r * (NSLOT * 9); if ((livemask >> r) & 1u) om[i] = tile[(size_t)r * KD_STRIDE + NSLOT * 3 + k]; }
  if (dynmask)
    for (int i = lane; i < 32 * DYNW; i += 32) { const int r = i / DYNW, k = i - r * DYNW; if ((dynmask >> r) & 1u) od[i] = tile[(size_t)r * KD_STRIDE + KD_POSE + k]; }
}

// Warp per env, after every kinematics refresh: the broad + mid phase that fills the work queues of substep `sub`, or - after
// the last substep (`last`) - the task layer on the refreshed poses ([upstream] mj_step1 refresh before the observables
// and the reward are read): observation delay rings, SO100HandOver reward, discount, time limit, outputs.
template <typename T>
__global__ void __launch_bounds__(WARPS_BROAD * 32) scene_broad_kernel(const __grid_constant__ SceneModel<T> sm, const __grid_constant__ StepCfg cfg,
                                                                      const EnvState<T> S, const PipeBuf<T> pb, const so101_step_out out, int sub, int last) {
  __shared__ BroadEnv<T> all[WARPS_BROAD];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int env = pb.env0 + blockIdx.x * WARPS_BROAD + wib;
  if (env >= pb.env0 + pb.nenv) return;
  if (!pb.active[env]) {  // (a settled nursery env idles, but keeps trying to publish its placement)
    if (last && S.mode[env] && lane == 0) settle_end_of_step(S, env, false);
    return;
  }
  BroadEnv<T> &s = all[wib];
  prof_begin(s, S, lane);
  load_poses(pb, s, env, lane);
  __syncwarp();
  const bool bad = pb.flags[env] == 1;   // (flag 2: a rejected placement sits out the rest of this control step)
  if (!last) {
    int dropped = 0;
    if (!pb.flags[env]) scene_broadphase(sm, s, pb, env, sub, dropped, lane);
    if (lane == 0 && dropped) { atomicAdd(S.diverged_count + 1, dropped); S.dropped_env[env] += dropped; }
  } else if (S.mode[env]) {
    if (lane == 0) settle_end_of_step(S, env, bad);
  } else {
    for (int i = lane; i < NQ; i += 32) s.q[i] = (T)S.qpos[(size_t)env * NQ + i];
    for (int i = lane; i < NV; i += 32) s.qd[i] = (T)S.qvel[(size_t)env * NV + i];
    if (lane < NA) s.ctrl[lane] = S.ctrl[(size_t)env * NA + lane];
    __syncwarp();
    const int t = S.step[env] + 1;
    float reward = scene_reward(sm, s), discount = 1.f;
    uint8_t st = (cfg.last_step > 0 && t >= cfg.last_step) ? SO101_STEP_LAST : SO101_STEP_MID;
    if (cfg.terminate_on_success && reward >= 1.f) { discount = 0.f; st = SO101_STEP_LAST; }  // so100_task.py:292-302
    if (bad) { reward = 0.f; discount = 0.f; st = SO101_STEP_LAST; }                         // task_suite.py:153
    __syncwarp();
    if (lane == 0) {
      S.step[env] = t; S.needs_reset[env] = st == SO101_STEP_LAST;
      if (bad) atomicAdd(S.diverged_count, 1);
    }
    write_obs_scene(cfg, S, out, s, env, t, reward, discount, st, lane);
  }
  prof_flush(s, S, lane);
}

// queue offsets: exclusive prefix sum of the per-geom item counts (each warp builds its own copy: 4 counts per lane)
__device__ __forceinline__ void build_qpref(int *qpref, const int *cnt, int cap, int lane) {
  int c[4], tot = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) { c[k] = min(cnt[4 * lane + k], cap); tot += c[k]; }
  int incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
  int run = incl - tot;
#pragma unroll
  for (int k = 0; k < 4; k++) { qpref[4 * lane + k] = run; run += c[k]; }
  if (lane == 31) qpref[WQ] = run;
  __syncwarp();
}
__device__ __forceinline__ int queue_of(const int *qpref, int item) {  // largest q with qpref[q] <= item
  int q = 0;
#pragma unroll
  for (int step = WQ / 2; step > 0; step >>= 1)
    if (qpref[q + step] <= item) q += step;
  return q;
}

// Per substep, ONE THREAD per candidate pair: boolean GJK.  Consecutive threads take consecutive items of one queue = the same
// hull for different envs, so the vertex scan is a broadcast stream and the warp runs the simplex logic for 32 pairs at once
// (the warp-per-pair version spent most of its instructions on scalar simplex code executed redundantly by 32 lanes).
// Intersecting pairs go, with their simplex, to the hit slots of their work queue for the EPA / manifold kernel.
constexpr int GJK_THREADS = 128;
template <typename T>
__global__ void __launch_bounds__(GJK_THREADS) scene_gjk_kernel(const __grid_constant__ SceneModel<T> sm, const EnvState<T> S, const PipeBuf<T> pb, int sub) {
  __shared__ int qpref[GJK_THREADS / 32][WQ + 1];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int *cnt = pb.nwork + WSTRIDE * sub;
  build_qpref(qpref[wib], cnt, pb.work_cap, lane);
  const int nwork = qpref[wib][WQ];
  const int stride = gridDim.x * GJK_THREADS;
  long long git_sum = 0;
#pragma unroll 1
  for (int item0 = blockIdx.x * GJK_THREADS + wib * 32; item0 < nwork; item0 += stride) {
    const int item = item0 + lane;
    bool hit = false;
    uint2 w = make_uint2(0u, 0u);
    MPoint<T> Sx[4];
    int n = 0, ha = 0, hb = 0;
    if (item < nwork) {
      const int q = queue_of(qpref[wib], item);
      w = pb.work[(size_t)q * pb.work_cap + (item - qpref[wib][q])];
      const int env = (int)w.x, g1 = (int)(w.y & 0xff), g2 = (int)((w.y >> 8) & 0xff);
      if (sm.geom_type[g1] == G_PLANE) hit = true;  // plane pairs need no simplex
      else {
        const T(*xpos)[3] = reinterpret_cast<const T(*)[3]>(pb.xpos + (size_t)env * (NSLOT * 3));
        const T(*xmat)[9] = reinterpret_cast<const T(*)[9]>(pb.xmat + (size_t)env * (NSLOT * 9));
        Shape<T> A, B;
        make_shape(sm, xpos, xmat, g1, A);
        make_shape(sm, xpos, xmat, g2, B);
        int it = 0;
        hit = gjk_intersect(sm, A, B, Sx, n, it) != 0;
        ha = A.hint; hb = B.hint;
        git_sum += it;
      }
    }
    if (hit) {
      // slot inside the queue's own item range (hits <= items of the queue)
      const int q = queue_of(qpref[wib], item);
      const int idx = qpref[wib][q] + atomicAdd(cnt + W_QHIT + q, 1);
      if (idx < pb.hit_cap) {
        HitRec<T> &r = pb.hits[idx];
        r.env = w.x; r.packed = w.y; r.n = n; r.hintA = ha; r.hintB = hb; r.pad = 0;
        for (int k = 0; k < n; k++)
#pragma unroll
          for (int c = 0; c < 3; c++) { r.S[k][c] = Sx[k].w[c]; r.S[k][3 + c] = Sx[k].a[c]; r.S[k][6 + c] = Sx[k].b[c]; }
      } else DROPCAT(6, 1);
    }
  }
  if (S.prof) {
    git_sum = warp_sum(git_sum);
    if (lane == 0 && git_sum) atomicAdd(S.prof + P_GJKIT, (unsigned long long)git_sum);
  }
}

// contacts of one pair -> the env's raw buffer, with the (pair, manifold index) sort key
template <typename T>
__device__ __forceinline__ void emit_pair_contacts(const PipeBuf<T> &pb, const PairContacts<T> &pc, int env, int g1, int g2, int pidx) {
  if (pc.n <= 0) return;
  const int base = atomicAdd(pb.ncon_raw + env, pc.n);
  for (int c = 0; c < pc.n; c++)
    if (base + c < CONBUF) {
      T *dst = pb.con + ((size_t)env * CONBUF + base + c) * 8;
      dst[0] = pc.normal[0]; dst[1] = pc.normal[1]; dst[2] = pc.normal[2];
      dst[3] = pc.pos[c][0]; dst[4] = pc.pos[c][1]; dst[5] = pc.pos[c][2];
      dst[6] = pc.dist[c];
      pb.con_key[(size_t)env * CONBUF + base + c] = (pidx << 20) | (c << 16) | (g1 << 8) | g2;
    }
}

// Per substep, ONE THREAD per intersecting pair: EPA -> manifold (scene_collide_seq.cuh), contacts to the env's raw buffer.
#ifndef SO101_NSEQ_MINCTAS
#define SO101_NSEQ_MINCTAS 12
#endif
constexpr int NSEQ_THREADS = 64, NSEQ_MINCTAS = SO101_NSEQ_MINCTAS;
template <typename T>
__global__ void __launch_bounds__(NSEQ_THREADS, NSEQ_MINCTAS) scene_narrow_seq_kernel(const __grid_constant__ SceneModel<T> sm, const EnvState<T> S, const PipeBuf<T> pb, int sub) {
  __shared__ int qpref[NSEQ_THREADS / 32][WQ + 1], hpref[NSEQ_THREADS / 32][WQ + 1];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int *cnt = pb.nwork + WSTRIDE * sub;
  build_qpref(qpref[wib], cnt, pb.work_cap, lane);            // where each queue's slots start
  build_qpref(hpref[wib], cnt + W_QHIT, pb.work_cap, lane);   // hits per queue -> item numbering of this kernel
  const int nhit = hpref[wib][WQ];
  const int stride = gridDim.x * NSEQ_THREADS;
  long long eit_sum = 0, nepa = 0;
#pragma unroll 1
  for (int item = blockIdx.x * NSEQ_THREADS + threadIdx.x; item < nhit; item += stride) {
    const long long t0 = clock64();
    const int hq = queue_of(hpref[wib], item), hslot = qpref[wib][hq] + (item - hpref[wib][hq]);
    if (hslot >= pb.hit_cap) continue;
    const HitRec<T> &rec = pb.hits[hslot];
    const int env = (int)rec.env, g1 = (int)(rec.packed & 0xff), g2 = (int)((rec.packed >> 8) & 0xff), pidx = (int)(rec.packed >> 16);
    const T(*xpos)[3] = reinterpret_cast<const T(*)[3]>(pb.xpos + (size_t)env * (NSLOT * 3));
    const T(*xmat)[9] = reinterpret_cast<const T(*)[9]>(pb.xmat + (size_t)env * (NSLOT * 9));
    Shape<T> A, B;
    make_shape(sm, xpos, xmat, g1, A);
    make_shape(sm, xpos, xmat, g2, B);
    A.hint = rec.hintA; B.hint = rec.hintB;  // continue the hill-climbing warm starts where GJK left them
    CollideScratch<T> cs;
    PairContacts<T> pc;
    pc.n = 0;
    long long t_epa = t0;
    if (A.type == G_PLANE) collide_plane_seq(sm, cs, A, B, pc);
    else {
      int eit = 0;
      collide_convex_seq(sm, cs, A, B, reinterpret_cast<const MPoint<T> *>(&rec.S[0][0]), rec.n, pc, eit, t_epa);
      eit_sum += eit; nepa++;
      if (S.prof) atomicAdd(&g_epahist[eit <= 2 ? 0 : eit <= 5 ? 1 : eit <= 10 ? 2 : eit <= 20 ? 3 : eit <= 40 ? 4 : eit < EPA_MAXIT ? 5 : 6], 1);
    }
    emit_pair_contacts(pb, pc, env, g1, g2, pidx);
    if (S.prof) {  // stage probe: how long the lanes that were active in this trip spent in EPA and in the manifold stage
      const long long t1 = clock64();
      long long e = t_epa - t0, m = t1 - t_epa;
      const unsigned act = __activemask();
      for (int o = 16; o > 0; o >>= 1) { e = max(e, __shfl_xor_sync(act, e, o)); m = max(m, __shfl_xor_sync(act, m, o)); }
      if (lane == __ffs(act) - 1) {
        const long long tot = e + m;
        atomicAdd(&g_nprof[0], (unsigned long long)e); atomicAdd(&g_nprof[1], (unsigned long long)m); atomicAdd(&g_nprof[2], 1ull);
        atomicMax(&g_nprof[3], ((unsigned long long)tot << 8) | (unsigned)g2);
        atomicMax(&g_nprof[4], (unsigned long long)e); atomicMax(&g_nprof[5], (unsigned long long)m);
        const int bin = tot < 100000 ? 0 : tot < 200000 ? 1 : tot < 400000 ? 2 : tot < 800000 ? 3 : tot < 1600000 ? 4 : tot < 3200000 ? 5 : 6;
        atomicAdd(&g_nprof[8 + bin], 1ull);
      }
    }
  }
  if (S.prof) {
    eit_sum = warp_sum(eit_sum); nepa = warp_sum(nepa);
    if (lane == 0 && nepa) { atomicAdd(S.prof + P_EPAIT, (unsigned long long)eit_sum); atomicAdd(S.prof + P_NEPA, (unsigned long long)nepa); }
  }
}

// The same work as scene_narrow_seq_kernel in TWO launches, for large batches: STAGE 0 runs EPA for every intersecting convex pair
// and leaves (normal, depth, witness points, hints) in the pair's hit record, where its GJK simplex was; STAGE 1 builds the
// manifold (or the plane contacts) and emits.  Each stage carries half of the fused kernel's 12 k instructions (the fused kernel
// spends 3 of its 13 stall cycles per instruction waiting for instruction fetch at full occupancy) and touches only its half
// of the per-thread scratch.  At small batches, where every launch lasts as long as its slowest warp, the second launch's own
// tail costs more than it saves, so launch_scene_step picks by group size.  Same arithmetic, same results.
template <typename T, int STAGE>
__global__ void __launch_bounds__(NSEQ_THREADS, NSEQ_MINCTAS) scene_narrow_split_kernel(const __grid_constant__ SceneModel<T> sm, const EnvState<T> S, const PipeBuf<T> pb, int sub) {
  __shared__ int qpref[NSEQ_THREADS / 32][WQ + 1], hpref[NSEQ_THREADS / 32][WQ + 1];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int *cnt = pb.nwork + WSTRIDE * sub;
  build_qpref(qpref[wib], cnt, pb.work_cap, lane);
  build_qpref(hpref[wib], cnt + W_QHIT, pb.work_cap, lane);
  const int nhit = hpref[wib][WQ];
  const int stride = gridDim.x * NSEQ_THREADS;
  int *cursor = cnt + W_CURSOR;
#pragma unroll 1
  for (int it0 = blockIdx.x * NSEQ_THREADS + threadIdx.x;; it0 += stride) {
    int item = it0;
    if (STAGE == 1) {
      // a warp takes 32 consecutive pairs (same queue = same second geom: its hull loads stay warp-wide broadcasts) from the group's
      // cursor: a 1000-vertex banana hull costs a warp twenty times a box, and a static assignment leaves the tail to the unlucky
      int base = 0;
      if (lane == 0) base = atomicAdd(cursor, 32);
      base = __shfl_sync(FULL, base, 0);
      if (base >= nhit) break;
      item = base + lane;
      if (item >= nhit) continue;
    } else if (item >= nhit) break;
    const int hq = queue_of(hpref[wib], item), hslot = qpref[wib][hq] + (item - hpref[wib][hq]);
    if (hslot >= pb.hit_cap) continue;
    HitRec<T> &rec = pb.hits[hslot];
    const int env = (int)rec.env, g1 = (int)(rec.packed & 0xff), g2 = (int)((rec.packed >> 8) & 0xff), pidx = (int)(rec.packed >> 16);
    if (STAGE == 0 && sm.geom_type[g1] == G_PLANE) continue;   // plane pairs have no EPA stage
    const T(*xpos)[3] = reinterpret_cast<const T(*)[3]>(pb.xpos + (size_t)env * (NSLOT * 3));
    const T(*xmat)[9] = reinterpret_cast<const T(*)[9]>(pb.xmat + (size_t)env * (NSLOT * 9));
    Shape<T> A, B;
    make_shape(sm, xpos, xmat, g1, A);
    make_shape(sm, xpos, xmat, g2, B);
    A.hint = rec.hintA; B.hint = rec.hintB;
    CollideScratch<T> cs;
    T *r = &rec.S[0][0];
    if (STAGE == 0) {
      T normal[3], depth = T(0), pa[3] = {T(0), T(0), T(0)}, pb_[3] = {T(0), T(0), T(0)};
      int eit = 0;
      const int ok = epa_seq(sm, cs, A, B, reinterpret_cast<const MPoint<T> *>(r), rec.n, normal, depth, pa, pb_, eit);
      // the simplex is dead now: the EPA result takes its place, together with the hill-climbing hints the manifold continues from
      if (ok) { r[0] = normal[0]; r[1] = normal[1]; r[2] = normal[2]; r[3] = depth; r[4] = pa[0]; r[5] = pa[1]; r[6] = pa[2]; r[7] = pb_[0]; r[8] = pb_[1]; r[9] = pb_[2]; }
      rec.pad = ok; rec.hintA = A.hint; rec.hintB = B.hint;
    } else {
      PairContacts<T> pc;
      pc.n = 0;
      if (A.type == G_PLANE) collide_plane_seq(sm, cs, A, B, pc);
      else {
        const T normal[3] = {r[0], r[1], r[2]}, pa[3] = {r[4], r[5], r[6]}, pb_[3] = {r[7], r[8], r[9]};
        finish_convex_pair(sm, cs, A, B, rec.pad, normal, r[3], pa, pb_, pc);
      }
      emit_pair_contacts(pb, pc, env, g1, g2, pidx);
    }
  }
}

// EPA stage of the two-launch narrow phase as a per-lane state machine: every lane owns one pair at a time and the warp runs
// ONE polytope expansion per round for all of its lanes, whatever pair and iteration each lane is at; a lane whose pair has
// converged takes the next pair from the group's cursor (warp-aggregated atomicAdd) once `refill_min` lanes are waiting, so the
// set-up code runs for many lanes at a time.  Run-to-completion per pair (scene_narrow_split_kernel<T, 0>) keeps 5-6 of 32
// lanes busy, because iteration counts range from 1 to 50 inside a warp; here the lanes a round occupies are those still
// expanding.  Same per-pair arithmetic, same results.
template <typename T>
__global__ void __launch_bounds__(NSEQ_THREADS, NSEQ_MINCTAS) scene_epa_kernel(const __grid_constant__ SceneModel<T> sm, const EnvState<T> S, const PipeBuf<T> pb, int sub, int refill_min) {
  __shared__ int qpref[NSEQ_THREADS / 32][WQ + 1], hpref[NSEQ_THREADS / 32][WQ + 1];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int *cnt = pb.nwork + WSTRIDE * sub;
  build_qpref(qpref[wib], cnt, pb.work_cap, lane);
  build_qpref(hpref[wib], cnt + W_QHIT, pb.work_cap, lane);
  const int nhit = hpref[wib][WQ];
  int *cursor = cnt + W_HITCURSOR;
  const unsigned lt = (1u << lane) - 1u;
  Shape<T> A, B;
  CollideScratch<T> cs;
  EpaState<T> st;
  HitRec<T> *rec = nullptr;
  bool busy = false, drained = false;
#pragma unroll 1
  while (true) {
    const unsigned idle = __ballot_sync(FULL, !busy && !drained), act = __ballot_sync(FULL, busy);
    if (!idle && !act) break;
    if (idle && (!act || __popc(idle) >= refill_min)) {
      const int leader = __ffs(idle) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(cursor, __popc(idle));
      base = __shfl_sync(FULL, base, leader);
      if (!busy && !drained) {
        const int item = base + __popc(idle & lt);
        if (item >= nhit) drained = true;
        else {
          const int hq = queue_of(hpref[wib], item), hslot = qpref[wib][hq] + (item - hpref[wib][hq]);
          if (hslot < pb.hit_cap) {
            rec = &pb.hits[hslot];
            const int env = (int)rec->env, g1 = (int)(rec->packed & 0xff), g2 = (int)((rec->packed >> 8) & 0xff);
            if (sm.geom_type[g1] != G_PLANE) {   // (plane pairs have no EPA stage)
              const T(*xpos)[3] = reinterpret_cast<const T(*)[3]>(pb.xpos + (size_t)env * (NSLOT * 3));
              const T(*xmat)[9] = reinterpret_cast<const T(*)[9]>(pb.xmat + (size_t)env * (NSLOT * 9));
              make_shape(sm, xpos, xmat, g1, A);
              make_shape(sm, xpos, xmat, g2, B);
              A.hint = rec->hintA; B.hint = rec->hintB;
              if (epa_begin(sm, cs, A, B, reinterpret_cast<const MPoint<T> *>(&rec->S[0][0]), rec->n, st)) busy = true;
              else { rec->pad = 0; rec->hintA = A.hint; rec->hintB = B.hint; }
            }
          }
        }
      }
    }
    if (busy) {
      const int r_ = epa_step(sm, cs, A, B, st);
      if (r_ != 0) {
        T normal[3], depth = T(0), pa[3], pb_[3];
        const int ok = r_ > 0 ? epa_end(cs, st, normal, depth, pa, pb_) : 0;
        // the simplex is dead now: the EPA result takes its place, together with the hill-climbing hints the manifold continues from
        T *r = &rec->S[0][0];
        if (ok) { r[0] = normal[0]; r[1] = normal[1]; r[2] = normal[2]; r[3] = depth; r[4] = pa[0]; r[5] = pa[1]; r[6] = pa[2]; r[7] = pb_[0]; r[8] = pb_[1]; r[9] = pb_[2]; }
        rec->pad = ok; rec->hintA = A.hint; rec->hintB = B.hint;
        busy = false;
      }
    }
  }
}

// Gather the env's raw contacts into shared memory in oracle order (pair order, then manifold order).  Returns false (and
// touches nothing) when the env needs more than NC contacts or NB Jacobian blocks: the caller defers it to the large tier.
template <typename T, typename SC>
__device__ __forceinline__ bool gather_contacts(const SceneModel<T> &sm, const PipeBuf<T> &pb, SC &s, int env, int &dropped, int lane) {
  auto &R = s.sol;
  constexpr int NC = sizeof(R.D0) / sizeof(T), NB = sizeof(R.w1[0]) / sizeof(T);
  int nraw = pb.ncon_raw[env];
  if (nraw > NC) { dropped += nraw - NC; if (lane == 0) DROPCAT(4, nraw - NC); nraw = NC; }  // (the classifier picked a tier that fits)
  const int *keys = pb.con_key + (size_t)env * CONBUF;
  constexpr int RSL = (NC + 31) / 32;
  int mykey[RSL], rank[RSL];
  int nblk = 0;
#pragma unroll
  for (int k = 0; k < RSL; k++) {
    const int i = lane + 32 * k;
    mykey[k] = i < nraw ? keys[i] : 0x7fffffff; rank[k] = 0;
    if (i < nraw) {
      const int s1 = sm.body_slot[sm.geom_body[(mykey[k] >> 8) & 0xff]], s2 = sm.body_slot[sm.geom_body[mykey[k] & 0xff]];
      const int a1 = s1 < 0 ? 31 : (s1 < NA ? s1 / NJ : s1), a2 = s2 < 0 ? 31 : (s2 < NA ? s2 / NJ : s2);   // dof block of each side
      nblk += (a1 != 31) + (a2 != 31 && a2 != a1);
    }
  }
  // rank = number of contacts with a smaller key (keys are unique: pair index and manifold index)
#pragma unroll
  for (int kk = 0; kk < RSL; kk++) {
#pragma unroll 1
    for (int jl = 0; jl < 32; jl++) {
      const int j = 32 * kk + jl;
      if (j >= nraw) break;
      const int kj = wshfl(mykey[kk], jl);
#pragma unroll
      for (int k = 0; k < RSL; k++) rank[k] += kj < mykey[k];
    }
  }
#pragma unroll
  for (int k = 0; k < RSL; k++) {
    const int i = lane + 32 * k;
    if (i < nraw) {
      const int c = rank[k];
      const T *src = pb.con + ((size_t)env * CONBUF + i) * 8;
      const T nrm[3] = {src[0], src[1], src[2]};
      T frame[9];
      frame_from_normal(nrm, frame);
#pragma unroll
      for (int e = 0; e < 9; e++) R.v.con.frame[e][c] = frame[e];
      R.v.con.pos[0][c] = src[3]; R.v.con.pos[1][c] = src[4]; R.v.con.pos[2][c] = src[5];
      R.v.con.dist[c] = src[6];
      R.v.con.g1[c] = (mykey[k] >> 8) & 0xff; R.v.con.g2[c] = mykey[k] & 0xff;
    }
  }
  if (lane == 0) { s.ncon = nraw; pb.ncon_raw[env] = 0; }
  __syncwarp();
  return true;
}

// One substep of one env (warp): constraint rows from the gathered contacts and the dyn record, Newton, semi-implicit Euler.  Returns false if the env must be
// handled by the large solver tier instead (nothing has been modified in that case).
template <typename T, typename SC>
__device__ __forceinline__ bool solve_env(const ArmSetT<T> &am, const SceneModel<T> &sm, const StepCfg &cfg, const EnvState<T> &S, const PipeBuf<T> &pb,
                                          const so101_step_out &out, int sub, int env, SC &s, int lane) {
  prof_begin(s, S, lane);
  for (int i = lane; i < NQ; i += 32) s.q[i] = (T)S.qpos[(size_t)env * NQ + i];
  for (int i = lane; i < NV; i += 32) { s.qd[i] = (T)S.qvel[(size_t)env * NV + i]; s.warm[i] = S.warm[(size_t)env * NV + i]; }
  if (lane == 0) s.dbg = 0;
  __syncwarp();
  int iters = 0, dropped = 0;
  const bool last = sub == cfg.nsub - 1;
  PROF_START(s);
  const bool frozen = pb.flags[env] != 0;  // diverged earlier in this control step: no more physics, no more collision work
  if (frozen && lane == 0) pb.ncon_raw[env] = 0;
  if (!frozen && !gather_contacts(sm, pb, s, env, dropped, lane)) return false;
  if (frozen && lane == 0) s.ncon = 0;
  __syncwarp();
  const int mode = S.mode[env];
  if (!frozen && mode && S.settle_sub[env] == 0) {
    // [upstream] PropPlacer(ignore_collisions=False): a fresh sample whose collision-checked prop penetrates anything at its
    // spawn pose is rejected and re-drawn (so100_hand_over.py:216-221), at most max_attempts times
    bool hit = false;
    for (int c = lane; c < s.ncon; c += 32) {
      if (!(s.sol.v.con.dist[c] < T(0))) continue;
      const int s1 = sm.body_slot[sm.geom_body[s.sol.v.con.g1[c]]], s2 = sm.body_slot[sm.geom_body[s.sol.v.con.g2[c]]];
      if ((s1 >= NA && S.place.check_collisions[s1 - NA]) || (s2 >= NA && S.place.check_collisions[s2 - NA])) hit = true;
    }
    if (__any_sync(FULL, hit)) {
      int give_up = 0;
      if (lane == 0) {
        const int a = S.attempt[env] + 1;
        if (a < S.place.max_attempts) { S.attempt[env] = a; S.sstate[env] = SETTLE_SAMPLE; pb.flags[env] = 2; atomicAdd(S.ring_ctr + RC_REJECTED, 1); }
        else { give_up = 1; atomicAdd(S.ring_ctr + RC_EXHAUSTED, 1); }
      }
      if (!__shfl_sync(FULL, give_up, 0)) return true;
    }
  }
  if (!frozen) {
    PROF_CNT(s, P_NCON, s.ncon, lane);
    load_poses(pb, s, env, lane);
    load_dyn(pb, s, env, lane);
    __syncwarp();
    PROF_ACC(s, P_DYN, lane);
    if (S.dbg_contacts && last) {  // parity probe: contacts of the last substep
      float *dst = S.dbg_contacts + (size_t)env * (1 + 9 * NCON);
      const int n = s.ncon < NCON ? s.ncon : NCON;
      if (lane == 0) dst[0] = (float)n;
      for (int c = lane; c < n; c += 32) {
        float *r = dst + 1 + 9 * c;
        r[0] = (float)s.sol.v.con.g1[c]; r[1] = (float)s.sol.v.con.g2[c]; r[2] = (float)s.sol.v.con.dist[c];
        for (int e = 0; e < 3; e++) { r[3 + e] = (float)s.sol.v.con.pos[e][c]; r[6 + e] = (float)s.sol.v.con.frame[e][c]; }
      }
    }
    build_rows(sm, s, dropped, lane);
    PROF_ACC(s, P_ROWS, lane);
    if (lane < NV) s.delta[lane] = s.warm[lane] - s.qacc_s[lane];
    __syncwarp();
    iters = scene_solve(am, sm.impratio, s, cfg.max_iter, (T)cfg.tol, lane);
    __syncwarp();
    PROF_ACC(s, P_SOLVE, lane);
    PROF_CNT(s, P_NEWTON, iters, lane);
    PROF_CNT(s, P_NSUB, 1, lane);
    // [upstream] mj_Euler
    T qacc = T(0);
    if (lane < NV) qacc = s.qacc_s[lane] + s.delta[lane];
    // [upstream] mj_checkAcc: a non-finite / huge acceleration ends the episode (task_suite.py:153); the env is frozen for the
    // rest of this control step (its state stays finite) and resets on the next step() call
    const bool badnow = __any_sync(FULL, lane < NV && !(t_abs(qacc) < T(1e10)));
    if (badnow) {
      if (lane == 0) pb.flags[env] = 1;
      qacc = T(0);
    }
    T qacc_int = qacc;   // acceleration the velocity update uses (differs from qacc under implicitfast only)
    // The update itself runs in float64 on the float64 state (env_state.cuh): velocity first, then positions with the new
    // velocity; free-joint quaternions are advanced by the body-frame angular velocity and renormalised.
    const TS h = am[0].dt_d;
    TS qd_new = TS(0);
    if (cfg.integrator == 1) {
      // implicitfast ([upstream] mj_implicit): the arm's velocity update uses x = qacc + (M - h D)^-1 h D qacc, D = the unclamped
      // actuators' velocity gain (arm_dynamics.cuh); the free props have no velocity-dependent smooth force (x = qacc)
#pragma unroll 1
      for (int k = 0; k < NARM; k++) {
        T qa6[NJ], dv[NJ], cr[NJ], Mm[21];
        const unsigned vm = (unsigned)pb.dyn[(size_t)env * DYNW + DYN_VELMASK + k];
#pragma unroll
        for (int i = 0; i < NJ; i++) { qa6[i] = __shfl_sync(FULL, qacc, NJ * k + i); dv[i] = ((vm >> i) & 1u) ? (T)am[k].bias_d[i][2] : T(0); }
#pragma unroll
        for (int i = 0; i < 21; i++) Mm[i] = s.Marm[k][i];
        implicitfast_correction<T>(Mm, dv, (T)h, qa6, cr);
        T mine = T(0);
#pragma unroll
        for (int i = 0; i < NJ; i++) if (lane == NJ * k + i) mine = cr[i];
        if (lane >= NJ * k && lane < NJ * (k + 1) && !badnow) qacc_int = qacc + mine;
      }
    }
    if (lane < NV) {
      // SETTLE mode: the arm is frozen ([upstream] JointStaticIsolator restores the non-prop joints after every settle step)
      qd_new = (badnow || (mode && lane < NA)) ? TS(0) : S.qvel[(size_t)env * NV + lane] + h * (TS)qacc_int;
      S.qvel[(size_t)env * NV + lane] = qd_new;
      S.warm[(size_t)env * NV + lane] = qacc;
    }
    if (mode && !badnow) {  // [upstream] PropPlacer settle test after every physics step: props' max |qvel|, max |qacc|
      T mv = (lane >= NA && lane < NV) ? (T)t_abs(qd_new) : T(0), ma = (lane >= NA && lane < NV) ? t_abs(qacc) : T(0);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { mv = max(mv, __shfl_xor_sync(FULL, mv, o)); ma = max(ma, __shfl_xor_sync(FULL, ma, o)); }
      if (lane == 0) {
        const int sub_done = S.settle_sub[env] + 1;
        S.settle_sub[env] = sub_done;
        const bool settled = mv < (T)S.place.qvel_tol && ma < (T)S.place.qacc_tol, timeout = sub_done >= S.place.max_settle_substeps;
        if (settled || timeout) {
          S.sstate[env] = SETTLE_DONE; pb.flags[env] = 2;
          if (!settled) atomicAdd(S.ring_ctr + RC_UNSETTLED, 1);
        }
      }
    }
    constexpr int PL = NV;   // lanes PL, PL + 1 integrate the two free joints
    static_assert(PL + NPROP <= 32, "prop lanes");
    TS pv[6];  // the six velocities of prop (lane - PL), gathered from the dof lanes
#pragma unroll
    for (int c = 0; c < 6; c++) {
      const TS a0 = __shfl_sync(FULL, qd_new, NA + c), a1 = __shfl_sync(FULL, qd_new, NA + 6 + c);
      pv[c] = lane == PL + 1 ? a1 : a0;
    }
    if (lane < NA) S.qpos[(size_t)env * NQ + lane] += h * qd_new;
    if (lane >= PL && lane < PL + NPROP) {
      TS *qp = S.qpos + (size_t)env * NQ + NA + 7 * (lane - PL);
      for (int c = 0; c < 3; c++) qp[c] += h * pv[c];
      const TS w[3] = {pv[3], pv[4], pv[5]};
      const TS nw = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]), ang = nw * h;
      TS qn[4] = {qp[3], qp[4], qp[5], qp[6]};
      if (ang > TS(0)) {
        TS sn, cn;
        sincos(TS(0.5) * ang, &sn, &cn);
        const TS qr[4] = {cn, w[0] / nw * sn, w[1] / nw * sn, w[2] / nw * sn};
        quat_mul(qn, qn, qr);
      }
      TS n = sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
      if (n < TS(1e-15)) { qn[0] = TS(1); qn[1] = qn[2] = qn[3] = TS(0); n = TS(1); }
      for (int c = 0; c < 4; c++) qp[3 + c] = qn[c] / n;
    }
  }
  if (last && lane == 0) { S.solver_iter[env] = iters; S.ncon[env] = s.ncon; }
  if (lane == 0 && dropped) { atomicAdd(S.diverged_count + 1, dropped); S.dropped_env[env] += dropped; }
  prof_flush(s, S, lane);
  return true;
}

// Per substep, thread per env: pick the solver tier whose contact / Jacobian-block capacity fits the env's raw contacts and
// queue tier-1 / tier-2 envs, so that the three tier kernels can run concurrently on separate streams.
template <typename T>
__global__ void scene_classify_kernel(const __grid_constant__ SceneModel<T> sm, const EnvState<T> S, const PipeBuf<T> pb, int sub) {
  const int env = pb.env0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= pb.env0 + pb.nenv || !pb.active[env]) return;
  int n = pb.ncon_raw[env];
  if (n > CONBUF) n = CONBUF;
  const int *keys = pb.con_key + (size_t)env * CONBUF;
  int nblk = 0;
  for (int i = 0; i < n; i++) {
    const int k = keys[i];
    const int s1 = sm.body_slot[sm.geom_body[(k >> 8) & 0xff]], s2 = sm.body_slot[sm.geom_body[k & 0xff]];
    const int a1 = s1 < 0 ? 31 : (s1 < NA ? s1 / NJ : s1), a2 = s2 < 0 ? 31 : (s2 < NA ? s2 / NJ : s2);
    nblk += (a1 != 31) + (a2 != 31 && a2 != a1);
  }
  const int tier = (n <= NC_S && nblk <= NB_S) ? 0 : ((n <= NC_M && nblk <= NB_M) ? 1 : 2);
  pb.tier[env] = (uint8_t)tier;
  if (tier > 0) pb.big[(size_t)(tier - 1) * pb.nenv + atomicAdd(pb.nwork + WSTRIDE * sub + W_NTIER + tier - 1, 1)] = env;
}

// Tier 0: warp per env (every env whose contacts fit the small scratch).
template <typename T>
__global__ void __launch_bounds__(WARPS_SOLVE * 32) scene_solve_kernel(const __grid_constant__ ArmSetT<T> am, const __grid_constant__ SceneModel<T> sm,
                                                                      const __grid_constant__ StepCfg cfg, const EnvState<T> S, const PipeBuf<T> pb,
                                                                      const so101_step_out out, int sub) {
  using SC = Scratch<T, NC_S, NB_S>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SC *all = reinterpret_cast<SC *>(smem_raw);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int env = pb.env0 + blockIdx.x * WARPS_SOLVE + wib;
  if (env >= pb.env0 + pb.nenv) return;
  if (!pb.active[env] || pb.tier[env] != 0) return;
  solve_env(am, sm, cfg, S, pb, out, sub, env, all[wib], lane);
}
// Tiers 1 and 2: persistent CTAs walk the tier's queue (filled by the previous tier during this substep).
template <typename T, int NC, int NB, int WARPS, int TIER>
__global__ void __launch_bounds__(WARPS * 32) scene_solve_tier_kernel(const __grid_constant__ ArmSetT<T> am, const __grid_constant__ SceneModel<T> sm,
                                                                     const __grid_constant__ StepCfg cfg, const EnvState<T> S, const PipeBuf<T> pb,
                                                                     const so101_step_out out, int sub) {
  using SC = Scratch<T, NC, NB>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SC *all = reinterpret_cast<SC *>(smem_raw);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  SC &s = all[wib];
  int *cnt = pb.nwork + WSTRIDE * sub;
  const int n = cnt[W_NTIER + TIER - 1];
  const int *list = pb.big + (size_t)(TIER - 1) * pb.nenv;
#pragma unroll 1
  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(cnt + W_TIERCURSOR + TIER - 1, 1);
    item = wshfl(item, 0);
    if (item >= n) break;
    const int env = list[item];
    solve_env(am, sm, cfg, S, pb, out, sub, env, s, lane);
    if (S.prof && lane == 0) atomicAdd(S.prof + P_BIGENV, 1ull);
    __syncwarp();
  }
}

template <typename T>
__global__ void scene_reset_kernel(const __grid_constant__ StepCfg cfg, const EnvState<T> S, const uint8_t *__restrict__ mask, const so101_step_out out) {
  __shared__ BroadEnv<T> all[WARPS_BROAD];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int env = blockIdx.x * WARPS_BROAD + wib;
  if (env >= S.NU || S.mode[env]) return;
  if (mask && !mask[env]) return;
  reset_env_scene(cfg, S, out, all[wib], env, lane);
}

// Parity probe for the reward geometry: the device 6-axis SAT on n caller-supplied box pairs, rows of 20 doubles
// (p0 3, q0 4, half0 3, p1 3, q1 4, half1 3), evaluated in T.  tests/test_scene_gpu.py runs the 240 cases produced by the
// reference's own oobb_utils.py (tests/golden/oobb_overlap.json) through it.
template <typename T>
__global__ void debug_overlap_kernel(const double *__restrict__ cases, int n, uint8_t *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double *c = cases + (size_t)i * 20;
  T v[20];
  for (int k = 0; k < 20; k++) v[k] = (T)c[k];
  out[i] = overlap_oobb_oobb<T>(v, v + 3, v + 7, v + 10, v + 13, v + 17) ? 1 : 0;
}
template <typename T>
void launch_debug_overlap(const double *cases, int n, uint8_t *out, cudaStream_t stream) {
  debug_overlap_kernel<T><<<(n + 63) / 64, 64, 0, stream>>>(cases, n, out);
}

template <typename T>
size_t scene_smem_bytes() {
  const size_t a = sizeof(Scratch<T, NC_M, NB_M>) * WARPS_M, b = sizeof(Scratch<T, NC_S, NB_S>) * WARPS_SOLVE, c = sizeof(Scratch<T, NC_L, NB_L>) * WARPS_L;
  return a > b ? (a > c ? a : c) : (b > c ? b : c);
}

template <typename T>
void scene_dropcat(int out[8]) { cudaMemcpyFromSymbol(out, g_dropcat, sizeof(int) * 8); }
template <typename T>
void scene_epahist(int out[8]) { cudaMemcpyFromSymbol(out, g_epahist, sizeof(int) * 8); }
template <typename T>
void scene_nprof(unsigned long long out[16]) { cudaMemcpyFromSymbol(out, g_nprof, sizeof(unsigned long long) * 16); }

// Launches of one control step: per pipeline group 1 memset + 3 + 8 * nsub kernels on the group's streams, forked from and
// joined back into the caller's stream.  Returns the kernel count.
template <typename T>
int launch_scene_step(const ArmSetT<T> &am, const ArmSetT<double> &am64, const SceneModel<T> &sm, const StepCfg &cfg, const EnvState<T> &S, const PipeBuf<T> *pbs,
                      TierExec *txs, int ngroups, const float *action, const so101_step_out &out, cudaStream_t stream, KernelTimer *kt) {
  int dev = 0;
  cudaGetDevice(&dev);
  const size_t smem_env = sizeof(Scratch<T, NC_S, NB_S>) * WARPS_SOLVE,
               smem_m = sizeof(Scratch<T, NC_M, NB_M>) * WARPS_M, smem_l = sizeof(Scratch<T, NC_L, NB_L>) * WARPS_L;
  // (no function-local statics here: template statics are process-wide unique symbols, and the one-arm and the two-arm builds
  // of this library can be loaded into the same process)
  const size_t smem_kd = sizeof(T) * 32 * KD_STRIDE;
  cudaFuncSetAttribute(scene_kindyn_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kd);
  cudaFuncSetAttribute(scene_solve_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_env);
  cudaFuncSetAttribute(scene_solve_tier_kernel<T, NC_M, NB_M, WARPS_M, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_m);
  cudaFuncSetAttribute(scene_solve_tier_kernel<T, NC_L, NB_L, WARPS_L, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l);
  KernelTimer none;
  KernelTimer &t = kt ? *kt : none;
  cudaEventRecord(txs[0].start, stream);  // every group starts after the work already queued on the caller's stream
  for (int g = 0; g < ngroups; g++) cudaStreamWaitEvent(txs[g].main, txs[0].start, 0);
  // kinematics (+ smooth dynamics) at the current state, then the broad phase for substep `sub` or the task layer
  auto refresh = [&](const PipeBuf<T> &pb, cudaStream_t st, int sub, bool last) {
    t.begin(6, st);
    scene_kindyn_kernel<T><<<(pb.nenv + KD_THREADS - 1) / KD_THREADS, KD_THREADS, smem_kd, st>>>(am, am64, sm, S, pb, last ? 0 : 1);
    t.end(6, st);
    t.begin(7, st);
    scene_broad_kernel<T><<<(pb.nenv + WARPS_BROAD - 1) / WARPS_BROAD, WARPS_BROAD * 32, 0, st>>>(sm, cfg, S, pb, out, sub, last ? 1 : 0);
    t.end(7, st);
  };
  // While the per-kernel event timers are on, every launch of every group goes to the caller's stream: a launch is then timed
  // ALONE, as under ncu, instead of sharing the SMs with the other group's kernels and the side-stream solver tiers (which made
  // a 54 us kinematics launch read 1.6 ms).  The untimed product path interleaves the groups on their own streams.
  const bool serial = kt && kt->on;
  // interleave the groups' launches substep by substep so that their kernels are in flight together
  for (int g = 0; g < ngroups; g++) {
    const PipeBuf<T> &pb = pbs[g];
    cudaStream_t st = serial ? stream : txs[g].main;
    cudaMemsetAsync(pb.nwork, 0, sizeof(int) * WSTRIDE * (cfg.nsub + 1), st);
    t.begin(0, st);
    scene_begin_kernel<T><<<(pb.nenv + WARPS_BROAD - 1) / WARPS_BROAD, WARPS_BROAD * 32, 0, st>>>(cfg, S, pb, action, out);
    t.end(0, st);
    refresh(pb, st, 0, false);
  }
  int sms = 0;   // SMs of the device the handle lives on (148 on B200): grids are sized in multiples of it
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  const int seq_per_sm = getenv("SO101_SEQ_CTAS") ? atoi(getenv("SO101_SEQ_CTAS")) : 16;  // narrow-phase CTAs of 64 threads per SM
  // groups of at least this many envs run the narrow phase as two launches (scene_narrow_split_kernel)
  const int split_min = getenv("SO101_NARROW_SPLIT") ? atoi(getenv("SO101_NARROW_SPLIT")) : 32768;
  // lanes of a warp that wait for a new pair before the EPA kernel fetches (0: run-to-completion EPA, scene_narrow_split_kernel<T, 0>)
  const int epa_refill = getenv("SO101_EPA_REFILL") ? atoi(getenv("SO101_EPA_REFILL")) : 20;
  int nlaunch = ngroups * (3 + 8 * cfg.nsub);
  for (int sub = 0; sub < cfg.nsub; sub++) {
    for (int g = 0; g < ngroups; g++) {
      const PipeBuf<T> &pb = pbs[g];
      TierExec &tx = txs[g];
      cudaStream_t st = serial ? stream : tx.main, st_m = serial ? stream : tx.sm, st_l = serial ? stream : tx.sl;
      const int grid_env = (pb.nenv + WARPS_SOLVE - 1) / WARPS_SOLVE;
      const int grid_gjk = max(1, min(sms * 16, (pb.nenv * 16 + GJK_THREADS - 1) / GJK_THREADS));
      const int grid_seq = max(1, min(sms * seq_per_sm, (pb.nenv * 12 + NSEQ_THREADS - 1) / NSEQ_THREADS));
      const int grid_m = pb.nenv < sms * 8 ? pb.nenv : sms * 8, grid_l = pb.nenv < sms * 4 ? pb.nenv : sms * 4;
      t.begin(5, st);
      scene_gjk_kernel<T><<<grid_gjk, GJK_THREADS, 0, st>>>(sm, S, pb, sub);
      t.end(5, st);
      if (pb.nenv >= split_min && !S.prof) {   // (the stage probes live in the fused kernel)
        t.begin(8, st);
        if (epa_refill > 0) scene_epa_kernel<T><<<grid_seq, NSEQ_THREADS, 0, st>>>(sm, S, pb, sub, epa_refill);
        else scene_narrow_split_kernel<T, 0><<<grid_seq, NSEQ_THREADS, 0, st>>>(sm, S, pb, sub);
        t.end(8, st);
        t.begin(9, st);
        scene_narrow_split_kernel<T, 1><<<grid_seq, NSEQ_THREADS, 0, st>>>(sm, S, pb, sub);
        t.end(9, st);
        nlaunch++;
      } else {
        t.begin(1, st);
        scene_narrow_seq_kernel<T><<<grid_seq, NSEQ_THREADS, 0, st>>>(sm, S, pb, sub);
        t.end(1, st);
      }
      scene_classify_kernel<T><<<(pb.nenv + 127) / 128, 128, 0, st>>>(sm, S, pb, sub);
      // the three solver tiers work on disjoint envs: tiers 1 and 2 run on side streams beside tier 0 and join before the
      // next kernel
      cudaEventRecord(tx.fork, st);
      cudaStreamWaitEvent(st_m, tx.fork, 0); cudaStreamWaitEvent(st_l, tx.fork, 0);
      // largest tier first: its few, long envs should not start last
      t.begin(10, st_l);
      scene_solve_tier_kernel<T, NC_L, NB_L, WARPS_L, 2><<<grid_l, WARPS_L * 32, smem_l, st_l>>>(am, sm, cfg, S, pb, out, sub);
      t.end(10, st_l);
      t.begin(3, st_m);
      scene_solve_tier_kernel<T, NC_M, NB_M, WARPS_M, 1><<<grid_m, WARPS_M * 32, smem_m, st_m>>>(am, sm, cfg, S, pb, out, sub);
      t.end(3, st_m);
      cudaEventRecord(tx.joinm, st_m); cudaEventRecord(tx.joinl, st_l);
      t.begin(2, st);
      scene_solve_kernel<T><<<grid_env, WARPS_SOLVE * 32, smem_env, st>>>(am, sm, cfg, S, pb, out, sub);
      t.end(2, st);
      cudaStreamWaitEvent(st, tx.joinm, 0); cudaStreamWaitEvent(st, tx.joinl, 0);
      refresh(pb, st, sub + 1, sub == cfg.nsub - 1);
    }
  }
  for (int g = 0; g < ngroups; g++) {
    cudaEventRecord(txs[g].done, serial ? stream : txs[g].main);
    cudaStreamWaitEvent(stream, txs[g].done, 0);
  }
  return nlaunch;
}
template <typename T>
void launch_scene_reset(const StepCfg &cfg, const EnvState<T> &S, const uint8_t *mask, const so101_step_out &out, cudaStream_t stream) {
  scene_reset_kernel<T><<<(S.NU + WARPS_BROAD - 1) / WARPS_BROAD, WARPS_BROAD * 32, 0, stream>>>(cfg, S, mask, out);
}

// so101_sample_and_settle support: put the user envs into SETTLE mode with a fresh draw / take the settled states as the envs'
// initial states and return to normal stepping
template <typename T>
__global__ void settle_enter_kernel(const EnvState<T> S, unsigned draw0) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= S.NU) return;
  S.mode[env] = 1; S.sstate[env] = SETTLE_SAMPLE; S.attempt[env] = 0; S.draws[env] = draw0; S.settle_sub[env] = 0;
  S.needs_reset[env] = 0;
}
template <typename T>
__global__ void settle_leave_kernel(const EnvState<T> S) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= S.NU) return;
  for (int i = 0; i < NQ; i++) S.init_qpos[(size_t)env * NQ + i] = i < NA ? TS(0) : S.qpos[(size_t)env * NQ + i];
  for (int i = 0; i < NV; i++) S.init_qvel[(size_t)env * NV + i] = i < NA ? TS(0) : S.qvel[(size_t)env * NV + i];
  S.mode[env] = 0; S.sstate[env] = SETTLE_SAMPLE; S.episode[env] = 0; S.draws[env] += 1;
}
template <typename T>
void launch_settle_enter(const EnvState<T> &S, unsigned draw0, cudaStream_t stream) { settle_enter_kernel<T><<<(S.NU + 127) / 128, 128, 0, stream>>>(S, draw0); }
template <typename T>
void launch_settle_leave(const EnvState<T> &S, cudaStream_t stream) { settle_leave_kernel<T><<<(S.NU + 127) / 128, 128, 0, stream>>>(S); }

}  // namespace so101
