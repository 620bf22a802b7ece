// Control step of the FULL contact scene (BASELINE config 3: SO100HandOverBanana, nq=20 nv=18) as a pipeline of
// three kernels per physics substep, all envs in lockstep:
//
//   scene_begin_kernel   (once per control step, warp per env)  auto-reset, action -> ctrl, kinematics, broad/mid phase
//   scene_narrow_kernel  (per substep, warp per candidate geom PAIR drawn from a device-wide work list)
//                        convex narrow phase: boolean GJK -> EPA -> support-feature clipping (multiccd manifold)
//   scene_solve_kernel   (per substep, warp per env) smooth dynamics, contact gather, constraint rows, elliptic-cone
//                        Newton, semi-implicit Euler, then kinematics + broad phase of the NEXT substep, or (last
//                        substep) the task layer: observation delay rings, SO100HandOver reward, discount, time limit.
//
// Splitting by stage keeps each kernel's code and shared-memory footprint small (more resident warps, no instruction-
// cache thrash) and turns the narrow phase — whose cost varies 10x between pairs — into a flat, dynamically balanced
// work list.  What crosses kernels (body poses, pair list, raw contacts; ~1 KB per env) stays L2-resident.
// Per substep order ([upstream] mj_step; legacy dm_control order is equivalent):
//   arm FK / CRB / RNE            (all lanes redundantly, registers; arm_dynamics.cuh)
//   prop kinematics, M, bias      (free joints: linear dofs world frame, angular dofs body frame)
//   collision                     (scene_collide.cuh: lanes over vertices / faces)
//   constraint rows               (lane per contact: parameter mixing, impedance, Jacobian blocks, aref)
//   Newton solve                  (elliptic cones, lane per contact for row work, lane per entry for the Hessian)
//   semi-implicit Euler           (quaternion integration for the free joints)
// Task layer: so100_task.py:266-368, so100_hand_over.py:238-275.
#include "scene_kernel.cuh"

namespace so101 {

constexpr int WARPS_SOLVE = 2;   // envs (warps) per CTA in the begin / solve kernels
constexpr int WARPS_NARROW = 4;  // pairs (warps) in flight per CTA in the narrow-phase kernel
constexpr int CSL = (NCON + 31) / 32;  // contact slots per lane
constexpr int NH = NV * (NV + 1) / 2;  // 171 packed lower-triangular entries
constexpr int NOUT = 8;                // contacts one pair can emit (manifold <= MAXMANI)

template <typename T>
struct SolveScratch {
  // Jacobian storage is a pool of 6x6 blocks (rows x dofs of ONE dynamic body); a contact owns one block per dynamic body
  // it touches (prop-vs-table: 1, grasp / prop-vs-prop: 2), block index fastest so that per-lane access is conflict-free.
  T J[36][NBLK];
  T w1[6][NBLK], w2[6][NBLK];  // J^T v1, J^T v2 per block: J^T Hc J = w1 w1^T - w2 w2^T + J^T diag(e) J
  T aref[6][NCON], e[6][NCON];
  T D0[NCON], mu[NCON], fri[3][NCON];
  int info[NCON];      // dim | baseA << 8 | baseB << 16   (dof base 0 / 6 / 12, 31 = no block)
  int blk[2][NCON];    // pool index of block A / block B (-1 = none)
};

template <typename T>
struct BroadScratch {
  unsigned pairq[PAIRCAP];
  T gcenter[3][96];  // world bounding-sphere centres of all geoms
};

// per-warp scratch of the narrow-phase kernel: one candidate pair at a time
template <typename T>
struct NarrowScratch {
  CollideScratch<T> col;
  int ncon, dbg, profon;
  long long prof[16];
  T c_pos[3][NOUT], c_normal[3][NOUT], c_dist[NOUT];
};

// per-env scratch of the begin / solve kernels
template <typename T>
struct Scratch {
  T xpos[NSLOT][3], xmat[NSLOT][9];
  T arm_p[NJ][3], arm_a[NJ][3];
  T q[NQ], qd[NV], warm[NV], ctrl[NJ];
  T Mprop[NPROP][21];
  T Marm[21];
  T H[NH];
  T qacc_s[NV], fsm[NV], delta[NV], grad[NV], search[NV], Md[NV], Ms[NV];
  int ncon, dbg, profon;
  long long prof[16];  // developer probe (SO101_PROFILE=1): per-stage clock64 sums and counters of this env
  // contacts gathered from the narrow phase, in oracle order
  T c_pos[3][NCON], c_frame[9][NCON], c_dist[NCON];
  int c_g1[NCON], c_g2[NCON];
  union U {
    BroadScratch<T> broad;
    SolveScratch<T> sol;
    __device__ U() {}
  } u;
};

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }  // i >= j

// stage profiler: lane 0 accumulates clock64 deltas into Scratch::prof when the handle was created with SO101_PROFILE=1
enum { P_DYN = 0, P_BROAD, P_PLANE, P_GJK, P_EPA, P_MANI, P_ROWS, P_SOLVE, P_INTEG, P_TASK, P_NPQ, P_NCON, P_NEWTON, P_LINE, P_NEPA, P_NSUB };
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

// positions: arm FK (registers) -> shared poses; prop poses.  Returns the arm kinematic state for the CRB/RNE sweep.
template <typename T>
__device__ __forceinline__ void scene_kinematics(const ArmModelT<T> &am, Scratch<T> &s, ArmKin<T> &k, int lane) {
  T qa[NJ];
#pragma unroll
  for (int i = 0; i < NJ; i++) qa[i] = s.q[i];
  T R[NJ][9];
  arm_fk<T>(am, qa, k, R);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NJ; i++) {
      s.xpos[i][0] = k.p[i].x; s.xpos[i][1] = k.p[i].y; s.xpos[i][2] = k.p[i].z;
      s.arm_p[i][0] = k.p[i].x; s.arm_p[i][1] = k.p[i].y; s.arm_p[i][2] = k.p[i].z;
      s.arm_a[i][0] = k.a[i].x; s.arm_a[i][1] = k.a[i].y; s.arm_a[i][2] = k.a[i].z;
#pragma unroll
      for (int e = 0; e < 9; e++) s.xmat[i][e] = R[i][e];
    }
  }
  if (lane < NPROP) {
    const T *qp = s.q + NJ + 7 * lane;
    T Rp[9];
    prop_rotation(qp + 3, Rp);
#pragma unroll
    for (int c = 0; c < 3; c++) s.xpos[NJ + lane][c] = qp[c];
#pragma unroll
    for (int e = 0; e < 9; e++) s.xmat[NJ + lane][e] = Rp[e];
  }
  __syncwarp();
}

// free-joint mass block (packed lower 6x6) and bias force for prop p (uniform; [upstream] mj_crb / mj_rne for a free body)
template <typename T>
__device__ __forceinline__ void prop_dynamics(const SceneModel<T> &sm, const ArmModelT<T> &am, const Scratch<T> &s, int p, T (&M)[21], T (&bias)[6]) {
  const T *R = s.xmat[NJ + p];
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
  const T *qd = s.qd + NJ + 6 * p;
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

// ------------------------------------------------------------------------------------------------ collision driver
template <typename T>
__device__ __forceinline__ void emit_contact(NarrowScratch<T> &s, int &ncon, int &dropped, int g1, int g2, const T *frame, const T *pos, T dist) {
  // lane 0 only; the pair's contacts are staged in shared memory and flushed to the env's raw contact buffer by the caller
  if (ncon >= NOUT) { dropped++; return; }
  const int c = ncon++;
  s.c_dist[c] = dist;
  for (int e = 0; e < 3; e++) { s.c_pos[e][c] = pos[e]; s.c_normal[e][c] = frame[e]; }
}

template <typename T>
__device__ int manifold(const SceneModel<T> &sm, NarrowScratch<T> &s, const Shape<T> &A, const Shape<T> &B, const T *n, T depth, int &ncon, int &dropped, int lane) {
  CollideScratch<T> &cs = s.col;
  T frame[9];
  frame_from_normal(n, frame);
  const T *t1 = frame + 3, *t2 = frame + 6;
  const T nn[3] = {-frame[0], -frame[1], -frame[2]};
  const T delta = depth + T(1e-7);
  const int na = feature(sm, cs, A, frame, t1, t2, delta, cs.FA, lane);
  const int nb = feature(sm, cs, B, nn, t1, t2, delta, cs.FB, lane);
  int u = 0;
  if (lane == 0) {
    if (s.dbg) {
      printf("MANIFOLD g=(%d,%d) n=(%.17g,%.17g,%.17g) depth=%.17g na=%d nb=%d\n", A.geom, B.geom, (double)n[0], (double)n[1], (double)n[2], (double)depth, na, nb);
      for (int i = 0; i < na; i++) printf("  FA[%d]=(%.17g,%.17g,%.17g)\n", i, (double)cs.FA[i].x, (double)cs.FA[i].y, (double)cs.FA[i].h);
      for (int i = 0; i < nb; i++) printf("  FB[%d]=(%.17g,%.17g,%.17g)\n", i, (double)cs.FB[i].x, (double)cs.FB[i].y, (double)cs.FB[i].h);
    }
    for (int i = 0; i < nb; i++) cs.FB[i].h = -cs.FB[i].h;
    int nr = 0;
    if (na >= 3 && nb >= 3) nr = clip_poly(cs, cs.FA, na, cs.FB, nb, cs.R);
    else if (na >= 3 && nb == 2) nr = clip_poly(cs, cs.FB, nb, cs.FA, na, cs.R);
    else if (nb >= 3 && na <= 2) nr = clip_poly(cs, cs.FA, na, cs.FB, nb, cs.R);
    else if (na >= 3 && nb == 1) nr = clip_poly(cs, cs.FB, nb, cs.FA, na, cs.R);
    int k = 0;
    for (int i = 0; i < nr; i++) {
      const T ha = feature_height(cs.FA, na, cs.R[i].x, cs.R[i].y), hb = feature_height(cs.FB, nb, cs.R[i].x, cs.R[i].y);
      const T di = hb - ha;
      if (di < T(0)) { cs.R[k] = cs.R[i]; cs.R[k].h = T(0.5) * (ha + hb); cs.mdist[k] = di; k++; }
    }
    for (int i = 0; i < k; i++) {
      int dup = 0;
      for (int j = 0; j < u; j++)
        if (t_abs(cs.R[i].x - cs.R[j].x) + t_abs(cs.R[i].y - cs.R[j].y) < T(1e-7)) {
          dup = 1;
          if (cs.mdist[i] < cs.mdist[j]) { cs.R[j] = cs.R[i]; cs.mdist[j] = cs.mdist[i]; }
          break;
        }
      if (!dup) { cs.R[u] = cs.R[i]; cs.mdist[u] = cs.mdist[i]; u++; }
    }
    if (s.dbg) { printf("  nr=%d k=%d u=%d\n", nr, k, u); for (int i = 0; i < u; i++) printf("  R[%d]=(%.17g,%.17g) d=%.17g\n", i, (double)cs.R[i].x, (double)cs.R[i].y, (double)cs.mdist[i]); }
    u = reduce_manifold(cs.R, cs.mdist, u);
    for (int i = 0; i < u; i++) {
      T pos[3];
      for (int c = 0; c < 3; c++) pos[c] = cs.R[i].x * t1[c] + cs.R[i].y * t2[c] + cs.R[i].h * frame[c];
      emit_contact(s, ncon, dropped, A.geom, B.geom, frame, pos, cs.mdist[i]);
    }
  }
  u = wshfl(u, 0);
  ncon = wshfl(ncon, 0); dropped = wshfl(dropped, 0);
  __syncwarp();
  return u;
}

template <typename T>
__device__ void collide_convex(const SceneModel<T> &sm, NarrowScratch<T> &s, const Shape<T> &A, const Shape<T> &B, int &ncon, int &dropped, int lane) {
  MPoint<T> S[4];
  int n = 0;
  PROF_START(s);
  const int hit = gjk_intersect(sm, A, B, S, n, lane);
  PROF_ACC(s, P_GJK, lane);
  if (!hit) return;
  T normal[3], depth, pa[3], pb[3];
  PROF_CNT(s, P_NEPA, 1, lane);
  const int ok = epa(sm, s.col, A, B, S, n, normal, depth, pa, pb, lane);
  PROF_ACC(s, P_EPA, lane);
  if (!ok) return;
  if (!(depth > T(0))) return;
  const int nm = manifold(sm, s, A, B, normal, depth, ncon, dropped, lane);
  PROF_ACC(s, P_MANI, lane);
  if (nm > 0) return;
  T frame[9], pos[3];
  frame_from_normal(normal, frame);
  for (int c = 0; c < 3; c++) pos[c] = T(0.5) * (pa[c] + pb[c]);
  if (lane == 0) emit_contact(s, ncon, dropped, A.geom, B.geom, frame, pos, -depth);
  ncon = wshfl(ncon, 0); dropped = wshfl(dropped, 0);
  __syncwarp();
}

template <typename T>
__device__ void collide_plane(const SceneModel<T> &sm, NarrowScratch<T> &s, const Shape<T> &P, const Shape<T> &B, int &ncon, int &dropped, int lane) {
  CollideScratch<T> &cs = s.col;
  const T n[3] = {P.mat[2], P.mat[5], P.mat[8]}, nn[3] = {-n[0], -n[1], -n[2]};
  T sp[3];
  support(sm, B, nn, sp, lane);
  const T off = dot3(n, P.pos), depth = off - dot3(sp, n);
  if (!(depth > T(0))) return;
  T frame[9];
  frame_from_normal(n, frame);
  const T *t1 = frame + 3, *t2 = frame + 6;
  int nb = feature(sm, cs, B, nn, t1, t2, depth + T(1e-7), cs.FB, lane);
  if (lane == 0) {
    for (int i = 0; i < nb; i++) { cs.FB[i].h = -cs.FB[i].h; cs.mdist[i] = cs.FB[i].h - off; }
    nb = reduce_manifold(cs.FB, cs.mdist, nb);
    for (int i = 0; i < nb; i++) {
      if (cs.mdist[i] >= T(0)) continue;
      T pos[3];
      for (int c = 0; c < 3; c++) pos[c] = cs.FB[i].x * t1[c] + cs.FB[i].y * t2[c] + (cs.FB[i].h - T(0.5) * cs.mdist[i]) * frame[c];
      emit_contact(s, ncon, dropped, P.geom, B.geom, frame, pos, cs.mdist[i]);
    }
  }
  ncon = wshfl(ncon, 0); dropped = wshfl(dropped, 0);
  __syncwarp();
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
template <typename T>
__device__ __forceinline__ void geom_pose(const SceneModel<T> &sm, const Scratch<T> &s, int g, T *pos, T *mat) {
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
template <typename T>
__device__ __forceinline__ void geom_obb(const SceneModel<T> &sm, const Scratch<T> &s, int g, T *pos, T *mat, T *half) {
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
template <typename T>
__device__ void scene_broadphase(const SceneModel<T> &sm, Scratch<T> &s, const PipeBuf<T> &pb, int env, int sub, int &dropped, int lane) {
  BroadScratch<T> &cs = s.u.broad;
  PROF_START(s);
  // publish the body poses for the narrow-phase kernel
  {
    T *gx = pb.xpos + (size_t)env * (NSLOT * 3), *gm = pb.xmat + (size_t)env * (NSLOT * 9);
    for (int i = lane; i < NSLOT * 3; i += 32) gx[i] = (&s.xpos[0][0])[i];
    for (int i = lane; i < NSLOT * 9; i += 32) gm[i] = (&s.xmat[0][0])[i];
  }
  // world bounding-sphere centres of all geoms
  for (int g = lane; g < sm.ngeom; g += 32) {
    const int slot = sm.geom_slot[g];
    const T bc[3] = {sm.geom_bcenter[3 * g], sm.geom_bcenter[3 * g + 1], sm.geom_bcenter[3 * g + 2]};
    if (slot < 0) { cs.gcenter[0][g] = bc[0]; cs.gcenter[1][g] = bc[1]; cs.gcenter[2][g] = bc[2]; }
    else {
      T t[3];
      mulmv(t, s.xmat[slot], bc);
      for (int c = 0; c < 3; c++) cs.gcenter[c][g] = s.xpos[slot][c] + t[c];
    }
  }
  __syncwarp();
  int npq = 0;
  for (int p = 0; p < sm.npair; p++) {  // uniform loop; broad-phase order == oracle order
    const int b1 = sm.bodypair[2 * p], b2 = sm.bodypair[2 * p + 1];
    if (b1 != 0) {
      T c1[3], c2[3], t[3];
      const int s1 = sm.body_slot[b1], s2 = sm.body_slot[b2];
      for (int k = 0; k < 2; k++) {
        const int b = k ? b2 : b1, sl = k ? s2 : s1;
        T *c = k ? c2 : c1;
        const T bc[3] = {sm.body_bcenter[3 * b], sm.body_bcenter[3 * b + 1], sm.body_bcenter[3 * b + 2]};
        if (sl < 0) { c[0] = bc[0]; c[1] = bc[1]; c[2] = bc[2]; }
        else { mulmv(t, s.xmat[sl], bc); for (int e = 0; e < 3; e++) c[e] = s.xpos[sl][e] + t[e]; }
      }
      sub3(t, c1, c2);
      const T r = sm.body_rbound[b1] + sm.body_rbound[b2];
      if (dot3(t, t) > r * r) continue;
    }
    const int a1 = sm.body_geomadr[b1], n1 = sm.body_geomnum[b1], a2 = sm.body_geomadr[b2], n2 = sm.body_geomnum[b2];
    const int total = n1 * n2;
    for (int base = 0; base < total; base += 32) {
      const int k = base + lane;
      bool keep = false;
      int g1 = 0, g2 = 0;
      if (k < total) {
        g1 = a1 + k / n2; g2 = a2 + k % n2;
        const int ty1 = sm.geom_type[g1];
        const T cA[3] = {cs.gcenter[0][g1], cs.gcenter[1][g1], cs.gcenter[2][g1]}, cB[3] = {cs.gcenter[0][g2], cs.gcenter[1][g2], cs.gcenter[2][g2]};
        const T rA = sm.geom_rbound[g1], rB = sm.geom_rbound[g2];
        if (ty1 == G_PLANE) {
          const T n[3] = {sm.geom_mat[9 * g1 + 2], sm.geom_mat[9 * g1 + 5], sm.geom_mat[9 * g1 + 8]};
          const T pp[3] = {sm.geom_pos[3 * g1], sm.geom_pos[3 * g1 + 1], sm.geom_pos[3 * g1 + 2]};
          keep = !(dot3(n, cB) - dot3(n, pp) - rB > T(0));
          if (keep) {
            T pos[3], mat[9], half[3];
            geom_obb(sm, s, g2, pos, mat, half);
            const T ext = t_abs(n[0] * mat[0] + n[1] * mat[3] + n[2] * mat[6]) * half[0] + t_abs(n[0] * mat[1] + n[1] * mat[4] + n[2] * mat[7]) * half[1] +
                          t_abs(n[0] * mat[2] + n[1] * mat[5] + n[2] * mat[8]) * half[2];
            keep = !(dot3(n, pos) - dot3(n, pp) - ext > T(0));
          }
        } else {
          T t[3];
          sub3(t, cA, cB);
          const T r = rA + rB;
          keep = !(dot3(t, t) > r * r);
          if (keep) {
            T p1[3], m1[9], h1[3], p2[3], m2[9], h2[3];
            geom_obb(sm, s, g1, p1, m1, h1); geom_obb(sm, s, g2, p2, m2, h2);
            keep = obb_overlap(p1, m1, h1, p2, m2, h2);
          }
        }
      }
      const unsigned m = __ballot_sync(FULL, keep);
      const int idx = npq + __popc(m & ((1u << lane) - 1));
      if (keep && idx < PAIRCAP) cs.pairq[idx] = (unsigned)g1 | ((unsigned)g2 << 8) | ((unsigned)idx << 16);
      npq += __popc(m);
    }
  }
  if (npq > PAIRCAP) { dropped += npq - PAIRCAP; npq = PAIRCAP; }
  __syncwarp();
  // append (env, g1 | g2 << 8 | pair index << 16) to the work list of this substep
  int base = 0;
  if (lane == 0 && npq > 0) base = atomicAdd(pb.nwork + 2 * sub, npq);
  base = wshfl(base, 0);
  for (int i = lane; i < npq; i += 32) pb.work[(size_t)base + i] = make_uint2((unsigned)env, cs.pairq[i]);
  PROF_ACC(s, P_BROAD, lane);
  PROF_CNT(s, P_NPQ, npq, lane);
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------ constraint rows (lane per contact)
template <typename T>
__device__ void build_rows(const SceneModel<T> &sm, const ArmModelT<T> &am, Scratch<T> &s, int &dropped, int lane) {
  SolveScratch<T> &R = s.u.sol;
  const int ncon = s.ncon;
  // NOTE: contact geometry lives outside the union; the collision scratch is dead from here on.
  int blk_base = 0;
  for (int c0 = 0; c0 < ncon; c0 += 32) {
    const int c = c0 + lane;
    const bool valid = c < ncon;
    // allocate Jacobian blocks: one per distinct dynamic body (dof base) of the contact, in contact order
    int nb = 0;
    if (valid) {
      const int t1 = sm.body_slot[sm.geom_body[s.c_g1[c]]], t2 = sm.body_slot[sm.geom_body[s.c_g2[c]]];
      const int a1 = t1 < 0 ? 31 : (t1 < NJ ? 0 : NJ + 6 * (t1 - NJ)), a2 = t2 < 0 ? 31 : (t2 < NJ ? 0 : NJ + 6 * (t2 - NJ));
      nb = (a1 != 31) + (a2 != 31 && a2 != a1);
    }
    int incl = nb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
    const int my_blk = blk_base + incl - nb;
    blk_base += wshfl(incl, 31);
    const bool fits = my_blk + nb <= NBLK;
    dropped += __popc(__ballot_sync(FULL, valid && !fits));
    if (!valid) continue;
    const int g1 = s.c_g1[c], g2 = s.c_g2[c];
    const int b1 = sm.geom_body[g1], b2 = sm.geom_body[g2];
    const int s1 = sm.body_slot[b1], s2 = sm.body_slot[b2];
    // [upstream] mj_contactParam
    const int dim = max(sm.geom_condim[g1], sm.geom_condim[g2]);
    const int p1 = sm.geom_priority[g1], p2 = sm.geom_priority[g2];
    T f[3], mix;
    if (p1 == p2) {
      for (int k = 0; k < 3; k++) f[k] = max(sm.geom_friction[3 * g1 + k], sm.geom_friction[3 * g2 + k]);
      const T m1 = sm.geom_solmix[g1], m2 = sm.geom_solmix[g2];
      if (m1 >= T(1e-15) && m2 >= T(1e-15)) mix = m1 / (m1 + m2);
      else if (m1 < T(1e-15) && m2 < T(1e-15)) mix = T(0.5);
      else mix = m1 < T(1e-15) ? T(0) : T(1);
    } else {
      const int g = p1 > p2 ? g1 : g2;
      for (int k = 0; k < 3; k++) f[k] = sm.geom_friction[3 * g + k];
      mix = p1 > p2 ? T(1) : T(0);
    }
    T solref[2], solimp[5];
    const T *r1 = sm.geom_solref + 2 * g1, *r2 = sm.geom_solref + 2 * g2;
    if (r1[0] > T(0) && r2[0] > T(0)) for (int k = 0; k < 2; k++) solref[k] = mix * r1[k] + (T(1) - mix) * r2[k];
    else for (int k = 0; k < 2; k++) solref[k] = min(r1[k], r2[k]);
    for (int k = 0; k < 5; k++) solimp[k] = mix * sm.geom_solimp[5 * g1 + k] + (T(1) - mix) * sm.geom_solimp[5 * g2 + k];
    const T margin = max(sm.geom_margin[g1], sm.geom_margin[g2]) - max(sm.geom_gap[g1], sm.geom_gap[g2]);
    const T dist = s.c_dist[c];
    // impedance, reference acceleration gains, regulariser
    const T imp = impedance(solimp, dist, margin);
    const T dmax = t_clamp(solimp[1], T(1e-4), T(0.9999));
    T K, B;
    if (solref[0] > T(0)) {
      const T tc = solref[0] > T(2) * sm.timestep ? solref[0] : T(2) * sm.timestep;
      K = T(1) / (dmax * dmax * tc * tc * solref[1] * solref[1]); B = T(2) / (dmax * tc);
    } else { K = -solref[0] / (dmax * dmax); B = -solref[1] / dmax; }
    const T tran = sm.body_invweight0[2 * b1] + sm.body_invweight0[2 * b2];
    T R0 = (T(1) - imp) / imp * tran;
    R0 = R0 > T(1e-15) ? R0 : T(1e-15);
    R.D0[c] = T(1) / R0;
    R.fri[0][c] = f[0]; R.fri[1][c] = f[1]; R.fri[2][c] = f[2];
    R.mu[c] = dim > 1 ? f[0] * t_sqrt(T(1) / (sm.impratio > T(1e-15) ? sm.impratio : T(1e-15))) : T(0);
    // Jacobian blocks
    const T pos[3] = {s.c_pos[0][c], s.c_pos[1][c], s.c_pos[2][c]};
    T fr[9];
    for (int e = 0; e < 9; e++) fr[e] = s.c_frame[e][c];
    int baseA = 31, baseB = 31, bA = -1, bB = -1;
    if (fits) {
      for (int k = 0; k < nb; k++)
        for (int e = 0; e < 36; e++) R.J[e][my_blk + k] = T(0);
      for (int side = 0; side < 2; side++) {
        const int sl = side ? s2 : s1;
        if (sl < 0) continue;
        const T sgn = side ? T(1) : T(-1);
        const int base = sl < NJ ? 0 : NJ + 6 * (sl - NJ);
        int bi;
        if (baseA == 31 || baseA == base) { baseA = base; bA = my_blk; bi = bA; }
        else { baseB = base; bB = my_blk + 1; bi = bB; }
        for (int col = 0; col < 6; col++) {
          T tr[3] = {T(0), T(0), T(0)}, ro[3] = {T(0), T(0), T(0)};
          if (sl < NJ) {
            if (col > sl) continue;
            const T a[3] = {s.arm_a[col][0], s.arm_a[col][1], s.arm_a[col][2]};
            const T r[3] = {pos[0] - s.arm_p[col][0], pos[1] - s.arm_p[col][1], pos[2] - s.arm_p[col][2]};
            cross3(tr, a, r);
            ro[0] = a[0]; ro[1] = a[1]; ro[2] = a[2];
          } else if (col < 3) {
            tr[col] = T(1);
          } else {
            const T *Rm = s.xmat[sl];
            const T a[3] = {Rm[col - 3], Rm[3 + col - 3], Rm[6 + col - 3]};
            const T r[3] = {pos[0] - s.xpos[sl][0], pos[1] - s.xpos[sl][1], pos[2] - s.xpos[sl][2]};
            cross3(tr, a, r);
            ro[0] = a[0]; ro[1] = a[1]; ro[2] = a[2];
          }
          for (int r = 0; r < dim; r++) {
            const T *ax = fr + 3 * (r % 3);
            const T v = sgn * (r < 3 ? dot3(ax, tr) : dot3(ax, ro));
            R.J[r * 6 + col][bi] += v;
          }
        }
      }
    }
    const int edim = fits ? dim : 0;  // a contact whose blocks do not fit the pool is dropped (counted by the caller)
    R.info[c] = edim | (baseA << 8) | (baseB << 16);
    R.blk[0][c] = bA; R.blk[1][c] = bB;
    // aref = -B vel - K imp (dist - margin) on the normal row; friction rows: -B vel
    for (int r = 0; r < 6; r++) {
      T vel = T(0);
      if (r < edim) {
        for (int col = 0; col < 6; col++) {
          if (bA >= 0) vel += R.J[r * 6 + col][bA] * s.qd[baseA + col];
          if (bB >= 0) vel += R.J[r * 6 + col][bB] * s.qd[baseB + col];
        }
      }
      R.aref[r][c] = r < edim ? (-B * vel - (r == 0 ? K * imp * (dist - margin) : T(0))) : T(0);
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------ elliptic cone (per lane)
template <typename T>
struct Cone {
  int dim;
  T mu, S[6], D[6];
};
template <typename T>
__device__ __forceinline__ void cone_setup(const SolveScratch<T> &R, int c, T impratio, Cone<T> &k) {
  k.dim = R.info[c] & 0xff;
  k.mu = R.mu[c];
  const T f0 = R.fri[0][c], ft = R.fri[1][c], fr = R.fri[2][c], D0 = R.D0[c];
  const T D1 = D0 * (impratio > T(1e-15) ? impratio : T(1e-15));
  k.S[0] = k.mu; k.S[1] = f0; k.S[2] = f0; k.S[3] = ft; k.S[4] = fr; k.S[5] = fr;
  k.D[0] = D0; k.D[1] = D1; k.D[2] = D1; k.D[3] = D1 * ft * ft / (f0 * f0); k.D[4] = D1 * fr * fr / (f0 * f0); k.D[5] = k.D[4];
}
// zone 0: satisfied; 1: quadratic (bottom); 2: cone surface.  force = -d cost / d jar.  Hessian = v1 v1^T - v2 v2^T + diag(e).
template <typename T>
__device__ __forceinline__ int cone_eval(const Cone<T> &k, const T *x, T &cost, T *force, T *v1, T *v2, T *e) {
  const int dim = k.dim;
  cost = T(0);
#pragma unroll
  for (int j = 0; j < 6; j++) { force[j] = T(0); v1[j] = T(0); v2[j] = T(0); e[j] = T(0); }
  if (dim == 1) {
    if (x[0] >= T(0)) return 0;
    cost = T(0.5) * k.D[0] * x[0] * x[0]; force[0] = -k.D[0] * x[0]; e[0] = k.D[0];
    return 1;
  }
  T U[6], T2 = T(0);
  U[0] = x[0] * k.mu;
#pragma unroll
  for (int j = 1; j < 6; j++) { U[j] = j < dim ? x[j] * k.S[j] : T(0); T2 += U[j] * U[j]; }
  const T N = U[0], Tt = t_sqrt(T2), mu = k.mu;
  if (N >= mu * Tt || (Tt <= T(0) && N >= T(0))) return 0;
  if (mu * N + Tt <= T(0) || (Tt <= T(0) && N < T(0))) {
#pragma unroll
    for (int j = 0; j < 6; j++)
      if (j < dim) { cost += T(0.5) * k.D[j] * x[j] * x[j]; force[j] = -k.D[j] * x[j]; e[j] = k.D[j]; }
    return 1;
  }
  const T Dm = k.D[0] / (mu * mu * (T(1) + mu * mu)), NmT = N - mu * Tt;
  cost = T(0.5) * Dm * NmT * NmT;
  force[0] = -Dm * NmT * mu;
  const T sDm = t_sqrt(Dm), c2 = -Dm * NmT * mu;  // c2 > 0 in the middle zone
  const T s2 = t_sqrt(c2 / (Tt * Tt * Tt));
  v1[0] = sDm * mu;  // S0 * g0, g0 = 1
#pragma unroll
  for (int j = 1; j < 6; j++)
    if (j < dim) {
      force[j] = -force[0] / Tt * U[j] * k.S[j];
      v1[j] = sDm * k.S[j] * (-mu * U[j] / Tt);
      v2[j] = s2 * k.S[j] * U[j];
      e[j] = c2 / Tt * k.S[j] * k.S[j];
    }
  return 2;
}

// ------------------------------------------------------------------------------------------------ Newton solver (warp)
template <typename T>
__device__ __forceinline__ T blockdiag_mv(const Scratch<T> &s, const T *x, int i) {  // (M x)_i for i < NV
  T acc = T(0);
  if (i < NJ) {
    for (int j = 0; j < NJ; j++) acc += (j <= i ? s.Marm[tri(i, j)] : s.Marm[tri(j, i)]) * x[j];
  } else {
    const int p = (i - NJ) / 6, li = (i - NJ) % 6;
    for (int j = 0; j < 6; j++) acc += (j <= li ? s.Mprop[p][tri(li, j)] : s.Mprop[p][tri(j, li)]) * x[NJ + 6 * p + j];
  }
  return acc;
}

// J_c x for the lane's contact: out[r] = sum_col J[r][col] x[base + col] over the contact's (<= 2) blocks
template <typename T>
__device__ __forceinline__ void contact_Jx(const SolveScratch<T> &R, int dim, int baseA, int baseB, int bA, int bB, const T *x, T *out) {
#pragma unroll
  for (int r = 0; r < 6; r++) {
    T acc = T(0);
    if (r < dim) {
#pragma unroll
      for (int col = 0; col < 6; col++) {
        if (bA >= 0) acc += R.J[r * 6 + col][bA] * x[baseA + col];
        if (bB >= 0) acc += R.J[r * 6 + col][bB] * x[baseB + col];
      }
    }
    out[r] = acc;
  }
}

template <typename T>
__device__ void cholesky_packed(T *H, int lane) {  // in-place lower Cholesky of the packed NV x NV matrix, lanes = rows
  for (int j = 0; j < NV; j++) {
    T sacc = T(0);
    if (lane >= j && lane < NV) {
      sacc = H[tri(lane, j)];
      for (int k = 0; k < j; k++) sacc -= H[tri(lane, k)] * H[tri(j, k)];
    }
    T dg = wshfl(sacc, j);
    dg = t_sqrt(dg > T(1e-15) ? dg : T(1e-15));
    if (lane >= j && lane < NV) H[tri(lane, j)] = lane == j ? dg : sacc / dg;
    __syncwarp();
  }
}
template <typename T>
__device__ T chol_solve_packed(const T *L, T b, int lane) {  // lane i holds b_i; returns x_i
  for (int k = 0; k < NV; k++) {
    const T yk = wshfl(b, k) / L[tri(k, k)];
    if (lane == k) b = yk;
    else if (lane > k && lane < NV) b -= L[tri(lane, k)] * yk;
  }
  for (int k = NV - 1; k >= 0; k--) {
    const T xk = wshfl(b, k) / L[tri(k, k)];
    if (lane == k) b = xk;
    else if (lane < k) b -= L[tri(k, lane)] * xk;
  }
  return b;
}

template <typename T>
__device__ int scene_solve(const SceneModel<T> &sm, const ArmModelT<T> &am, Scratch<T> &s, const ArmRows<T> &arows, int max_iter, T tol, int lane) {
  SolveScratch<T> &R = s.u.sol;
  const int ncon = s.ncon;
  const T xeps = sizeof(T) == 8 ? T(1e-14) : T(2e-6);
  // per-lane contact cache
  Cone<T> cone[CSL];
  T jar0[CSL][6];
  int cdim[CSL], cA[CSL], cB[CSL], kA[CSL], kB[CSL];
#pragma unroll
  for (int k = 0; k < CSL; k++) {
    const int c = lane + 32 * k;
    cdim[k] = 0; cA[k] = 31; cB[k] = 31; kA[k] = -1; kB[k] = -1;
    if (c < ncon) {
      cone_setup(R, c, sm.impratio, cone[k]);
      cdim[k] = cone[k].dim; cA[k] = (R.info[c] >> 8) & 0xff; cB[k] = (R.info[c] >> 16) & 0xff;
      kA[k] = R.blk[0][c]; kB[k] = R.blk[1][c];
      contact_Jx(R, cdim[k], cA[k], cB[k], kA[k], kB[k], s.qacc_s, jar0[k]);
#pragma unroll
      for (int r = 0; r < 6; r++) jar0[k][r] -= R.aref[r][c];
    }
  }
  // cost of the contact + arm rows at x (+ alpha * dx): returns warp-uniform (cost, d/dalpha, d2/dalpha2)
  auto rows_line = [&](const T *x, const T *dx, T alpha, T &c, T &g, T &h) {
    PROF_CNT(s, P_LINE, 1, lane);
    T lc = T(0), lg = T(0), lh = T(0);
#pragma unroll
    for (int k = 0; k < CSL; k++) {
      const int ci = lane + 32 * k;
      if (ci < ncon) {
        T jx[6], jv[6], xx[6], force[6], v1[6], v2[6], e[6], cc;
        contact_Jx(R, cdim[k], cA[k], cB[k], kA[k], kB[k], x, jx);
        contact_Jx(R, cdim[k], cA[k], cB[k], kA[k], kB[k], dx, jv);
#pragma unroll
        for (int r = 0; r < 6; r++) xx[r] = jar0[k][r] + jx[r] + alpha * jv[r];
        cone_eval(cone[k], xx, cc, force, v1, v2, e);
        T a1 = T(0), a2 = T(0);
        lc += cc;
#pragma unroll
        for (int r = 0; r < 6; r++) { lg -= force[r] * jv[r]; a1 += v1[r] * jv[r]; a2 += v2[r] * jv[r]; lh += e[r] * jv[r] * jv[r]; }
        lh += a1 * a1 - a2 * a2;
      }
    }
    lc = warp_sum(lc); lg = warp_sum(lg); lh = warp_sum(lh);
    // arm friction / limit rows (uniform)
    T xa[NJ], da[NJ];
#pragma unroll
    for (int i = 0; i < NJ; i++) { xa[i] = x[i]; da[i] = dx[i]; }
    arm_rows_line(am, arows, xa, da, alpha, lc, lg, lh);
    c += lc; g += lg; h += lh;
  };
  const T *zero = s.Ms;  // scratch vector, zeroed here
  if (lane < NV) s.Ms[lane] = T(0);
  __syncwarp();
  // warm start: keep delta only if it beats delta = 0
  {
    T c0 = T(0), cw = T(0), g = T(0), h = T(0);
    rows_line(zero, zero, T(0), c0, g, h);
    T md = lane < NV ? blockdiag_mv(s, s.delta, lane) * s.delta[lane] : T(0);
    cw = T(0.5) * warp_sum(md);
    rows_line(s.delta, zero, T(0), cw, g, h);
    if (!(cw < c0)) { if (lane < NV) s.delta[lane] = T(0); }
    __syncwarp();
  }
  int iter = 0;
  for (; iter < max_iter; iter++) {
    // gradient and the factored per-contact Hessians
    const T mdl = lane < NV ? blockdiag_mv(s, s.delta, lane) : T(0);
    T gl = mdl;  // lane i < NV accumulates grad_i
    T frc[CSL][6];
#pragma unroll
    for (int k = 0; k < CSL; k++) {
      const int ci = lane + 32 * k;
#pragma unroll
      for (int r = 0; r < 6; r++) frc[k][r] = T(0);
      if (ci < ncon) {
        T jx[6], xx[6], v1[6], v2[6], e[6], cc;
        contact_Jx(R, cdim[k], cA[k], cB[k], kA[k], kB[k], s.delta, jx);
#pragma unroll
        for (int r = 0; r < 6; r++) xx[r] = jar0[k][r] + jx[r];
        cone_eval(cone[k], xx, cc, frc[k], v1, v2, e);
        // w = J^T v per block
#pragma unroll
        for (int side = 0; side < 2; side++) {
          const int bi = side ? kB[k] : kA[k];
          if (bi < 0) continue;
#pragma unroll
          for (int col = 0; col < 6; col++) {
            T a1 = T(0), a2 = T(0);
#pragma unroll
            for (int r = 0; r < 6; r++) { const T j = R.J[r * 6 + col][bi]; a1 += j * v1[r]; a2 += j * v2[r]; }
            R.w1[col][bi] = a1; R.w2[col][bi] = a2;
          }
        }
#pragma unroll
        for (int r = 0; r < 6; r++) R.e[r][ci] = e[r];
      }
    }
    // grad -= J^T force: one warp reduction per dof
    for (int dof = 0; dof < NV; dof++) {
      T v = T(0);
#pragma unroll
      for (int k = 0; k < CSL; k++) {
        const int ci = lane + 32 * k;
        if (ci < ncon) {
          int col = -1, bi = -1;
          if (kA[k] >= 0 && dof >= cA[k] && dof < cA[k] + 6) { col = dof - cA[k]; bi = kA[k]; }
          else if (kB[k] >= 0 && dof >= cB[k] && dof < cB[k] + 6) { col = dof - cB[k]; bi = kB[k]; }
          if (col >= 0) {
#pragma unroll
            for (int r = 0; r < 6; r++) v += R.J[r * 6 + col][bi] * frc[k][r];
          }
        }
      }
      v = warp_sum(v);
      if (lane == dof) gl -= v;
    }
    // arm rows: force and diagonal Hessian (uniform)
    T hdiag = T(0);
    if (lane < NJ) {
      const int i = lane;
      T f = T(0);
      const T x = arows.jar0_f[i] + s.delta[i], eta = am.frictionloss[i], rf = am.fr_R[i] * eta;
      if (x <= -rf) f = eta;
      else if (x >= rf) f = -eta;
      else { f = -am.fr_D[i] * x; hdiag = am.fr_D[i]; }
      if (arows.D_l[i] > T(0)) {
        const T xl = arows.jar0_l[i] + arows.js[i] * s.delta[i];
        if (xl < T(0)) { f += arows.js[i] * (-arows.D_l[i] * xl); hdiag += arows.D_l[i]; }
      }
      gl -= f;
    }
    if (lane < NV) { s.grad[lane] = gl; s.Md[lane] = mdl; }
    const T gn = t_sqrt(warp_sum(lane < NV ? gl * gl : T(0)));
    __syncwarp();
    if (am.solver_scale * gn < tol) break;
    // Hessian: lane per packed entry
    for (int en = lane; en < NH; en += 32) {
      // unpack (i, j), i >= j
      int i = 0;
      while ((i + 1) * (i + 2) / 2 <= en) i++;
      const int j = en - i * (i + 1) / 2;
      T acc = T(0);
      if (i < NJ) acc = s.Marm[tri(i, j)];
      else if ((i - NJ) / 6 == (j - NJ) / 6 && j >= NJ) acc = s.Mprop[(i - NJ) / 6][tri((i - NJ) % 6, (j - NJ) % 6)];
      for (int c = 0; c < ncon; c++) {
        const int inf = R.info[c], bA = (inf >> 8) & 0xff, bB = (inf >> 16) & 0xff, pA = R.blk[0][c], pB = R.blk[1][c];
        int ci = -1, cj = -1, bi = -1, bj = -1;
        if (pA >= 0 && i >= bA && i < bA + 6) { ci = i - bA; bi = pA; } else if (pB >= 0 && i >= bB && i < bB + 6) { ci = i - bB; bi = pB; }
        if (pA >= 0 && j >= bA && j < bA + 6) { cj = j - bA; bj = pA; } else if (pB >= 0 && j >= bB && j < bB + 6) { cj = j - bB; bj = pB; }
        if (ci < 0 || cj < 0) continue;
        T a = R.w1[ci][bi] * R.w1[cj][bj] - R.w2[ci][bi] * R.w2[cj][bj];
#pragma unroll
        for (int r = 0; r < 6; r++) a += R.e[r][c] * R.J[r * 6 + ci][bi] * R.J[r * 6 + cj][bj];
        acc += a;
      }
      s.H[en] = acc;
    }
    __syncwarp();
    if (lane < NJ) s.H[tri(lane, lane)] += hdiag;
    __syncwarp();
    cholesky_packed(s.H, lane);
    const T sr = chol_solve_packed(s.H, lane < NV ? -gl : T(0), lane);
    if (lane < NV) s.search[lane] = sr;
    __syncwarp();
    const T msl = lane < NV ? blockdiag_mv(s, s.search, lane) : T(0);
    const T q0 = T(0.5) * warp_sum(lane < NV ? s.delta[lane] * mdl : T(0));
    const T q1 = warp_sum(lane < NV ? sr * mdl : T(0));
    const T q2 = warp_sum(lane < NV ? sr * msl : T(0));
    T f0 = q0, df0 = q1, ddf0 = q2;
    rows_line(s.delta, s.search, T(0), f0, df0, ddf0);
    if (df0 >= T(0) || ddf0 <= T(0)) break;
    T alpha = -df0 / ddf0, lo = T(0), hi = T(-1), f = f0;
    for (int ls = 0; ls < 30; ls++) {
      T df = q1 + alpha * q2, ddf = q2;
      f = q0 + alpha * q1 + T(0.5) * alpha * alpha * q2;
      rows_line(s.delta, s.search, alpha, f, df, ddf);
      if (t_abs(df) <= T(sizeof(T) == 8 ? 1e-13 : 1e-6) * t_abs(df0)) break;
      if (df < T(0)) lo = alpha; else hi = alpha;
      T next = alpha - df / ddf;
      if (hi > T(0) && (next <= lo || next >= hi)) next = T(0.5) * (lo + hi);
      else if (hi < T(0) && next <= lo) next = T(2) * alpha;
      if (next == alpha || t_abs(next - alpha) <= xeps * t_abs(alpha) || (hi > T(0) && hi - lo <= xeps * hi)) { alpha = next; break; }
      alpha = next;
    }
    T st = T(0), am_ = T(1);
    if (lane < NV) {
      st = alpha * sr;
      s.delta[lane] += st;
      am_ = t_abs(s.qacc_s[lane]) + t_abs(s.delta[lane]);
      st = t_abs(st);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { st = max(st, __shfl_xor_sync(FULL, st, o)); am_ = max(am_, __shfl_xor_sync(FULL, am_, o)); }
    __syncwarp();
    if (am.solver_scale * (f0 - f) < tol || st <= xeps * max(am_, T(1))) { iter++; break; }
  }
  return iter;
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
template <typename T>
__device__ float scene_reward(const SceneModel<T> &sm, const Scratch<T> &s) {
  // success_detector_utils.py:22-28 — linear velocity of either prop >= 1e-3 -> 0
  for (int p = 0; p < NPROP; p++) {
    T mx = T(0);
    for (int c = 0; c < 3; c++) mx = max(mx, t_abs(s.qd[NJ + 6 * p + c]));
    if (mx >= T(1e-3)) return 0.f;
  }
  // object OOBB: root BVH box at xipos / ximat (oobb_utils.py:137-148,165-172)
  const T *R0 = s.xmat[NJ], *X0 = s.xpos[NJ];
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
  const T *qb = s.q + NJ + 7 + 3;
  T q1[4] = {qb[0], qb[1], qb[2], qb[3]}, p1[3];
  const T n = t_sqrt(q1[0] * q1[0] + q1[1] * q1[1] + q1[2] * q1[2] + q1[3] * q1[3]);
  for (int i = 0; i < 4; i++) q1[i] /= n;
  mulmv(t, s.xmat[NJ + 1], sm.reward_box_pos);
  for (int c = 0; c < 3; c++) p1[c] = t[c] + s.xpos[NJ + 1][c];
  return overlap_oobb_oobb(p0, q0, sm.reward_obj_box + 3, p1, q1, sm.reward_box_half) ? 1.f : 0.f;
}

// ------------------------------------------------------------------------------------------------ task layer: observations
template <typename T>
__device__ void write_obs_scene(const StepCfg &cfg, const EnvState<T> &S, const so101_step_out &out, const Scratch<T> &s, int env, int t, float reward,
                                float discount, uint8_t st, int lane) {
  constexpr int SD = NQ + NV;
  const int dj = cfg.dj + 1, dp = cfg.dp + 1;
  const size_t N = S.N;
  float *rj = S.ring_joints + ((size_t)(t % dj) * N + env) * 6;
  float *rp = S.ring_phys + ((size_t)(t % dp) * N + env) * SD;
  const int tj = t - cfg.dj > 0 ? t - cfg.dj : 0, tp = t - cfg.dp > 0 ? t - cfg.dp : 0;
  const float *sj = S.ring_joints + ((size_t)(tj % dj) * N + env) * 6;
  const float *sp = S.ring_phys + ((size_t)(tp % dp) * N + env) * SD;
  for (int i = lane; i < SD; i += 32) {
    const float v = i < NQ ? (float)s.q[i] : (float)s.qd[i - NQ];
    rp[i] = v;
    if (out.physics_state) out.physics_state[(size_t)env * SD + i] = v;
    if (out.delayed_physics_state) out.delayed_physics_state[(size_t)env * SD + i] = tp == t ? v : sp[i];
  }
  if (lane < 6) {
    const float v = (float)s.q[lane];
    rj[lane] = v;
    if (out.undelayed_joints_pos) out.undelayed_joints_pos[(size_t)env * 6 + lane] = v;
    if (out.joints_pos) out.joints_pos[(size_t)env * 6 + lane] = tj == t ? v : sj[lane];
    if (out.commanded_joints_pos) out.commanded_joints_pos[(size_t)env * 6 + lane] = (float)s.ctrl[lane];
  }
  if (lane == 0) {
    if (out.reward) out.reward[env] = reward;
    if (out.discount) out.discount[env] = discount;
    if (out.step_type) out.step_type[env] = st;
  }
}

template <typename T>
__device__ void reset_env_scene(const StepCfg &cfg, const EnvState<T> &S, const so101_step_out &out, Scratch<T> &s, int env, int lane) {
  for (int i = lane; i < NQ; i += 32) { s.q[i] = S.init_qpos[(size_t)env * NQ + i]; S.qpos[(size_t)env * NQ + i] = s.q[i]; }
  for (int i = lane; i < NV; i += 32) { s.qd[i] = S.init_qvel[(size_t)env * NV + i]; S.qvel[(size_t)env * NV + i] = s.qd[i]; S.warm[(size_t)env * NV + i] = T(0); }
  if (lane < NJ) { s.ctrl[lane] = (T)cfg.home[lane] + (T)cfg.offsets[lane]; S.ctrl[(size_t)env * 6 + lane] = s.ctrl[lane]; }
  if (lane == 0) { S.step[env] = 0; S.needs_reset[env] = 0; }
  __syncwarp();
  write_obs_scene(cfg, S, out, s, env, 0, 0.f, 1.f, SO101_STEP_FIRST, lane);
}

// ------------------------------------------------------------------------------------------------ kernels
template <typename T>
__device__ __forceinline__ void prof_begin(Scratch<T> &s, const EnvState<T> &S, int lane) {
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

// Once per control step: dm_control auto-reset, action -> ctrl, kinematics and broad phase of substep 0.
template <typename T>
__global__ void __launch_bounds__(WARPS_SOLVE * 32) scene_begin_kernel(const __grid_constant__ ArmModelT<T> am, const __grid_constant__ SceneModel<T> sm,
                                                                      const __grid_constant__ StepCfg cfg, const EnvState<T> S, const PipeBuf<T> pb,
                                                                      const float *__restrict__ action, const so101_step_out out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Scratch<T> *all = reinterpret_cast<Scratch<T> *>(smem_raw);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int env = blockIdx.x * WARPS_SOLVE + wib;
  if (env >= S.N) return;
  Scratch<T> &s = all[wib];
  if (lane == 0) pb.ncon_raw[env] = 0;
  if (S.needs_reset[env]) {  // the step() after a LAST step resets and returns FIRST; no physics this call
    reset_env_scene(cfg, S, out, s, env, lane);
    if (lane == 0) pb.active[env] = 0;
    return;
  }
  prof_begin(s, S, lane);
  for (int i = lane; i < NQ; i += 32) s.q[i] = S.qpos[(size_t)env * NQ + i];
  if (lane < NJ) S.ctrl[(size_t)env * 6 + lane] = (T)action[(size_t)env * 6 + lane] + (T)cfg.offsets[lane];  // so100_task.py:266-287
  if (lane == 0) { pb.active[env] = 1; pb.flags[env] = 0; }
  __syncwarp();
  ArmKin<T> k;
  scene_kinematics(am, s, k, lane);
  int dropped = 0;
  scene_broadphase(sm, s, pb, env, 0, dropped, lane);
  if (lane == 0 && dropped) atomicAdd(S.diverged_count + 1, dropped);
  prof_flush(s, S, lane);
}

// Per substep: one warp per candidate geom pair, pulled from the work list with an atomic cursor (pairs differ 10x in
// cost: a GJK miss vs GJK + EPA + manifold).  Contacts go to the env's raw buffer tagged with (pair index, manifold index).
template <typename T>
__global__ void __launch_bounds__(WARPS_NARROW * 32) scene_narrow_kernel(const __grid_constant__ SceneModel<T> sm, const __grid_constant__ StepCfg cfg,
                                                                        const EnvState<T> S, const PipeBuf<T> pb, int sub) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  NarrowScratch<T> *all = reinterpret_cast<NarrowScratch<T> *>(smem_raw);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  NarrowScratch<T> &s = all[wib];
  const int nwork = pb.nwork[2 * sub];
  if (lane == 0) {
    s.profon = S.prof != nullptr; s.dbg = 0;
    for (int i = 0; i < 16; i++) s.prof[i] = 0;
  }
  __syncwarp();
  int dropped = 0;
  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(pb.nwork + 2 * sub + 1, 1);
    item = wshfl(item, 0);
    if (item >= nwork) break;
    const uint2 w = pb.work[item];
    const int env = (int)w.x, g1 = (int)(w.y & 0xff), g2 = (int)((w.y >> 8) & 0xff), pidx = (int)(w.y >> 16);
    const T(*xpos)[3] = reinterpret_cast<const T(*)[3]>(pb.xpos + (size_t)env * (NSLOT * 3));
    const T(*xmat)[9] = reinterpret_cast<const T(*)[9]>(pb.xmat + (size_t)env * (NSLOT * 9));
    if (cfg.dbg_env >= 0) {
      if (lane == 0) s.dbg = (env == cfg.dbg_env && S.step[env] == cfg.dbg_step);
      __syncwarp();
    }
    Shape<T> A, B;
    make_shape(sm, xpos, xmat, g1, A);
    make_shape(sm, xpos, xmat, g2, B);
    int ncon = 0;
    if (A.type == G_PLANE) { PROF_START(s); collide_plane(sm, s, A, B, ncon, dropped, lane); PROF_ACC(s, P_PLANE, lane); }
    else collide_convex(sm, s, A, B, ncon, dropped, lane);
    if (ncon > 0) {
      int base = 0;
      if (lane == 0) base = atomicAdd(pb.ncon_raw + env, ncon);
      base = wshfl(base, 0);
      if (lane < ncon) {
        if (base + lane < CONBUF) {
          T *dst = pb.con + ((size_t)env * CONBUF + base + lane) * 8;
          dst[0] = s.c_normal[0][lane]; dst[1] = s.c_normal[1][lane]; dst[2] = s.c_normal[2][lane];
          dst[3] = s.c_pos[0][lane]; dst[4] = s.c_pos[1][lane]; dst[5] = s.c_pos[2][lane];
          dst[6] = s.c_dist[lane];
          pb.con_key[(size_t)env * CONBUF + base + lane] = (pidx << 20) | (lane << 16) | (g1 << 8) | g2;
        }
      }
      __syncwarp();
    }
  }
  if (lane == 0 && dropped) atomicAdd(S.diverged_count + 1, dropped);
  prof_flush(s, S, lane);
}

// gather the env's raw contacts into shared memory in oracle order (pair order, then manifold order)
template <typename T>
__device__ void gather_contacts(const PipeBuf<T> &pb, Scratch<T> &s, int env, int &dropped, int lane) {
  int nraw = pb.ncon_raw[env];
  if (nraw > CONBUF) { dropped += nraw - CONBUF; nraw = CONBUF; }
  const int *keys = pb.con_key + (size_t)env * CONBUF;
  constexpr int RSL = (CONBUF + 31) / 32;
  int mykey[RSL], rank[RSL];
#pragma unroll
  for (int k = 0; k < RSL; k++) { const int i = lane + 32 * k; mykey[k] = i < nraw ? keys[i] : 0x7fffffff; rank[k] = 0; }
  // rank = number of contacts with a smaller key (keys are unique: pair index and manifold index)
#pragma unroll
  for (int kk = 0; kk < RSL; kk++) {
    for (int jl = 0; jl < 32; jl++) {
      const int j = 32 * kk + jl;
      if (j >= nraw) break;
      const int kj = wshfl(mykey[kk], jl);
#pragma unroll
      for (int k = 0; k < RSL; k++) rank[k] += kj < mykey[k];
    }
  }
  const int n = nraw < NCON ? nraw : NCON;
  if (nraw > NCON) dropped += nraw - NCON;
#pragma unroll
  for (int k = 0; k < RSL; k++) {
    const int i = lane + 32 * k;
    if (i < nraw && rank[k] < NCON) {
      const int c = rank[k];
      const T *src = pb.con + ((size_t)env * CONBUF + i) * 8;
      const T nrm[3] = {src[0], src[1], src[2]};
      T frame[9];
      frame_from_normal(nrm, frame);
#pragma unroll
      for (int e = 0; e < 9; e++) s.c_frame[e][c] = frame[e];
      s.c_pos[0][c] = src[3]; s.c_pos[1][c] = src[4]; s.c_pos[2][c] = src[5];
      s.c_dist[c] = src[6];
      s.c_g1[c] = (mykey[k] >> 8) & 0xff; s.c_g2[c] = mykey[k] & 0xff;
    }
  }
  if (lane == 0) { s.ncon = n; pb.ncon_raw[env] = 0; }
  __syncwarp();
}

// Per substep, warp per env: smooth dynamics, constraint rows from the gathered contacts, Newton, Euler; then either the
// kinematics + broad phase of the next substep or (last substep) the task layer.
template <typename T>
__global__ void __launch_bounds__(WARPS_SOLVE * 32) scene_solve_kernel(const __grid_constant__ ArmModelT<T> am, const __grid_constant__ SceneModel<T> sm,
                                                                      const __grid_constant__ StepCfg cfg, const EnvState<T> S, const PipeBuf<T> pb,
                                                                      const so101_step_out out, int sub) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Scratch<T> *all = reinterpret_cast<Scratch<T> *>(smem_raw);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int env = blockIdx.x * WARPS_SOLVE + wib;
  if (env >= S.N) return;
  if (!pb.active[env]) return;
  Scratch<T> &s = all[wib];
  prof_begin(s, S, lane);
  for (int i = lane; i < NQ; i += 32) s.q[i] = S.qpos[(size_t)env * NQ + i];
  for (int i = lane; i < NV; i += 32) { s.qd[i] = S.qvel[(size_t)env * NV + i]; s.warm[i] = S.warm[(size_t)env * NV + i]; }
  if (lane < NJ) s.ctrl[lane] = S.ctrl[(size_t)env * 6 + lane];
  if (lane == 0) s.dbg = 0;
  __syncwarp();
  int iters = 0, dropped = 0;
  const bool last = sub == cfg.nsub - 1;
  {
    PROF_START(s);
    ArmRows<T> arows;
    {
      ArmKin<T> k;
      scene_kinematics(am, s, k, lane);
      T qa[NJ], qda[NJ], ca[NJ], M[21], bias[NJ], frc[NJ], qs[NJ];
#pragma unroll
      for (int i = 0; i < NJ; i++) { qa[i] = s.q[i]; qda[i] = s.qd[i]; ca[i] = s.ctrl[i]; }
      arm_crb_rne(am, k, qda, M, bias);
      arm_actuation(am, qa, qda, ca, frc);
      T L[21];
#pragma unroll
      for (int i = 0; i < 21; i++) L[i] = M[i];
      chol6(L);
#pragma unroll
      for (int i = 0; i < NJ; i++) qs[i] = frc[i] - bias[i];
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 21; i++) s.Marm[i] = M[i];
#pragma unroll
        for (int i = 0; i < NJ; i++) s.fsm[i] = qs[i];
      }
      chol6_solve(L, qs);
      arm_make_rows(am, qa, qda, qs, arows);
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NJ; i++) s.qacc_s[i] = qs[i];
      }
    }
    if (lane < NPROP) {  // one lane per prop
      T M[21], bias[6], x[6];
      prop_dynamics(sm, am, s, lane, M, bias);
#pragma unroll
      for (int i = 0; i < 21; i++) s.Mprop[lane][i] = M[i];
#pragma unroll
      for (int i = 0; i < 6; i++) { x[i] = -bias[i]; s.fsm[NJ + 6 * lane + i] = x[i]; }
      chol6(M);
      chol6_solve(M, x);
#pragma unroll
      for (int i = 0; i < 6; i++) s.qacc_s[NJ + 6 * lane + i] = x[i];
    }
    __syncwarp();
    PROF_ACC(s, P_DYN, lane);
    gather_contacts(pb, s, env, dropped, lane);
    PROF_CNT(s, P_NCON, s.ncon, lane);
    if (S.dbg_contacts && last) {  // parity probe: contacts of the last substep
      float *dst = S.dbg_contacts + (size_t)env * (1 + 9 * NCON);
      if (lane == 0) dst[0] = (float)s.ncon;
      for (int c = lane; c < s.ncon; c += 32) {
        float *r = dst + 1 + 9 * c;
        r[0] = (float)s.c_g1[c]; r[1] = (float)s.c_g2[c]; r[2] = (float)s.c_dist[c];
        for (int e = 0; e < 3; e++) { r[3 + e] = (float)s.c_pos[e][c]; r[6 + e] = (float)s.c_frame[e][c]; }
      }
    }
    build_rows(sm, am, s, dropped, lane);
    PROF_ACC(s, P_ROWS, lane);
    if (lane < NV) s.delta[lane] = s.warm[lane] - s.qacc_s[lane];
    __syncwarp();
    iters = scene_solve(sm, am, s, arows, cfg.max_iter, (T)cfg.tol, lane);
    __syncwarp();
    PROF_ACC(s, P_SOLVE, lane);
    PROF_CNT(s, P_NEWTON, iters, lane);
    PROF_CNT(s, P_NSUB, 1, lane);
    // [upstream] mj_Euler
    T qacc = T(0);
    if (lane < NV) {
      qacc = s.qacc_s[lane] + s.delta[lane];
      s.warm[lane] = qacc;
      s.qd[lane] += sm.timestep * qacc;
    }
    const bool badnow = __any_sync(FULL, lane < NV && !(t_abs(qacc) < T(1e10)));
    if (badnow && lane == 0) pb.flags[env] = 1;
    __syncwarp();
    if (lane < NJ) s.q[lane] += sm.timestep * s.qd[lane];
    if (lane >= 8 && lane < 8 + NPROP) {
      const int p = lane - 8;
      T *qp = s.q + NJ + 7 * p;
      const T *v = s.qd + NJ + 6 * p;
      for (int c = 0; c < 3; c++) qp[c] += sm.timestep * v[c];
      const T w[3] = {v[3], v[4], v[5]};
      const T nw = t_sqrt(dot3(w, w)), ang = nw * sm.timestep;
      T qn[4] = {qp[3], qp[4], qp[5], qp[6]};
      if (ang > T(0)) {
        T sn, cn;
        t_sincos(T(0.5) * ang, &sn, &cn);
        const T qr[4] = {cn, w[0] / nw * sn, w[1] / nw * sn, w[2] / nw * sn};
        quat_mul(qn, qn, qr);
      }
      T n = t_sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
      if (n < T(1e-15)) { qn[0] = T(1); qn[1] = qn[2] = qn[3] = T(0); n = T(1); }
      for (int c = 0; c < 4; c++) qp[3 + c] = qn[c] / n;
    }
    __syncwarp();
    PROF_ACC(s, P_INTEG, lane);
  }
  for (int i = lane; i < NQ; i += 32) S.qpos[(size_t)env * NQ + i] = s.q[i];
  for (int i = lane; i < NV; i += 32) { S.qvel[(size_t)env * NV + i] = s.qd[i]; S.warm[(size_t)env * NV + i] = s.warm[i]; }
  // poses at the new state: next substep's collision, or (mj_step1 refresh) the task layer
  {
    ArmKin<T> k;
    scene_kinematics(am, s, k, lane);
  }
  if (!last) {
    scene_broadphase(sm, s, pb, env, sub + 1, dropped, lane);
  } else {
    PROF_START(s);
    const bool bad = pb.flags[env] != 0;
    const int t = S.step[env] + 1;
    float reward = scene_reward(sm, s), discount = 1.f;
    uint8_t st = (cfg.last_step > 0 && t >= cfg.last_step) ? SO101_STEP_LAST : SO101_STEP_MID;
    if (cfg.terminate_on_success && reward >= 1.f) { discount = 0.f; st = SO101_STEP_LAST; }  // so100_task.py:292-302
    if (bad) { reward = 0.f; discount = 0.f; st = SO101_STEP_LAST; }                         // task_suite.py:153
    if (lane == 0) {
      S.step[env] = t; S.needs_reset[env] = st == SO101_STEP_LAST; S.solver_iter[env] = iters; S.ncon[env] = s.ncon;
      if (bad) atomicAdd(S.diverged_count, 1);
    }
    write_obs_scene(cfg, S, out, s, env, t, reward, discount, st, lane);
    PROF_ACC(s, P_TASK, lane);
  }
  if (lane == 0 && dropped) atomicAdd(S.diverged_count + 1, dropped);
  prof_flush(s, S, lane);
}

template <typename T>
__global__ void scene_reset_kernel(const __grid_constant__ StepCfg cfg, const EnvState<T> S, const uint8_t *__restrict__ mask, const so101_step_out out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Scratch<T> *all = reinterpret_cast<Scratch<T> *>(smem_raw);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int env = blockIdx.x * WARPS_SOLVE + wib;
  if (env >= S.N) return;
  if (mask && !mask[env]) return;
  reset_env_scene(cfg, S, out, all[wib], env, lane);
}

template <typename T>
size_t scene_smem_bytes() {
  const size_t a = sizeof(Scratch<T>) * WARPS_SOLVE, b = sizeof(NarrowScratch<T>) * WARPS_NARROW;
  return a > b ? a : b;
}

template <typename T>
int scene_narrow_grid() {
  const size_t smem_nar = sizeof(NarrowScratch<T>) * WARPS_NARROW;
  cudaFuncSetAttribute(scene_narrow_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_nar);
  int nb = 0, dev = 0, sms = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, scene_narrow_kernel<T>, WARPS_NARROW * 32, smem_nar);
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (nb > 0 ? nb : 1) * (sms > 0 ? sms : 1);
}

// Launches of one control step: 1 memset + 1 + 2 * nsub kernels, all on the caller's stream.  Returns the kernel count.
template <typename T>
int launch_scene_step(const ArmModelT<T> &am, const SceneModel<T> &sm, const StepCfg &cfg, const EnvState<T> &S, const PipeBuf<T> &pb,
                      const float *action, const so101_step_out &out, cudaStream_t stream) {
  static bool configured = false;
  const size_t smem_env = sizeof(Scratch<T>) * WARPS_SOLVE, smem_nar = sizeof(NarrowScratch<T>) * WARPS_NARROW;
  if (!configured) {
    cudaFuncSetAttribute(scene_begin_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_env);
    cudaFuncSetAttribute(scene_solve_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_env);
    cudaFuncSetAttribute(scene_narrow_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_nar);
    cudaFuncSetAttribute(scene_reset_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_env);
    configured = true;
  }
  cudaMemsetAsync(pb.nwork, 0, sizeof(int) * 2 * (cfg.nsub + 1), stream);
  const int grid_env = (S.N + WARPS_SOLVE - 1) / WARPS_SOLVE;
  scene_begin_kernel<T><<<grid_env, WARPS_SOLVE * 32, smem_env, stream>>>(am, sm, cfg, S, pb, action, out);
  // narrow phase: persistent grid sized to the machine (pairs are pulled with an atomic cursor)
  const int grid_nar = pb.narrow_grid;
  for (int sub = 0; sub < cfg.nsub; sub++) {
    scene_narrow_kernel<T><<<grid_nar, WARPS_NARROW * 32, smem_nar, stream>>>(sm, cfg, S, pb, sub);
    scene_solve_kernel<T><<<grid_env, WARPS_SOLVE * 32, smem_env, stream>>>(am, sm, cfg, S, pb, out, sub);
  }
  return 1 + 2 * cfg.nsub;
}
template <typename T>
void launch_scene_reset(const StepCfg &cfg, const EnvState<T> &S, const uint8_t *mask, const so101_step_out &out, cudaStream_t stream) {
  const size_t smem = sizeof(Scratch<T>) * WARPS_SOLVE;
  cudaFuncSetAttribute(scene_reset_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  scene_reset_kernel<T><<<(S.N + WARPS_SOLVE - 1) / WARPS_SOLVE, WARPS_SOLVE * 32, smem, stream>>>(cfg, S, mask, out);
}

}  // namespace so101
