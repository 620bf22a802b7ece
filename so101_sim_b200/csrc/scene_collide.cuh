// Warp-cooperative collision detection for one environment (one warp = one env).
//
// Replaces [upstream] mj_collision for the SO100 scene: body-pair broad phase (static filters resolved at model-compile
// time), bounding-volume mid phase, convex narrow phase = boolean GJK -> EPA -> support-feature clipping (the
// multiccd manifold, so100_task.py:151).  The 32 lanes split the vertex loops (support mapping, slab candidates) and the
// EPA face loops; scalar polygon work (2-D hull, Sutherland-Hodgman clip, manifold reduction) runs on lane 0.
// The algorithm and its tie-breaking mirror oracle/so101_collide.c so that float64 runs agree to round-off.
#pragma once
#include "scene_model.cuh"

namespace so101 {

constexpr int NCON = 64;         // contacts per env the parity probe (debug_contacts) can report
constexpr int PAIRCAP = 64;      // candidate geom pairs per env per substep (after the OBB mid phase)
constexpr int CONBUF = 128;      // raw contacts per env the narrow phase may write between two kernels
constexpr int EPA_MAXV = 96, EPA_MAXF = 256;
constexpr int MAXCAND = 64, MAXFEAT = 32, MAXMANI = 4;
constexpr unsigned FULL = 0xffffffffu;

template <typename T>
struct Shape {
  int type, geom, vadr, vnum;
  T pos[3], mat[9], size[3], center[3], rbound;
};

template <typename T>
struct FPt {
  T x, y, h;
};

// Height interpolant of a support feature over its (t1, t2) plane: constant (vertex), line (edge) or plane (face).
template <typename T>
struct HPlane {
  int mode;
  T x0, y0, h0, ex, ey, eh, el, fx, fy, fh, det;
};

// Per-warp scratch of one narrow-phase pair.  The EPA polytope is dead once epa() has returned (normal, depth and
// witness points are in registers), so the manifold stage re-uses its storage.
template <typename T>
struct CollideScratch {
  union {
    struct {  // EPA polytope
      T Vw[3][EPA_MAXV], Va[3][EPA_MAXV], Vb[3][EPA_MAXV];
      unsigned char Fv[3][EPA_MAXF];  // vertex ids < EPA_MAXV
      T Fn[3][EPA_MAXF], Fd[EPA_MAXF];
      unsigned char Falive[EPA_MAXF];
      unsigned char horizon[EPA_MAXF][2];
    };
    struct {  // manifold
      T cand[3][MAXCAND];
      FPt<T> P[MAXCAND], Hh[2 * MAXCAND + 2], FA[MAXFEAT], FB[MAXFEAT], R[2 * MAXFEAT + 8], bufA[2 * MAXFEAT + 8], bufB[2 * MAXFEAT + 8];
      T mdist[2 * MAXFEAT + 8], mdist2[2 * MAXFEAT + 8];
      HPlane<T> hp[2];  // height interpolants of the two features
    };
  };
};

template <typename T> __device__ __forceinline__ T wshfl(T v, int src) { return __shfl_sync(FULL, v, src); }
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

template <typename T> __device__ __forceinline__ T dot3(const T *a, const T *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <typename T> __device__ __forceinline__ void cross3(T *r, const T *a, const T *b) {
  const T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> __device__ __forceinline__ void sub3(T *r, const T *a, const T *b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
template <typename T> __device__ __forceinline__ void mulmv(T *r, const T *m, const T *v) {
  const T x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2], z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> __device__ __forceinline__ void mulmtv(T *r, const T *m, const T *v) {
  const T x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2], z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> __device__ __forceinline__ void local2world(const Shape<T> &s, const T *l, T *w) {
  mulmv(w, s.mat, l);
  w[0] += s.pos[0]; w[1] += s.pos[1]; w[2] += s.pos[2];
}

// world pose of geom g.  xpos/xmat: the env's dynamic body poses (shared memory).
template <typename T>
__device__ __noinline__ void make_shape(const SceneModel<T> &sm, const T (*xpos)[3], const T (*xmat)[9], int g, Shape<T> &s) {
  s.type = sm.geom_type[g]; s.geom = g; s.vadr = sm.geom_vertadr[g]; s.vnum = sm.geom_vertnum[g];
  s.rbound = sm.geom_rbound[g];
  const int slot = sm.geom_slot[g];
#pragma unroll
  for (int c = 0; c < 3; c++) s.size[c] = sm.geom_size[3 * g + c];
  if (slot < 0) {
#pragma unroll
    for (int c = 0; c < 3; c++) { s.pos[c] = sm.geom_pos[3 * g + c]; s.center[c] = sm.geom_bcenter[3 * g + c]; }
#pragma unroll
    for (int c = 0; c < 9; c++) s.mat[c] = sm.geom_mat[9 * g + c];
  } else {
    const T *X = xpos[slot], *R = xmat[slot];
    T gp[3] = {sm.geom_pos[3 * g], sm.geom_pos[3 * g + 1], sm.geom_pos[3 * g + 2]}, t[3];
    mulmv(t, R, gp);
#pragma unroll
    for (int c = 0; c < 3; c++) s.pos[c] = X[c] + t[c];
    T gc[3] = {sm.geom_bcenter[3 * g], sm.geom_bcenter[3 * g + 1], sm.geom_bcenter[3 * g + 2]};
    mulmv(t, R, gc);
#pragma unroll
    for (int c = 0; c < 3; c++) s.center[c] = X[c] + t[c];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        T v = T(0);
#pragma unroll
        for (int k = 0; k < 3; k++) v += R[3 * i + k] * sm.geom_mat[9 * g + 3 * k + j];
        s.mat[3 * i + j] = v;
      }
  }
}

// support point in world direction dir (warp-cooperative for hulls; result uniform across lanes)
template <typename T>
__device__ __forceinline__ void support(const SceneModel<T> &sm, const Shape<T> &s, const T *dir, T *out, int lane) {
  T dl[3], p[3];
  mulmtv(dl, s.mat, dir);
  if (s.type == G_HULL) {
    T bv = -INFINITY;
    int bi = 0x7fffffff;
    #pragma unroll 4
    for (int i = lane; i < s.vnum; i += 32) {
      const Vec4<T> v = sm.hull_vert[s.vadr + i];
      const T val = v.x * dl[0] + v.y * dl[1] + v.z * dl[2];
      if (val > bv) { bv = val; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T ov = __shfl_xor_sync(FULL, bv, o);
      const int oi = __shfl_xor_sync(FULL, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (bi == 0x7fffffff) bi = 0;  // non-finite direction (diverged env): any vertex, but never out of bounds
    const Vec4<T> v = sm.hull_vert[s.vadr + bi];
    p[0] = v.x; p[1] = v.y; p[2] = v.z;
  } else if (s.type == G_BOX) {
#pragma unroll
    for (int c = 0; c < 3; c++) p[c] = dl[c] >= T(0) ? s.size[c] : -s.size[c];
  } else if (s.type == G_CYLINDER) {
    const T n = t_sqrt(dl[0] * dl[0] + dl[1] * dl[1]);
    p[0] = n > T(1e-15) ? s.size[0] * dl[0] / n : T(0); p[1] = n > T(1e-15) ? s.size[0] * dl[1] / n : T(0);
    p[2] = dl[2] >= T(0) ? s.size[1] : -s.size[1];
  } else if (s.type == G_CAPSULE || s.type == G_SPHERE) {
    const T n = t_sqrt(dot3(dl, dl));
#pragma unroll
    for (int c = 0; c < 3; c++) p[c] = n > T(1e-15) ? s.size[0] * dl[c] / n : T(0);
    if (s.type == G_CAPSULE) p[2] += dl[2] >= T(0) ? s.size[1] : -s.size[1];
  } else { p[0] = p[1] = p[2] = T(0); }
  local2world(s, p, out);
}

template <typename T>
struct MPoint {
  T w[3], a[3], b[3];
};
template <typename T>
__device__ __noinline__ void msupport(const SceneModel<T> &sm, const Shape<T> &A, const Shape<T> &B, const T *d, MPoint<T> &p, int lane) {
  const T nd[3] = {-d[0], -d[1], -d[2]};
  support(sm, A, d, p.a, lane); support(sm, B, nd, p.b, lane);
  sub3(p.w, p.a, p.b);
}

// ---- one THREAD per pair (boolean-GJK kernel): the thread scans all hull vertices itself.  Lanes of a warp work on the same
// hull for different envs (queue order), so every vertex load is a warp-wide broadcast.  Same first-maximum tie-break as the
// warp-cooperative support() above.
template <typename T>
__device__ __forceinline__ void support_seq(const SceneModel<T> &sm, const Shape<T> &s, const T *dir, T *out) {
  T dl[3], p[3];
  mulmtv(dl, s.mat, dir);
  if (s.type == G_HULL) {
    T bv = -INFINITY;
    int bi = 0x7fffffff;
    const Vec4<T> *vt = sm.hull_vert + s.vadr;
#pragma unroll 4
    for (int i = 0; i < s.vnum; i++) {
      const Vec4<T> v = vt[i];
      const T val = v.x * dl[0] + v.y * dl[1] + v.z * dl[2];
      if (val > bv) { bv = val; bi = i; }
    }
    if (bi == 0x7fffffff) bi = 0;
    const Vec4<T> v = vt[bi];
    p[0] = v.x; p[1] = v.y; p[2] = v.z;
  } else if (s.type == G_BOX) {
#pragma unroll
    for (int c = 0; c < 3; c++) p[c] = dl[c] >= T(0) ? s.size[c] : -s.size[c];
  } else if (s.type == G_CYLINDER) {
    const T n = t_sqrt(dl[0] * dl[0] + dl[1] * dl[1]);
    p[0] = n > T(1e-15) ? s.size[0] * dl[0] / n : T(0); p[1] = n > T(1e-15) ? s.size[0] * dl[1] / n : T(0);
    p[2] = dl[2] >= T(0) ? s.size[1] : -s.size[1];
  } else if (s.type == G_CAPSULE || s.type == G_SPHERE) {
    const T n = t_sqrt(dot3(dl, dl));
#pragma unroll
    for (int c = 0; c < 3; c++) p[c] = n > T(1e-15) ? s.size[0] * dl[c] / n : T(0);
    if (s.type == G_CAPSULE) p[2] += dl[2] >= T(0) ? s.size[1] : -s.size[1];
  } else { p[0] = p[1] = p[2] = T(0); }
  local2world(s, p, out);
}
template <typename T>
__device__ __noinline__ void msupport_seq(const SceneModel<T> &sm, const Shape<T> &A, const Shape<T> &B, const T *d, MPoint<T> &p) {
  const T nd[3] = {-d[0], -d[1], -d[2]};
  support_seq(sm, A, d, p.a); support_seq(sm, B, nd, p.b);
  sub3(p.w, p.a, p.b);
}

// boolean GJK; mirrors gjk_intersect() of the oracle.  Uniform control flow across the warp.
// SEQ = true: one thread per pair (msupport_seq); false: one warp per pair (msupport)
template <typename T, bool SEQ>
__device__ __noinline__ int gjk_intersect(const SceneModel<T> &sm, const Shape<T> &A, const Shape<T> &B, MPoint<T> *S, int &np, int &iters, int lane) {
  T d[3];
  sub3(d, B.center, A.center);
  if (dot3(d, d) < T(1e-20)) { d[0] = T(1); d[1] = T(0); d[2] = T(0); }
  int n = 0;
  #pragma unroll 1
  for (int it = 0; it < 64; it++) {
    MPoint<T> p;
    iters = it + 1;
    if constexpr (SEQ) msupport_seq(sm, A, B, d, p); else msupport(sm, A, B, d, p, lane);
    if (dot3(p.w, d) < T(0)) { np = n; return 0; }
    S[n++] = p;
    if (n == 1) { d[0] = -S[0].w[0]; d[1] = -S[0].w[1]; d[2] = -S[0].w[2]; }
    else if (n == 2) {
      T ab[3], ao[3] = {-S[1].w[0], -S[1].w[1], -S[1].w[2]}, t[3];
      sub3(ab, S[0].w, S[1].w);
      if (dot3(ab, ao) > T(0)) { cross3(t, ab, ao); cross3(d, t, ab); }
      else { S[0] = S[1]; n = 1; d[0] = ao[0]; d[1] = ao[1]; d[2] = ao[2]; }
    } else if (n == 3) {
      T *a = S[2].w, *b = S[1].w, *c = S[0].w, ab[3], ac[3], ao[3] = {-a[0], -a[1], -a[2]}, abc[3], t[3], u[3];
      sub3(ab, b, a); sub3(ac, c, a); cross3(abc, ab, ac);
      cross3(t, abc, ac);
      bool star = false;
      if (dot3(t, ao) > T(0)) {
        if (dot3(ac, ao) > T(0)) { S[1] = S[2]; n = 2; cross3(u, ac, ao); cross3(d, u, ac); }
        else star = true;
      } else {
        cross3(t, ab, abc);
        if (dot3(t, ao) > T(0)) star = true;
        else if (dot3(abc, ao) > T(0)) { d[0] = abc[0]; d[1] = abc[1]; d[2] = abc[2]; }
        else { MPoint<T> tmp = S[0]; S[0] = S[1]; S[1] = tmp; d[0] = -abc[0]; d[1] = -abc[1]; d[2] = -abc[2]; }
      }
      if (star) {
        if (dot3(ab, ao) > T(0)) { S[0] = S[1]; S[1] = S[2]; n = 2; cross3(u, ab, ao); cross3(d, u, ab); }
        else { S[0] = S[2]; n = 1; d[0] = ao[0]; d[1] = ao[1]; d[2] = ao[2]; }
      }
    } else {
      T *a = S[3].w, *b = S[2].w, *c = S[1].w, *e = S[0].w, ab[3], ac[3], ad[3], ao[3] = {-a[0], -a[1], -a[2]}, abc[3], acd[3], adb[3];
      sub3(ab, b, a); sub3(ac, c, a); sub3(ad, e, a);
      cross3(abc, ab, ac); cross3(acd, ac, ad); cross3(adb, ad, ab);
      if (dot3(abc, ad) > T(0)) { abc[0] = -abc[0]; abc[1] = -abc[1]; abc[2] = -abc[2]; }
      if (dot3(acd, ab) > T(0)) { acd[0] = -acd[0]; acd[1] = -acd[1]; acd[2] = -acd[2]; }
      if (dot3(adb, ac) > T(0)) { adb[0] = -adb[0]; adb[1] = -adb[1]; adb[2] = -adb[2]; }
      const T da = dot3(abc, ao), db = dot3(acd, ao), dc = dot3(adb, ao);
      if (da > T(0) && da >= db && da >= dc) { S[0] = S[1]; S[1] = S[2]; S[2] = S[3]; n = 3; d[0] = abc[0]; d[1] = abc[1]; d[2] = abc[2]; }
      else if (db > T(0) && db >= dc) { S[2] = S[3]; n = 3; d[0] = acd[0]; d[1] = acd[1]; d[2] = acd[2]; }
      else if (dc > T(0)) { S[1] = S[2]; S[2] = S[3]; n = 3; d[0] = adb[0]; d[1] = adb[1]; d[2] = adb[2]; }
      else { np = 4; return 1; }
    }
    if (dot3(d, d) < T(1e-30)) { np = n; return 1; }
  }
  np = n;
  return n == 4;
}

// --------------------------------------------------------------------------------------------- EPA (shared memory)
template <typename T>
__device__ __forceinline__ void epa_getv(const CollideScratch<T> &cs, int i, T *w) { w[0] = cs.Vw[0][i]; w[1] = cs.Vw[1][i]; w[2] = cs.Vw[2][i]; }
template <typename T>
__device__ __forceinline__ void epa_putv(CollideScratch<T> &cs, int i, const MPoint<T> &p, int lane) {
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < 3; c++) { cs.Vw[c][i] = p.w[c]; cs.Va[c][i] = p.a[c]; cs.Vb[c][i] = p.b[c]; }
  }
  __syncwarp();
}
// uniform; returns new face index or -1
template <typename T>
__device__ __noinline__ int epa_add_face(CollideScratch<T> &cs, int &nf, int a, int b, int c, const T *inside, int lane) {
  if (nf >= EPA_MAXF) return -1;
  T va[3], vb[3], vc[3], ab[3], ac[3], n[3], t[3];
  epa_getv(cs, a, va); epa_getv(cs, b, vb); epa_getv(cs, c, vc);
  sub3(ab, vb, va); sub3(ac, vc, va);
  cross3(n, ab, ac);
  const T l = t_sqrt(dot3(n, n));
  if (l < T(1e-30)) return -1;
  n[0] /= l; n[1] /= l; n[2] /= l;
  sub3(t, va, inside);
  int v1 = b, v2 = c;
  if (dot3(n, t) < T(0)) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; v1 = c; v2 = b; }
  if (lane == 0) {
    cs.Fv[0][nf] = a; cs.Fv[1][nf] = v1; cs.Fv[2][nf] = v2;
    cs.Fn[0][nf] = n[0]; cs.Fn[1][nf] = n[1]; cs.Fn[2][nf] = n[2];
    cs.Fd[nf] = dot3(n, va);
    cs.Falive[nf] = 1;
  }
  __syncwarp();
  return nf++;
}

template <typename T>
__device__ __noinline__ int epa(const SceneModel<T> &sm, CollideScratch<T> &cs, const Shape<T> &A, const Shape<T> &B, const MPoint<T> *S, int n, T *normal,
                   T &depth, T *pa, T *pb, int &iters, int lane) {
  int nv = 0, nf = 0;
  iters = 0;
  if (n == 1) return 0;
  #pragma unroll 1
  for (int i = 0; i < n; i++) epa_putv(cs, nv++, S[i], lane);
  if (nv == 2) {
    T v0[3], v1[3], ab[3], ax[3] = {T(0), T(0), T(0)}, d[3];
    epa_getv(cs, 0, v0); epa_getv(cs, 1, v1);
    sub3(ab, v1, v0);
    const int k = t_abs(ab[0]) < t_abs(ab[1]) ? (t_abs(ab[0]) < t_abs(ab[2]) ? 0 : 2) : (t_abs(ab[1]) < t_abs(ab[2]) ? 1 : 2);
    ax[k] = T(1); cross3(d, ab, ax);
    MPoint<T> p;
    msupport(sm, A, B, d, p, lane);
    T t[3], cr[3];
    sub3(t, p.w, v0); cross3(cr, ab, t);
    if (dot3(cr, cr) < T(1e-24)) { d[0] = -d[0]; d[1] = -d[1]; d[2] = -d[2]; msupport(sm, A, B, d, p, lane); }
    epa_putv(cs, nv++, p, lane);
  }
  if (nv == 3) {
    T v0[3], v1[3], v2[3], ab[3], ac[3], nn[3], t[3];
    epa_getv(cs, 0, v0); epa_getv(cs, 1, v1); epa_getv(cs, 2, v2);
    sub3(ab, v1, v0); sub3(ac, v2, v0); cross3(nn, ab, ac);
    if (dot3(nn, nn) < T(1e-30)) return 0;
    MPoint<T> p;
    msupport(sm, A, B, nn, p, lane);
    sub3(t, p.w, v0);
    if (t_abs(dot3(t, nn)) < T(1e-12) * t_sqrt(dot3(nn, nn))) {
      const T m[3] = {-nn[0], -nn[1], -nn[2]};
      msupport(sm, A, B, m, p, lane);
      sub3(t, p.w, v0);
      if (t_abs(dot3(t, nn)) < T(1e-12) * t_sqrt(dot3(nn, nn))) return 0;
    }
    epa_putv(cs, nv++, p, lane);
  }
  T inside[3] = {T(0), T(0), T(0)};
  for (int i = 0; i < 4; i++) { T v[3]; epa_getv(cs, i, v); for (int k = 0; k < 3; k++) inside[k] += T(0.25) * v[k]; }
  if (epa_add_face(cs, nf, 0, 1, 2, inside, lane) < 0 || epa_add_face(cs, nf, 0, 1, 3, inside, lane) < 0 ||
      epa_add_face(cs, nf, 0, 2, 3, inside, lane) < 0 || epa_add_face(cs, nf, 1, 2, 3, inside, lane) < 0) return 0;
  int best = -1;
  auto find_best = [&]() {
    T bd = INFINITY;
    int bi = 0x7fffffff;
    #pragma unroll 1
    for (int f = lane; f < nf; f += 32)
      if (cs.Falive[f] && cs.Fd[f] < bd) { bd = cs.Fd[f]; bi = f; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T ov = __shfl_xor_sync(FULL, bd, o);
      const int oi = __shfl_xor_sync(FULL, bi, o);
      if (ov < bd || (ov == bd && oi < bi)) { bd = ov; bi = oi; }
    }
    return bi == 0x7fffffff ? -1 : bi;
  };
  #pragma unroll 1
  for (int it = 0; it < 80; it++) {
    best = find_best();
    if (best < 0) return 0;
    if (nv >= EPA_MAXV) break;
    iters = it + 1;
    const T bn[3] = {cs.Fn[0][best], cs.Fn[1][best], cs.Fn[2][best]}, bd = cs.Fd[best];
    MPoint<T> p;
    msupport(sm, A, B, bn, p, lane);
    const T adv = dot3(p.w, bn) - bd;
    if (adv < T(sizeof(T) == 8 ? 1e-9 : 1e-6)) break;
    epa_putv(cs, nv, p, lane);
    // visibility (lanes over faces), then the horizon in face order on lane 0 (same order as the oracle)
    #pragma unroll 1
    for (int f = lane; f < nf; f += 32) {
      if (!cs.Falive[f]) continue;
      T v0[3], t[3], fn[3] = {cs.Fn[0][f], cs.Fn[1][f], cs.Fn[2][f]};
      epa_getv(cs, cs.Fv[0][f], v0);
      sub3(t, p.w, v0);
      if (dot3(fn, t) > T(sizeof(T) == 8 ? 1e-12 : 1e-9)) cs.Falive[f] = 2;  // 2 = visible, to be removed
    }
    __syncwarp();
    int nh = 0;
    if (lane == 0) {
      #pragma unroll 1
      for (int f = 0; f < nf; f++) {
        if (cs.Falive[f] != 2) continue;
        cs.Falive[f] = 0;
        for (int e = 0; e < 3; e++) {
          const int a = cs.Fv[e][f], b = cs.Fv[(e + 1) % 3][f];
          int found = 0;
          #pragma unroll 1
          for (int h = 0; h < nh; h++)
            if (cs.horizon[h][0] == b && cs.horizon[h][1] == a) {
              cs.horizon[h][0] = cs.horizon[nh - 1][0]; cs.horizon[h][1] = cs.horizon[nh - 1][1]; nh--; found = 1;
              break;
            }
          if (!found && nh < EPA_MAXF) { cs.horizon[nh][0] = a; cs.horizon[nh][1] = b; nh++; }
        }
      }
    }
    nh = wshfl(nh, 0);
    __syncwarp();
    if (nh == 0) break;
    int failed = 0;
    #pragma unroll 1
    for (int h = 0; h < nh; h++)
      if (epa_add_face(cs, nf, cs.horizon[h][0], cs.horizon[h][1], nv, inside, lane) < 0) failed = 1;
    nv++;
    if (failed) break;
  }
  if (best < 0 || !cs.Falive[best]) {
    best = find_best();
    if (best < 0) return 0;
  }
  const T fn[3] = {cs.Fn[0][best], cs.Fn[1][best], cs.Fn[2][best]}, fd = cs.Fd[best];
  normal[0] = fn[0]; normal[1] = fn[1]; normal[2] = fn[2];
  depth = fd > T(0) ? fd : T(0);
  const T p[3] = {fn[0] * fd, fn[1] * fd, fn[2] * fd};
  const int i0 = cs.Fv[0][best], i1 = cs.Fv[1][best], i2 = cs.Fv[2][best];
  T a[3], b[3], c[3], v0[3], v1[3], v2[3];
  epa_getv(cs, i0, a); epa_getv(cs, i1, b); epa_getv(cs, i2, c);
  sub3(v0, b, a); sub3(v1, c, a); sub3(v2, p, a);
  const T d00 = dot3(v0, v0), d01 = dot3(v0, v1), d11 = dot3(v1, v1), d20 = dot3(v2, v0), d21 = dot3(v2, v1), den = d00 * d11 - d01 * d01;
  T bv = T(1.0 / 3), bw = T(1.0 / 3);
  if (t_abs(den) > T(1e-30)) { bv = (d11 * d20 - d01 * d21) / den; bw = (d00 * d21 - d01 * d20) / den; }
  const T bu = T(1) - bv - bw;
  for (int k = 0; k < 3; k++) {
    pa[k] = bu * cs.Va[k][i0] + bv * cs.Va[k][i1] + bw * cs.Va[k][i2];
    pb[k] = bu * cs.Vb[k][i0] + bv * cs.Vb[k][i1] + bw * cs.Vb[k][i2];
  }
  __syncwarp();  // the manifold stage re-uses the polytope's storage: every lane must be done reading it
  return 1;
}

// --------------------------------------------------------------------------------------------- support features
template <typename T>
__device__ __forceinline__ void frame_from_normal(const T *n, T *frame) {  // [upstream] mju_makeFrame
  T *x = frame, *y = frame + 3, *z = frame + 6;
  T l = t_sqrt(dot3(n, n));
  for (int c = 0; c < 3; c++) x[c] = n[c] / l;
  y[0] = y[1] = y[2] = T(0);
  if (x[1] < T(0.5) && x[1] > T(-0.5)) y[1] = T(1); else y[2] = T(1);
  const T dd = dot3(x, y);
  for (int c = 0; c < 3; c++) y[c] -= dd * x[c];
  l = t_sqrt(dot3(y, y));
  for (int c = 0; c < 3; c++) y[c] /= l;
  cross3(z, x, y);
}

// vertices of s within delta of the support plane along dir -> CCW 2-D convex polygon in (t1,t2) with heights.  Result in out (smem).
template <typename T>
__device__ __noinline__ int feature(const SceneModel<T> &sm, CollideScratch<T> &cs, const Shape<T> &s, const T *dir, const T *t1, const T *t2, T delta,
                       FPt<T> *out, int lane) {
  T sp[3];
  support(sm, s, dir, sp, lane);
  const T hmax = dot3(sp, dir);
  int nc = 0;
  if (s.type == G_HULL) {
    T dl[3];
    mulmtv(dl, s.mat, dir);
    const T off = dot3(s.pos, dir);
    #pragma unroll 1
    for (int base = 0; base < s.vnum && nc < MAXCAND; base += 32) {
      const int i = base + lane;
      bool in = false;
      Vec4<T> v{};
      if (i < s.vnum) { v = sm.hull_vert[s.vadr + i]; in = (v.x * dl[0] + v.y * dl[1] + v.z * dl[2]) + off >= hmax - delta; }
      const unsigned m = __ballot_sync(FULL, in);
      const int idx = nc + __popc(m & ((1u << lane) - 1));
      if (in && idx < MAXCAND) {
        const T l[3] = {v.x, v.y, v.z};
        T w[3];
        local2world(s, l, w);
        cs.cand[0][idx] = w[0]; cs.cand[1][idx] = w[1]; cs.cand[2][idx] = w[2];
      }
      nc += __popc(m);
    }
    if (nc > MAXCAND) nc = MAXCAND;
  } else if (lane == 0) {
    T w[3];
    if (s.type == G_BOX) {
      #pragma unroll 1
      for (int i = 0; i < 8; i++) {
        const T l[3] = {(i & 1 ? T(1) : T(-1)) * s.size[0], (i & 2 ? T(1) : T(-1)) * s.size[1], (i & 4 ? T(1) : T(-1)) * s.size[2]};
        local2world(s, l, w);
        if (dot3(w, dir) >= hmax - delta) { cs.cand[0][nc] = w[0]; cs.cand[1][nc] = w[1]; cs.cand[2][nc] = w[2]; nc++; }
      }
    } else if (s.type == G_CYLINDER) {
      #pragma unroll 1
      for (int cap = -1; cap <= 1; cap += 2)
        #pragma unroll 1
        for (int i = 0; i < 16; i++) {
          T sn, cn;
          t_sincos(T(2 * 3.14159265358979323846 / 16) * T(i), &sn, &cn);
          const T l[3] = {s.size[0] * cn, s.size[0] * sn, T(cap) * s.size[1]};
          local2world(s, l, w);
          if (dot3(w, dir) >= hmax - delta) { cs.cand[0][nc] = w[0]; cs.cand[1][nc] = w[1]; cs.cand[2][nc] = w[2]; nc++; }
        }
    } else if (s.type == G_CAPSULE) {
      #pragma unroll 1
      for (int e = -1; e <= 1; e += 2) {
        const T l[3] = {T(0), T(0), T(e) * s.size[1]};
        local2world(s, l, w);
        for (int c = 0; c < 3; c++) w[c] += s.size[0] * dir[c];
        if (dot3(w, dir) >= hmax - delta) { cs.cand[0][nc] = w[0]; cs.cand[1][nc] = w[1]; cs.cand[2][nc] = w[2]; nc++; }
      }
    }
  }
  nc = wshfl(nc, 0);
  if (nc == 0) {
    if (lane == 0) { cs.cand[0][0] = sp[0]; cs.cand[1][0] = sp[1]; cs.cand[2][0] = sp[2]; }
    nc = 1;
  }
  __syncwarp();
  // project; stable rank sort by (x, y); monotone-chain hull on lane 0
  FPt<T> mine[2];
  int rank[2] = {0, 0};
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int i = lane + 32 * k;
    if (i < nc) {
      const T w[3] = {cs.cand[0][i], cs.cand[1][i], cs.cand[2][i]};
      mine[k].x = dot3(w, t1); mine[k].y = dot3(w, t2); mine[k].h = dot3(w, dir);
    }
  }
  // ranks need everyone's projected coordinates: stage them in P (unsorted) then scatter into Hh[0..nc) as the sorted list
#pragma unroll
  for (int k = 0; k < 2; k++) { const int i = lane + 32 * k; if (i < nc) cs.P[i] = mine[k]; }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int i = lane + 32 * k;
    if (i < nc) {
      int r = 0;
      #pragma unroll 1
      for (int j = 0; j < nc; j++) {
        const FPt<T> o = cs.P[j];
        if (o.x < mine[k].x || (o.x == mine[k].x && (o.y < mine[k].y || (o.y == mine[k].y && j < i)))) r++;
      }
      rank[k] = r;
    }
  }
  __syncwarp();
  FPt<T> *sorted = cs.bufA;  // nc <= MAXCAND = 2*MAXFEAT <= size of bufA
#pragma unroll
  for (int k = 0; k < 2; k++) { const int i = lane + 32 * k; if (i < nc) sorted[rank[k]] = mine[k]; }
  __syncwarp();
  int nout = 0;
  if (lane == 0) {
    if (nc <= 2) { for (int i = 0; i < nc; i++) out[i] = sorted[i]; nout = nc; }
    else {
      FPt<T> *H = cs.Hh;
      int k = 0;
      #pragma unroll 1
      for (int i = 0; i < nc; i++) {
        while (k >= 2 && (H[k - 1].x - H[k - 2].x) * (sorted[i].y - H[k - 2].y) - (H[k - 1].y - H[k - 2].y) * (sorted[i].x - H[k - 2].x) <= T(1e-14)) k--;
        H[k++] = sorted[i];
      }
      #pragma unroll 1
      for (int i = nc - 2, t = k + 1; i >= 0; i--) {
        while (k >= t && (H[k - 1].x - H[k - 2].x) * (sorted[i].y - H[k - 2].y) - (H[k - 1].y - H[k - 2].y) * (sorted[i].x - H[k - 2].x) <= T(1e-14)) k--;
        H[k++] = sorted[i];
      }
      k--;
      if (k > MAXFEAT) k = MAXFEAT;
      #pragma unroll 1
      for (int i = 0; i < k; i++) out[i] = H[i];
      nout = k;
    }
  }
  nout = wshfl(nout, 0);
  __syncwarp();
  return nout;
}

// ---- manifold helpers mirroring the oracle (same expressions, evaluated once per feature / in parallel over points)
// Height of feature P at (x, y): the oracle's feature_height() split into a per-feature setup (lane 0) ...
template <typename T>
__device__ __noinline__ void feature_plane(const FPt<T> *P, int n, HPlane<T> &hp) {
  hp.mode = 0; hp.x0 = P[0].x; hp.y0 = P[0].y; hp.h0 = P[0].h;
  if (n == 1) return;
  int i1 = 1;
  T best = T(-1);
#pragma unroll 1
  for (int i = 1; i < n; i++) { const T dx = P[i].x - P[0].x, dy = P[i].y - P[0].y, l = dx * dx + dy * dy; if (l > best) { best = l; i1 = i; } }
  const T ex = P[i1].x - P[0].x, ey = P[i1].y - P[0].y, eh = P[i1].h - P[0].h, el = ex * ex + ey * ey;
  hp.ex = ex; hp.ey = ey; hp.eh = eh; hp.el = el;
  if (el < T(1e-20)) return;
  hp.mode = 1;
  if (n == 2) return;
  int i2 = -1;
  best = T(0);
#pragma unroll 1
  for (int i = 1; i < n; i++) { const T a = t_abs(ex * (P[i].y - P[0].y) - ey * (P[i].x - P[0].x)); if (a > best) { best = a; i2 = i; } }
  if (i2 < 0 || best < T(1e-12) * el) return;
  hp.mode = 2;
  hp.fx = P[i2].x - P[0].x; hp.fy = P[i2].y - P[0].y; hp.fh = P[i2].h - P[0].h; hp.det = ex * hp.fy - ey * hp.fx;
}
// ... and a per-point evaluation (any lane)
template <typename T>
__device__ __forceinline__ T plane_height(const HPlane<T> &hp, T x, T y) {
  if (hp.mode == 0) return hp.h0;
  const T px = x - hp.x0, py = y - hp.y0;
  if (hp.mode == 1) { const T t = (px * hp.ex + py * hp.ey) / hp.el; return hp.h0 + t * hp.eh; }
  const T u = (px * hp.fy - py * hp.fx) / hp.det, v = (hp.ex * py - hp.ey * px) / hp.det;
  return hp.h0 + u * hp.eh + v * hp.fh;
}

// Sutherland-Hodgman clip of `subj` (n points; polygon, segment or point) against the convex polygon `clip` (m >= 3), all
// lanes: one lane per subject vertex, ballot compaction keeps the oracle's output order.  Result in `out`, count returned.
template <typename T>
__device__ __noinline__ int clip_poly(CollideScratch<T> &cs, const FPt<T> *subj, int n, const FPt<T> *clip, int m, FPt<T> *out, int lane) {
  constexpr int CAP = 2 * MAXFEAT + 8;
  int na = n;
  FPt<T> *in = cs.bufA, *res = cs.bufB;
#pragma unroll 1
  for (int i = lane; i < n; i += 32) in[i] = subj[i];
  __syncwarp();
#pragma unroll 1
  for (int e = 0; e < m && na > 0; e++) {
    const T ax = clip[e].x, ay = clip[e].y, bx = clip[(e + 1) % m].x, by = clip[(e + 1) % m].y;
    const T ex = bx - ax, ey = by - ay, tol = T(1e-12);
    int nr = 0;
    if (na == 2) {  // open segment (uniform across lanes; written by lane 0)
      const FPt<T> P = in[0], Q = in[1];
      const T sp = ex * (P.y - ay) - ey * (P.x - ax), sq = ex * (Q.y - ay) - ey * (Q.x - ax);
      const bool pin = sp >= -tol, qin = sq >= -tol;
      FPt<T> r0 = P, r1 = Q;
      if (pin && qin) nr = 2;
      else if (pin || qin) {
        const T t = sp / (sp - sq);
        const FPt<T> I = {P.x + t * (Q.x - P.x), P.y + t * (Q.y - P.y), P.h + t * (Q.h - P.h)};
        if (pin) r1 = I; else r0 = I;
        nr = 2;
      }
      if (lane == 0 && nr) { res[0] = r0; res[1] = r1; }
    } else {
#pragma unroll 1
      for (int i0 = 0; i0 < na; i0 += 32) {
        const int i = i0 + lane;
        bool pin = false, cross = false;
        FPt<T> P{}, I{};
        if (i < na && !(na == 1 && i > 0)) {
          P = in[i];
          const FPt<T> Q = in[(i + 1) % na];
          const T sp = ex * (P.y - ay) - ey * (P.x - ax), sq = ex * (Q.y - ay) - ey * (Q.x - ax);
          const bool qin = sq >= -tol;
          pin = sp >= -tol;
          cross = na > 1 && pin != qin;
          if (cross) { const T t = sp / (sp - sq); I = FPt<T>{P.x + t * (Q.x - P.x), P.y + t * (Q.y - P.y), P.h + t * (Q.h - P.h)}; }
        }
        const unsigned mp = __ballot_sync(FULL, pin), mx = __ballot_sync(FULL, cross), lt = (1u << lane) - 1;
        const int o = nr + __popc(mp & lt) + __popc(mx & lt);
        if (pin && o < CAP) res[o] = P;
        if (cross && o + (pin ? 1 : 0) < CAP) res[o + (pin ? 1 : 0)] = I;
        nr += __popc(mp) + __popc(mx);
      }
      if (nr > CAP) nr = CAP;
    }
    __syncwarp();
    FPt<T> *tmp = in; in = res; res = tmp;
    na = nr;
    if (na > 2 * MAXFEAT) na = 2 * MAXFEAT;
  }
#pragma unroll 1
  for (int i = lane; i < na; i += 32) out[i] = in[i];
  __syncwarp();
  return na;
}

template <typename T>
__device__ __noinline__ int reduce_manifold(FPt<T> *P, T *dist, int n) {
  if (n <= MAXMANI) return n;
  int sel[4] = {0, -1, -1, -1};
  // tolerant comparisons: the first candidate in polygon order wins a tie in any arithmetic (see the oracle)
  #pragma unroll 1
  for (int i = 1; i < n; i++) if (dist[i] < dist[sel[0]] - T(1e-7)) sel[0] = i;
  T best = T(-1);
  #pragma unroll 1
  for (int i = 0; i < n; i++) { const T dx = P[i].x - P[sel[0]].x, dy = P[i].y - P[sel[0]].y, l = dx * dx + dy * dy; if (l > best * T(1.0001) + T(1e-12)) { best = l; sel[1] = i; } }
  const T ex = P[sel[1]].x - P[sel[0]].x, ey = P[sel[1]].y - P[sel[0]].y;
  T bp = T(0), bn = T(0);
  #pragma unroll 1
  for (int i = 0; i < n; i++) {
    if (i == sel[0] || i == sel[1]) continue;  // their cross product is 0 up to round-off (FMA contraction makes it +-eps)
    const T s = ex * (P[i].y - P[sel[0]].y) - ey * (P[i].x - P[sel[0]].x);
    if (s > bp * T(1.0001) + T(1e-12)) { bp = s; sel[2] = i; }
    if (s < bn * T(1.0001) - T(1e-12)) { bn = s; sel[3] = i; }
  }
  FPt<T> Q[4];
  T qd[4];
  int k = 0;
  for (int i = 0; i < 4; i++) if (sel[i] >= 0) { Q[k] = P[sel[i]]; qd[k] = dist[sel[i]]; k++; }
  #pragma unroll 1
  for (int i = 0; i < k; i++) { P[i] = Q[i]; dist[i] = qd[i]; }
  return k;
}

}  // namespace so101
