// Newton solver for the arm-only constraint set, one env per thread, all state in registers.
//
// Rows present without contacts ([upstream] mj_makeConstraint order): 6 dof friction-loss rows (always on,
// scene_pbr.xml:10) and at most one joint-limit row per hinge.  Every Jacobian row is +-e_i, so the primal Hessian is
// M + diag(h) and the whole solve is a handful of 6x6 Cholesky factorisations.
// Follows [upstream] mj_fwdConstraint with solver=Newton (primal, warm-started, exact 1-D line search), written in
// terms of the increment delta = qacc - qacc_smooth to avoid the float32 cancellation in (M qacc - qfrc_smooth).
#pragma once
#include "arm_dynamics.cuh"

namespace so101 {

template <typename T>
struct ArmRows {
  T jar0_f[NJ];            // friction rows: jar at delta = 0
  T jar0_l[NJ], D_l[NJ];   // limit rows (D_l == 0: inactive)
  T js[NJ];                // limit row Jacobian sign
};

template <typename T>
__device__ __forceinline__ void arm_make_rows(const ArmModelT<T> &am, const T (&q)[NJ], const T (&qd)[NJ], const T (&qacc_s)[NJ], ArmRows<T> &r) {
#pragma unroll
  for (int i = 0; i < NJ; i++) {
    r.jar0_f[i] = qacc_s[i] + am.fr_B[i] * qd[i];  // aref = -B * vel, K = 0 for friction rows
    r.D_l[i] = T(0); r.jar0_l[i] = T(0); r.js[i] = T(1);
    if (am.limited[i]) {
      const T dlo = q[i] - am.range[i][0], dhi = am.range[i][1] - q[i];
      T dist = T(0); bool act = false;
      if (dlo < T(0)) { dist = dlo; r.js[i] = T(1); act = true; }
      else if (dhi < T(0)) { dist = dhi; r.js[i] = T(-1); act = true; }
      if (act) {
        const T imp = impedance(am.lim_solimp[i], dist, T(0));
        T R = (T(1) - imp) / imp * am.invweight0[i];
        R = R > T(1e-15) ? R : T(1e-15);
        r.D_l[i] = T(1) / R;
        const T aref = -am.lim_B[i] * (r.js[i] * qd[i]) - am.lim_K[i] * imp * dist;
        r.jar0_l[i] = r.js[i] * qacc_s[i] - aref;
      }
    }
  }
}

// cost, first and second derivative of the row terms along delta + alpha * s
template <typename T>
__device__ __forceinline__ void arm_rows_line(const ArmModelT<T> &am, const ArmRows<T> &r, const T (&delta)[NJ], const T (&s)[NJ], T alpha,
                                              T &c, T &g, T &h) {
#pragma unroll
  for (int i = 0; i < NJ; i++) {
    {
      const T x = r.jar0_f[i] + delta[i] + alpha * s[i], eta = am.frictionloss[i], rf = am.fr_R[i] * eta;
      if (x <= -rf) { c += eta * (T(-0.5) * rf - x); g -= eta * s[i]; }
      else if (x >= rf) { c += eta * (T(-0.5) * rf + x); g += eta * s[i]; }
      else { c += T(0.5) * am.fr_D[i] * x * x; g += am.fr_D[i] * x * s[i]; h += am.fr_D[i] * s[i] * s[i]; }
    }
    if (r.D_l[i] > T(0)) {
      const T jv = r.js[i] * s[i], x = r.jar0_l[i] + r.js[i] * delta[i] + alpha * jv;
      if (x < T(0)) { c += T(0.5) * r.D_l[i] * x * x; g += r.D_l[i] * x * jv; h += r.D_l[i] * jv * jv; }
    }
  }
}

// Returns the number of Newton iterations.  delta: in = warm start (qacc_warmstart - qacc_smooth), out = solution.
template <typename T>
__device__ __forceinline__ int arm_solve(const ArmModelT<T> &am, const T (&M)[21], const ArmRows<T> &r, T (&delta)[NJ], int max_iter, T tol) {
  const T xeps = sizeof(T) == 8 ? T(1e-14) : T(2e-6);  // relative resolution of a line-search / Newton step
  T Md[NJ];
  // warm start: keep delta only if it beats delta = 0 (qacc_smooth)
  {
    T zero[NJ] = {T(0), T(0), T(0), T(0), T(0), T(0)};
    T c0 = T(0), cw = T(0), g = T(0), h = T(0);
    arm_rows_line(am, r, zero, zero, T(0), c0, g, h);
    symmv6(M, delta, Md);
#pragma unroll
    for (int i = 0; i < NJ; i++) cw += T(0.5) * delta[i] * Md[i];
    arm_rows_line(am, r, delta, zero, T(0), cw, g, h);
    if (!(cw < c0)) {
#pragma unroll
      for (int i = 0; i < NJ; i++) delta[i] = T(0);
    }
  }
  int iter = 0;
  for (; iter < max_iter; iter++) {
    T grad[NJ], H[21];
    symmv6(M, delta, Md);
#pragma unroll
    for (int i = 0; i < 21; i++) H[i] = M[i];
    T gn = T(0);
#pragma unroll
    for (int i = 0; i < NJ; i++) {
      T f = T(0), hd = T(0);
      const T x = r.jar0_f[i] + delta[i], eta = am.frictionloss[i], rf = am.fr_R[i] * eta;
      if (x <= -rf) f = eta;
      else if (x >= rf) f = -eta;
      else { f = -am.fr_D[i] * x; hd = am.fr_D[i]; }
      if (r.D_l[i] > T(0)) {
        const T xl = r.jar0_l[i] + r.js[i] * delta[i];
        if (xl < T(0)) { f += r.js[i] * (-r.D_l[i] * xl); hd += r.D_l[i]; }
      }
      grad[i] = Md[i] - f;
      gn += grad[i] * grad[i];
      H[i * (i + 1) / 2 + i] += hd;
    }
    if (am.solver_scale * t_sqrt(gn) < tol) break;
    chol6(H);
    T s[NJ];
#pragma unroll
    for (int i = 0; i < NJ; i++) s[i] = -grad[i];
    chol6_solve(H, s);
    T Ms[NJ];
    symmv6(M, s, Ms);
    T q0 = T(0), q1 = T(0), q2 = T(0);
#pragma unroll
    for (int i = 0; i < NJ; i++) { q0 += T(0.5) * delta[i] * Md[i]; q1 += s[i] * Md[i]; q2 += s[i] * Ms[i]; }
    T f0 = q0, df0 = q1, ddf0 = q2;
    arm_rows_line(am, r, delta, s, T(0), f0, df0, ddf0);
    if (df0 >= T(0) || ddf0 <= T(0)) break;
    // exact line search: safeguarded 1-D Newton on the convex piecewise-quadratic cost
    T alpha = -df0 / ddf0, lo = T(0), hi = T(-1), f = f0;
    for (int ls = 0; ls < 30; ls++) {
      T df = q1 + alpha * q2, ddf = q2;
      f = q0 + alpha * q1 + T(0.5) * alpha * alpha * q2;
      arm_rows_line(am, r, delta, s, alpha, f, df, ddf);
      if (t_abs(df) <= T(0.01) * t_abs(df0)) break;  // [upstream] mjOption.ls_tolerance = 0.01 (same constant in the oracle)
      if (df < T(0)) lo = alpha; else hi = alpha;
      T next = alpha - df / ddf;
      if (hi > T(0) && (next <= lo || next >= hi)) next = T(0.5) * (lo + hi);
      else if (hi < T(0) && next <= lo) next = T(2) * alpha;
      // float32: the derivative cannot be driven below round-off; stop when the bracket / step is at machine resolution
      if (next == alpha || t_abs(next - alpha) <= xeps * t_abs(alpha) || (hi > T(0) && hi - lo <= xeps * hi)) { alpha = next; break; }
      alpha = next;
    }
    T dmax = T(0), amax = T(1);
#pragma unroll
    for (int i = 0; i < NJ; i++) {
      const T st = alpha * s[i];
      delta[i] += st;
      dmax = t_abs(st) > dmax ? t_abs(st) : dmax;
      amax = t_abs(r.jar0_f[i]) + t_abs(delta[i]) > amax ? t_abs(r.jar0_f[i]) + t_abs(delta[i]) : amax;
    }
    // converged when the cost stops improving (MuJoCo's test) or the Newton step is below the arithmetic's resolution
    if (am.solver_scale * (f0 - f) < tol || dmax <= xeps * amax) { iter++; break; }
  }
  return iter;
}

}  // namespace so101
