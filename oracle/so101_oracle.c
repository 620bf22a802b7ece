/* so101 CPU oracle (float64, plain C) — TEST INFRASTRUCTURE ONLY, see so101_oracle.h.
 *
 * Restates, stage by stage, what the reference's env.step() makes MuJoCo compute for the SO100 scene
 * (call stack: SURVEY.md §3.1; reference so101_sim/task_suite.py:148-155 -> dm_control -> mj_step2 + mj_step1):
 *   kinematics / com            [upstream mj_kinematics, mj_comPos]       so_forward_position()
 *   mass matrix                 [upstream mj_crb]                          so_forward_position()
 *   collision                   [upstream mj_collision]                    so_collide()  (so101_collide.c)
 *   constraint rows             [upstream mj_makeConstraint/makeImpedance] make_constraint()
 *   bias forces                 [upstream mj_rne]                          rne_bias()
 *   actuation                   [upstream mj_fwdActuation]                 scene_pbr.xml:11,153-160
 *   Newton solver               [upstream mj_fwdConstraint, solver=Newton] solve_newton()
 *   semi-implicit Euler         [upstream mj_Euler]                        integrate()
 *   reward                      so100_hand_over.py:238-275, oobb_utils.py:114-273, success_detector_utils.py:19-28
 * MuJoCo's source is NOT in the reference checkout (pip dependency mujoco>=3.3.3, requirements.txt:2); its
 * published algorithm is restated from the documented computation pipeline.  Dense textbook formulas are used
 * (M = sum J^T I J, bias by a world-frame Newton-Euler pass) instead of MuJoCo's sparse spatial-vector code.
 */
#include "so101_oracle.h"

/* [upstream] mjOption.ls_tolerance = 0.01: the 1-D search stops once the slope has dropped to 1 % of its initial value;
   the outer Newton tolerance decides the accuracy of the solution. */
#define LS_TOLERANCE 0.01

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MINVAL 1e-15
#define MINIMP 0.0001
#define MAXIMP 0.9999

/* ------------------------------------------------------------------------------------------ small math */
static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(double *r, const double *a, const double *b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void mulmv3(double *r, const double *m, const double *v) { /* r = m v, m row-major 3x3 */
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2], z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void mulmtv3(double *r, const double *m, const double *v) { /* r = m^T v */
  double x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2], z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void mulmm3(double *r, const double *a, const double *b) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  memcpy(r, t, sizeof t);
}
static void quat_mul(double *r, const double *a, const double *b) {
  double t[4] = {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                 a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]};
  memcpy(r, t, sizeof t);
}
static void quat_norm(double *q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  for (int i = 0; i < 4; i++) q[i] /= n;
}
static void quat2mat(double *m, const double *q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[4] = w * w - x * x + y * y - z * z; m[8] = w * w - x * x - y * y + z * z;
  m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y); m[3] = 2 * (x * y + w * z);
  m[5] = 2 * (y * z - w * x); m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x);
}
static void quat_rot(double *r, const double *v, const double *q) { /* mju_rotVecQuat */
  double m[9]; quat2mat(m, q); mulmv3(r, m, v);
}
static void mat2quat(double *q, const double *m) { /* mju_mat2Quat */
  if (m[0] + m[4] + m[8] > 0) {
    q[0] = 0.5 * sqrt(1 + m[0] + m[4] + m[8]);
    q[1] = 0.25 * (m[7] - m[5]) / q[0]; q[2] = 0.25 * (m[2] - m[6]) / q[0]; q[3] = 0.25 * (m[3] - m[1]) / q[0];
  } else if (m[0] > m[4] && m[0] > m[8]) {
    q[1] = 0.5 * sqrt(1 + m[0] - m[4] - m[8]);
    q[0] = 0.25 * (m[7] - m[5]) / q[1]; q[2] = 0.25 * (m[1] + m[3]) / q[1]; q[3] = 0.25 * (m[2] + m[6]) / q[1];
  } else if (m[4] > m[8]) {
    q[2] = 0.5 * sqrt(1 - m[0] + m[4] - m[8]);
    q[0] = 0.25 * (m[2] - m[6]) / q[2]; q[1] = 0.25 * (m[1] + m[3]) / q[2]; q[3] = 0.25 * (m[5] + m[7]) / q[2];
  } else {
    q[3] = 0.5 * sqrt(1 - m[0] - m[4] + m[8]);
    q[0] = 0.25 * (m[3] - m[1]) / q[3]; q[1] = 0.25 * (m[2] + m[6]) / q[3]; q[2] = 0.25 * (m[5] + m[7]) / q[3];
  }
  quat_norm(q);
}

/* ------------------------------------------------------------------------------------------ model blob */
typedef struct { char name[24]; uint32_t dtype, count; uint64_t offset; } blob_entry;

static const void *blob_find(const char *base, const char *name, uint32_t *count, int dtype) {
  uint32_t n = *(const uint32_t *)(base + 8);
  const blob_entry *e = (const blob_entry *)(base + 16);
  for (uint32_t i = 0; i < n; i++)
    if (strncmp(e[i].name, name, 24) == 0) {
      if ((int)e[i].dtype != dtype) { fprintf(stderr, "so101_oracle: blob entry %s has wrong dtype\n", name); return NULL; }
      if (count) *count = e[i].count;
      return base + e[i].offset;
    }
  fprintf(stderr, "so101_oracle: blob entry %s missing\n", name);
  return NULL;
}
#define BF(field) m->field = (const double *)blob_find(b, #field, NULL, 0)
#define BI(field) m->field = (const int *)blob_find(b, #field, NULL, 1)
static int blob_int(const char *b, const char *name) { const int *p = (const int *)blob_find(b, name, NULL, 1); return p ? *p : 0; }

so_model *so_model_load(const void *blob, size_t len) {
  if (len < 16 || memcmp(blob, "SO1B", 4) != 0) return NULL;
  so_model *m = (so_model *)calloc(1, sizeof(so_model));
  char *b = (char *)malloc(len);
  memcpy(b, blob, len);
  m->blob_copy = b;
  m->nq = blob_int(b, "nq"); m->nv = blob_int(b, "nv"); m->nu = blob_int(b, "nu"); m->nbody = blob_int(b, "nbody");
  m->njnt = blob_int(b, "njnt"); m->ngeom = blob_int(b, "ngeom"); m->nprop = blob_int(b, "nprop");
  const double *opt = (const double *)blob_find(b, "opt", NULL, 0);
  m->timestep = opt[0]; m->gravity[0] = opt[1]; m->gravity[1] = opt[2]; m->gravity[2] = opt[3];
  m->impratio = opt[4]; m->tolerance = opt[5]; m->iterations = (int)opt[6]; m->elliptic = (int)opt[7]; m->meaninertia = opt[8];
  BI(body_parent); BI(body_weld); BI(body_geomadr); BI(body_geomnum); BI(body_jntadr); BI(body_jntnum);
  BF(body_pos); BF(body_quat); BF(body_ipos); BF(body_iquat); BF(body_mass); BF(body_inertia); BF(body_invweight0);
  BF(body_bcenter); BF(body_rbound);
  BI(jnt_type); BI(jnt_body); BI(jnt_qposadr); BI(jnt_dofadr); BI(jnt_limited);
  BF(jnt_pos); BF(jnt_axis); BF(jnt_range); BF(jnt_solreflimit); BF(jnt_solimplimit); BF(jnt_solreffriction); BF(jnt_solimpfriction);
  BF(jnt_margin);
  BI(dof_body); BI(dof_jnt); BF(dof_armature); BF(dof_frictionloss); BF(dof_damping); BF(dof_invweight0); BF(qpos0);
  BI(act_jnt); BF(act_gain); BF(act_bias); BF(act_ctrlrange); BF(act_forcerange); BF(act_gear);
  BI(geom_type); BI(geom_body); BI(geom_condim); BI(geom_priority); BI(geom_vertadr); BI(geom_vertnum); BI(geom_faceadr); BI(geom_facenum);
  BF(geom_pos); BF(geom_mat); BF(geom_size); BF(geom_bcenter); BF(geom_rbound); BF(geom_friction); BF(geom_solref); BF(geom_solimp);
  BF(geom_solmix); BF(geom_margin); BF(geom_gap);
  BF(hull_vert); BI(hull_face); BI(hull_nbradr); BI(hull_nbr);
  uint32_t c = 0;
  m->bodypair = (const int *)blob_find(b, "bodypair", &c, 1); m->npair = (int)c / 2;
  blob_find(b, "hull_vert", &c, 0); m->nvert = (int)c / 3;
  BI(prop_body); BF(reward_obj_box); BF(reward_box_pos); BF(reward_box_half);
  blob_find(b, "reward_box_pos", &c, 0); m->nreward_box = (int)c / 3;
  if (m->nq > SO_NQMAX || m->nv > SO_NVMAX || m->nbody > SO_NBMAX || m->nu > SO_NUMAX) { so_model_free(m); return NULL; }
  return m;
}
void so_model_free(so_model *m) { if (m) { free(m->blob_copy); free(m); } }
so_data *so_data_new(const so_model *m) {
  so_data *d = (so_data *)calloc(1, sizeof(so_data));
  d->collide_enabled = 1;
  so_reset(m, d);
  return d;
}
void so_data_free(so_data *d) { free(d); }
void so_reset(const so_model *m, so_data *d) {
  int ce = d->collide_enabled, ig = d->integrator;
  memset(d, 0, sizeof(so_data));
  d->collide_enabled = ce; d->integrator = ig;
  memcpy(d->qpos, m->qpos0, sizeof(double) * m->nq);
}

/* ------------------------------------------------------------------------------------------ kinematics */
static void kinematics(const so_model *m, so_data *d) {
  d->xpos[0][0] = d->xpos[0][1] = d->xpos[0][2] = 0;
  d->xquat[0][0] = 1; d->xquat[0][1] = d->xquat[0][2] = d->xquat[0][3] = 0;
  quat2mat(d->xmat[0], d->xquat[0]);
  memcpy(d->xipos[0], d->xpos[0], sizeof d->xpos[0]); memcpy(d->ximat[0], d->xmat[0], sizeof d->xmat[0]);
  for (int i = 1; i < m->nbody; i++) {
    int p = m->body_parent[i];
    double *xp = d->xpos[i], *xq = d->xquat[i];
    int j0 = m->body_jntadr[i], nj = m->body_jntnum[i];
    if (nj == 1 && m->jnt_type[j0] == SO_JNT_FREE) {
      const double *q = d->qpos + m->jnt_qposadr[j0];
      xp[0] = q[0]; xp[1] = q[1]; xp[2] = q[2];
      xq[0] = q[3]; xq[1] = q[4]; xq[2] = q[5]; xq[3] = q[6];
      quat_norm(xq);
      quat2mat(d->xmat[i], xq);
      int dof = m->jnt_dofadr[j0];
      for (int k = 0; k < 3; k++) { /* translational dofs: world axes; rotational dofs: body-local axes */
        d->dof_trans[dof + k] = 1; d->dof_trans[dof + 3 + k] = 0;
        for (int c = 0; c < 3; c++) {
          d->dof_axis[dof + k][c] = (c == k); d->dof_anchor[dof + k][c] = 0;
          d->dof_axis[dof + 3 + k][c] = d->xmat[i][3 * c + k]; d->dof_anchor[dof + 3 + k][c] = xp[c];
        }
      }
    } else {
      double t[3];
      mulmv3(t, d->xmat[p], m->body_pos + 3 * i);
      for (int c = 0; c < 3; c++) xp[c] = d->xpos[p][c] + t[c];
      quat_mul(xq, d->xquat[p], m->body_quat + 4 * i);
      for (int j = j0; j < j0 + nj; j++) { /* hinge joints: rotate about the anchor by qpos - qpos0 */
        double mat[9], anchor[3], axis_w[3];
        quat2mat(mat, xq);
        mulmv3(t, mat, m->jnt_pos + 3 * j);
        for (int c = 0; c < 3; c++) anchor[c] = xp[c] + t[c];
        mulmv3(axis_w, mat, m->jnt_axis + 3 * j);
        double ang = d->qpos[m->jnt_qposadr[j]] - m->qpos0[m->jnt_qposadr[j]];
        double s = sin(0.5 * ang), qr[4] = {cos(0.5 * ang), m->jnt_axis[3 * j] * s, m->jnt_axis[3 * j + 1] * s, m->jnt_axis[3 * j + 2] * s};
        quat_mul(xq, xq, qr);
        quat2mat(mat, xq);
        mulmv3(t, mat, m->jnt_pos + 3 * j);
        for (int c = 0; c < 3; c++) xp[c] = anchor[c] - t[c];
        int dof = m->jnt_dofadr[j];
        d->dof_trans[dof] = 0;
        memcpy(d->dof_axis[dof], axis_w, sizeof axis_w); memcpy(d->dof_anchor[dof], anchor, sizeof anchor);
      }
      quat_norm(xq);
      quat2mat(d->xmat[i], xq);
    }
    double t[3], im[9];
    mulmv3(t, d->xmat[i], m->body_ipos + 3 * i);
    for (int c = 0; c < 3; c++) d->xipos[i][c] = xp[c] + t[c];
    quat2mat(im, m->body_iquat + 4 * i);
    mulmm3(d->ximat[i], d->xmat[i], im);
  }
}

static int is_ancestor_dof(const so_model *m, int body, int dof) { /* does `dof` move `body`? */
  int b = body, db = m->dof_body[dof];
  while (b != 0) { if (b == db) return 1; b = m->body_parent[b]; }
  return 0;
}

void so_jac(const so_model *m, const so_data *d, int body, const double point[3], double *jacp, double *jacr) {
  int nv = m->nv;
  memset(jacp, 0, sizeof(double) * 3 * nv); memset(jacr, 0, sizeof(double) * 3 * nv);
  for (int k = 0; k < nv; k++) {
    if (!is_ancestor_dof(m, body, k)) continue;
    if (d->dof_trans[k]) {
      for (int c = 0; c < 3; c++) jacp[c * nv + k] = d->dof_axis[k][c];
    } else {
      double r[3] = {point[0] - d->dof_anchor[k][0], point[1] - d->dof_anchor[k][1], point[2] - d->dof_anchor[k][2]}, t[3];
      cross3(t, d->dof_axis[k], r);
      for (int c = 0; c < 3; c++) { jacr[c * nv + k] = d->dof_axis[k][c]; jacp[c * nv + k] = t[c]; }
    }
  }
}

static void mass_matrix(const so_model *m, so_data *d) {
  int nv = m->nv;
  double jp[3 * SO_NVMAX], jr[3 * SO_NVMAX];
  memset(d->M, 0, sizeof(double) * nv * nv);
  for (int k = 0; k < nv; k++) d->M[k * nv + k] = m->dof_armature[k];
  for (int b = 1; b < m->nbody; b++) {
    double mass = m->body_mass[b];
    if (mass <= 0 || m->body_weld[b] == 0) continue;
    so_jac(m, d, b, d->xipos[b], jp, jr);
    double Iw[9], t[9], dg[9] = {m->body_inertia[3 * b], 0, 0, 0, m->body_inertia[3 * b + 1], 0, 0, 0, m->body_inertia[3 * b + 2]};
    mulmm3(t, d->ximat[b], dg);
    double imT[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) imT[3 * i + j] = d->ximat[b][3 * j + i];
    mulmm3(Iw, t, imT);
    for (int i = 0; i < nv; i++)
      for (int j = 0; j < nv; j++) {
        double s = 0;
        for (int c = 0; c < 3; c++) {
          s += mass * jp[c * nv + i] * jp[c * nv + j];
          for (int e = 0; e < 3; e++) s += jr[c * nv + i] * Iw[3 * c + e] * jr[e * nv + j];
        }
        d->M[i * nv + j] += s;
      }
  }
}

/* world-frame Newton-Euler pass with qacc = 0: qfrc_bias = C(q,qd) qd - gravity forces */
static void rne_bias(const so_model *m, so_data *d) {
  int nv = m->nv, nb = m->nbody;
  double w[SO_NBMAX][3] = {{0}}, al[SO_NBMAX][3] = {{0}}, vo[SO_NBMAX][3] = {{0}}, ao[SO_NBMAX][3] = {{0}};
  double jp[3 * SO_NVMAX], jr[3 * SO_NVMAX];
  memset(d->qfrc_bias, 0, sizeof(double) * nv);
  for (int i = 1; i < nb; i++) {
    int p = m->body_parent[i], j0 = m->body_jntadr[i], nj = m->body_jntnum[i];
    if (nj == 1 && m->jnt_type[j0] == SO_JNT_FREE) {
      int dof = m->jnt_dofadr[j0];
      mulmv3(w[i], d->xmat[i], d->qvel + dof + 3);
      for (int c = 0; c < 3; c++) { vo[i][c] = d->qvel[dof + c]; al[i][c] = 0; ao[i][c] = 0; }
    } else {
      /* quantities of the point of the parent that coincides with this body's origin path are rebuilt joint by joint */
      double wi[3], ali[3];
      memcpy(wi, w[p], sizeof wi); memcpy(ali, al[p], sizeof ali);
      /* carry a reference point: start at the parent origin */
      double ref[3], vref[3], aref[3];
      memcpy(ref, d->xpos[p], sizeof ref); memcpy(vref, vo[p], sizeof vref); memcpy(aref, ao[p], sizeof aref);
      for (int j = j0; j < j0 + nj; j++) {
        int dof = m->jnt_dofadr[j];
        const double *ax = d->dof_axis[dof], *an = d->dof_anchor[dof];
        /* move reference point to the anchor with the current (pre-joint) angular motion */
        double r[3] = {an[0] - ref[0], an[1] - ref[1], an[2] - ref[2]}, t1[3], t2[3];
        cross3(t1, wi, r); cross3(t2, wi, t1);
        double t3[3]; cross3(t3, ali, r);
        for (int c = 0; c < 3; c++) { vref[c] += t1[c]; aref[c] += t3[c] + t2[c]; ref[c] = an[c]; }
        /* add joint motion */
        double qd = d->qvel[dof], wxa[3];
        cross3(wxa, wi, ax);
        for (int c = 0; c < 3; c++) { ali[c] += wxa[c] * qd; wi[c] += ax[c] * qd; }
      }
      double r[3] = {d->xpos[i][0] - ref[0], d->xpos[i][1] - ref[1], d->xpos[i][2] - ref[2]}, t1[3], t2[3], t3[3];
      cross3(t1, wi, r); cross3(t2, wi, t1); cross3(t3, ali, r);
      for (int c = 0; c < 3; c++) { vo[i][c] = vref[c] + t1[c]; ao[i][c] = aref[c] + t3[c] + t2[c]; w[i][c] = wi[c]; al[i][c] = ali[c]; }
    }
    double mass = m->body_mass[i];
    if (mass <= 0 || m->body_weld[i] == 0) continue;
    /* com acceleration, inertial force and torque */
    double c3[3] = {d->xipos[i][0] - d->xpos[i][0], d->xipos[i][1] - d->xpos[i][1], d->xipos[i][2] - d->xpos[i][2]};
    double t1[3], t2[3], t3[3], f[3], tau[3];
    cross3(t1, w[i], c3); cross3(t2, w[i], t1); cross3(t3, al[i], c3);
    for (int c = 0; c < 3; c++) f[c] = mass * (ao[i][c] + t3[c] + t2[c] - m->gravity[c]);
    double wl[3], all[3], Iw[3], Ia[3];
    mulmtv3(wl, d->ximat[i], w[i]); mulmtv3(all, d->ximat[i], al[i]);
    for (int c = 0; c < 3; c++) { Iw[c] = m->body_inertia[3 * i + c] * wl[c]; Ia[c] = m->body_inertia[3 * i + c] * all[c]; }
    double wxIw[3], tl[3];
    cross3(wxIw, wl, Iw);
    for (int c = 0; c < 3; c++) tl[c] = Ia[c] + wxIw[c];
    mulmv3(tau, d->ximat[i], tl);
    so_jac(m, d, i, d->xipos[i], jp, jr);
    for (int k = 0; k < nv; k++)
      for (int c = 0; c < 3; c++) d->qfrc_bias[k] += jp[c * nv + k] * f[c] + jr[c * nv + k] * tau[c];
  }
}

/* ------------------------------------------------------------------------------------------ linear algebra */
static int cholesky(double *A, int n) { /* in place lower Cholesky, row-major; returns rank deficiency count */
  int bad = 0;
  for (int j = 0; j < n; j++) {
    double s = A[j * n + j];
    for (int k = 0; k < j; k++) s -= A[j * n + k] * A[j * n + k];
    if (s < MINVAL) { s = MINVAL; bad++; }
    s = sqrt(s);
    A[j * n + j] = s;
    for (int i = j + 1; i < n; i++) {
      double t = A[i * n + j];
      for (int k = 0; k < j; k++) t -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = t / s;
    }
  }
  return bad;
}
static void chol_solve(const double *L, int n, double *x) {
  for (int i = 0; i < n; i++) { double s = x[i]; for (int k = 0; k < i; k++) s -= L[i * n + k] * x[k]; x[i] = s / L[i * n + i]; }
  for (int i = n - 1; i >= 0; i--) { double s = x[i]; for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k]; x[i] = s / L[i * n + i]; }
}

/* ------------------------------------------------------------------------------------------ constraint rows */
static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* [upstream] getimpedance: power-law sigmoid between d0 and dmax over `width` */
static double impedance(const double *solimp, double pos, double margin) {
  double d0 = clampd(solimp[0], MINIMP, MAXIMP), dmax = clampd(solimp[1], MINIMP, MAXIMP);
  double width = solimp[2] > MINVAL ? solimp[2] : MINVAL, mid = clampd(solimp[3], MINIMP, MAXIMP), power = solimp[4] > 1 ? solimp[4] : 1;
  double x = fabs(pos - margin) / width;
  if (x >= 1) return dmax;
  if (x <= 0) return d0;
  double y;
  if (power == 1) y = x;
  else if (x <= mid) y = pow(1 / mid, power - 1) * pow(x, power);
  else y = 1 - pow(1 / (1 - mid), power - 1) * pow(1 - x, power);
  return d0 + y * (dmax - d0);
}
/* [upstream] getsolparam: stiffness K and damping B of the reference acceleration */
static void solparam(const so_model *m, const double *solref, const double *solimp, double *K, double *B) {
  double dmax = clampd(solimp[1], MINIMP, MAXIMP);
  if (solref[0] > 0) {
    double tc = solref[0] > 2 * m->timestep ? solref[0] : 2 * m->timestep, dr = solref[1];
    *K = 1 / (dmax * dmax * tc * tc * dr * dr); *B = 2 / (dmax * tc);
  } else { *K = -solref[0] / (dmax * dmax); *B = -solref[1] / dmax; }
}

static void add_row(const so_model *m, so_data *d, int type, int id, const double *J, double pos, double margin, double floss,
                    double diagApprox, const double *solref, const double *solimp, int friction_row) {
  int i = d->nefc, nv = m->nv;
  if (i >= SO_NEFCMAX) return;
  d->efc_type[i] = type; d->efc_id[i] = id;
  memcpy(d->efc_J + i * nv, J, sizeof(double) * nv);
  d->efc_pos[i] = pos; d->efc_margin[i] = margin; d->efc_frictionloss[i] = floss; d->efc_diagApprox[i] = diagApprox;
  double vel = 0;
  for (int k = 0; k < nv; k++) vel += J[k] * d->qvel[k];
  d->efc_vel[i] = vel;
  double imp = impedance(solimp, pos, margin), K, B;
  solparam(m, solref, solimp, &K, &B);
  if (friction_row) K = 0;
  double R = (1 - imp) / imp * diagApprox;
  d->efc_R[i] = R > MINVAL ? R : MINVAL;
  d->efc_aref[i] = -B * vel - K * imp * (pos - margin);
  d->nefc++;
}

static void make_constraint(const so_model *m, so_data *d) {
  int nv = m->nv;
  double J[SO_NVMAX];
  d->nefc = 0;
  /* 1. dof friction loss (always-on rows for the arm: scene_pbr.xml:10) */
  for (int k = 0; k < nv; k++) {
    if (m->dof_frictionloss[k] <= 0) continue;
    memset(J, 0, sizeof J); J[k] = 1;
    int j = m->dof_jnt[k];
    add_row(m, d, SO_ROW_FRICTION, k, J, 0, 0, m->dof_frictionloss[k], m->dof_invweight0[k], m->jnt_solreffriction + 2 * j,
            m->jnt_solimpfriction + 5 * j, 1);
  }
  d->ne_fric = d->nefc;
  /* 2. joint limits: active when dist < margin */
  for (int j = 0; j < m->njnt; j++) {
    if (m->jnt_type[j] != SO_JNT_HINGE || !m->jnt_limited[j]) continue;
    int k = m->jnt_dofadr[j];
    double q = d->qpos[m->jnt_qposadr[j]], margin = m->jnt_margin[j];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->jnt_range[2 * j + (side + 1) / 2] - q);
      if (dist < margin) {
        memset(J, 0, sizeof J); J[k] = -side;
        add_row(m, d, SO_ROW_LIMIT, j, J, dist, margin, 0, m->dof_invweight0[k], m->jnt_solreflimit + 2 * j, m->jnt_solimplimit + 5 * j, 0);
      }
    }
  }
  d->ne_limit = d->nefc - d->ne_fric;
  /* 3. contacts: elliptic cones, one block of `dim` rows per contact */
  double jp1[3 * SO_NVMAX], jr1[3 * SO_NVMAX], jp2[3 * SO_NVMAX], jr2[3 * SO_NVMAX];
  for (int c = 0; c < d->ncon; c++) {
    so_contact *con = d->contact + c;
    con->efc_address = -1;
    if (con->dist >= con->includemargin) continue;
    if (d->nefc + con->dim > SO_NEFCMAX) break;
    int b1 = m->geom_body[con->geom1], b2 = m->geom_body[con->geom2];
    so_jac(m, d, b1, con->pos, jp1, jr1); so_jac(m, d, b2, con->pos, jp2, jr2);
    double tran = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2], rot = m->body_invweight0[2 * b1 + 1] + m->body_invweight0[2 * b2 + 1];
    int first = d->nefc;
    con->efc_address = first;
    for (int r = 0; r < con->dim; r++) {
      const double *ax = con->frame + 3 * (r % 3);
      for (int k = 0; k < nv; k++) {
        double s = 0;
        if (r < 3) for (int e = 0; e < 3; e++) s += ax[e] * (jp2[e * nv + k] - jp1[e * nv + k]);
        else for (int e = 0; e < 3; e++) s += ax[e] * (jr2[e * nv + k] - jr1[e * nv + k]);
        J[k] = s;
      }
      add_row(m, d, r == 0 ? SO_ROW_CONTACT : SO_ROW_CONTACT_FR, c, J, r == 0 ? con->dist : 0, r == 0 ? con->includemargin : 0, 0,
              r < 3 ? tran : rot, con->solref, con->solimp, r > 0);
    }
    /* [upstream] elliptic: friction regularisation from impratio, cone coefficient mu */
    if (con->dim > 1) {
      double ir = m->impratio > MINVAL ? m->impratio : MINVAL;
      d->efc_R[first + 1] = d->efc_R[first] / ir;
      for (int r = 2; r < con->dim; r++)
        d->efc_R[first + r] = d->efc_R[first + 1] * con->friction[0] * con->friction[0] / (con->friction[r - 1] * con->friction[r - 1]);
      con->mu = con->friction[0] * sqrt(d->efc_R[first + 1] / d->efc_R[first]);
    } else con->mu = 0;
  }
  for (int i = 0; i < d->nefc; i++) d->efc_D[i] = 1 / d->efc_R[i];
}

/* ------------------------------------------------------------------------------------------ Newton solver */
/* cost / force of one elliptic contact given jar (dim values).  Optionally the dim x dim Hessian.  Returns zone:
 * 0 = satisfied (zero), 1 = quadratic (bottom), 2 = cone (middle). */
static int cone_eval(const so_contact *con, const double *D, const double *jar, double *cost, double *force, double *H) {
  int dim = con->dim;
  double mu = con->mu, U[6], S[6];
  S[0] = mu; U[0] = jar[0] * mu;
  double T2 = 0;
  for (int j = 1; j < dim; j++) { S[j] = con->friction[j - 1]; U[j] = jar[j] * S[j]; T2 += U[j] * U[j]; }
  double N = U[0], T = sqrt(T2);
  *cost = 0;
  for (int j = 0; j < dim; j++) force[j] = 0;
  if (H) memset(H, 0, sizeof(double) * dim * dim);
  if (dim == 1) { /* frictionless */
    if (jar[0] >= 0) return 0;
    *cost = 0.5 * D[0] * jar[0] * jar[0]; force[0] = -D[0] * jar[0];
    if (H) H[0] = D[0];
    return 1;
  }
  if (N >= mu * T || (T <= 0 && N >= 0)) return 0;                      /* top zone */
  if (mu * N + T <= 0 || (T <= 0 && N < 0)) {                           /* bottom zone */
    for (int j = 0; j < dim; j++) { *cost += 0.5 * D[j] * jar[j] * jar[j]; force[j] = -D[j] * jar[j]; if (H) H[j * dim + j] = D[j]; }
    return 1;
  }
  double Dm = D[0] / (mu * mu * (1 + mu * mu)), NmT = N - mu * T;      /* middle zone */
  *cost = 0.5 * Dm * NmT * NmT;
  force[0] = -Dm * NmT * mu;
  for (int j = 1; j < dim; j++) force[j] = -force[0] / T * U[j] * S[j];
  if (H) {
    double g[6]; /* d(N - mu T)/dU */
    g[0] = 1;
    for (int j = 1; j < dim; j++) g[j] = -mu * U[j] / T;
    for (int a = 0; a < dim; a++)
      for (int b = 0; b < dim; b++) {
        double h = Dm * g[a] * g[b];
        if (a > 0 && b > 0) h += Dm * NmT * (-mu) * ((a == b ? 1.0 / T : 0.0) - U[a] * U[b] / (T * T * T));
        H[a * dim + b] = h * S[a] * S[b];
      }
  }
  return 2;
}

/* evaluate constraint cost, forces (and optionally accumulate J^T H J into Hq, nv x nv) at jar */
static double constraint_update(const so_model *m, so_data *d, const double *jar, double *force, double *Hq) {
  int nv = m->nv;
  double cost = 0;
  for (int i = 0; i < d->nefc; i++) {
    int type = d->efc_type[i];
    double D = d->efc_D[i], R = d->efc_R[i];
    if (type == SO_ROW_FRICTION) {
      double f = d->efc_frictionloss[i], rf = R * f, h = 0;
      if (jar[i] <= -rf) { cost += f * (-0.5 * rf - jar[i]); force[i] = f; }
      else if (jar[i] >= rf) { cost += f * (-0.5 * rf + jar[i]); force[i] = -f; }
      else { cost += 0.5 * D * jar[i] * jar[i]; force[i] = -D * jar[i]; h = D; }
      if (Hq && h > 0) { const double *J = d->efc_J + i * nv; for (int a = 0; a < nv; a++) for (int b = 0; b < nv; b++) Hq[a * nv + b] += h * J[a] * J[b]; }
    } else if (type == SO_ROW_LIMIT) {
      if (jar[i] < 0) {
        cost += 0.5 * D * jar[i] * jar[i]; force[i] = -D * jar[i];
        if (Hq) { const double *J = d->efc_J + i * nv; for (int a = 0; a < nv; a++) for (int b = 0; b < nv; b++) Hq[a * nv + b] += D * J[a] * J[b]; }
      } else force[i] = 0;
    } else if (type == SO_ROW_CONTACT) {
      const so_contact *con = d->contact + d->efc_id[i];
      int dim = con->dim;
      double c, Hc[36];
      int zone = cone_eval(con, d->efc_D + i, jar + i, &c, force + i, Hq ? Hc : NULL);
      cost += c;
      if (Hq && zone != 0) {
        /* Hq += Jc^T Hc Jc */
        double tmp[6 * SO_NVMAX];
        for (int a = 0; a < dim; a++) for (int k = 0; k < nv; k++) {
          double s = 0;
          for (int b = 0; b < dim; b++) s += Hc[a * dim + b] * d->efc_J[(i + b) * nv + k];
          tmp[a * nv + k] = s;
        }
        for (int k = 0; k < nv; k++) for (int l = 0; l < nv; l++) {
          double s = 0;
          for (int a = 0; a < dim; a++) s += d->efc_J[(i + a) * nv + k] * tmp[a * nv + l];
          Hq[k * nv + l] += s;
        }
      }
      i += dim - 1;
    }
  }
  return cost;
}

/* total cost at qacc: Gauss term + constraint term; also jar, force */
static double total_cost(const so_model *m, so_data *d, const double *qacc, double *jar, double *force, double *Hq, double *Ma_out) {
  int nv = m->nv;
  double Ma[SO_NVMAX], gauss = 0;
  for (int i = 0; i < nv; i++) { double s = 0; for (int j = 0; j < nv; j++) s += d->M[i * nv + j] * qacc[j]; Ma[i] = s; }
  for (int i = 0; i < nv; i++) gauss += 0.5 * (Ma[i] - d->qfrc_smooth[i]) * (qacc[i] - d->qacc_smooth[i]);
  for (int i = 0; i < d->nefc; i++) {
    double s = -d->efc_aref[i];
    for (int k = 0; k < nv; k++) s += d->efc_J[i * nv + k] * qacc[k];
    jar[i] = s;
  }
  if (Ma_out) memcpy(Ma_out, Ma, sizeof(double) * nv);
  return gauss + constraint_update(m, d, jar, force, Hq);
}

/* 1-D cost along qacc + alpha * search: value, first and second derivative */
static void line_eval(const so_model *m, so_data *d, const double *jar, const double *jv, double alpha, double quadGauss[3], double *f,
                      double *df, double *ddf) {
  double c = quadGauss[0] + alpha * quadGauss[1] + 0.5 * alpha * alpha * quadGauss[2];
  double g = quadGauss[1] + alpha * quadGauss[2], h = quadGauss[2];
  for (int i = 0; i < d->nefc; i++) {
    int type = d->efc_type[i];
    double x = jar[i] + alpha * jv[i], D = d->efc_D[i], R = d->efc_R[i];
    if (type == SO_ROW_FRICTION) {
      double fl = d->efc_frictionloss[i], rf = R * fl;
      if (x <= -rf) { c += fl * (-0.5 * rf - x); g += -fl * jv[i]; }
      else if (x >= rf) { c += fl * (-0.5 * rf + x); g += fl * jv[i]; }
      else { c += 0.5 * D * x * x; g += D * x * jv[i]; h += D * jv[i] * jv[i]; }
    } else if (type == SO_ROW_LIMIT) {
      if (x < 0) { c += 0.5 * D * x * x; g += D * x * jv[i]; h += D * jv[i] * jv[i]; }
    } else if (type == SO_ROW_CONTACT) {
      const so_contact *con = d->contact + d->efc_id[i];
      int dim = con->dim;
      double xx[6], force[6], Hc[36], cc;
      for (int j = 0; j < dim; j++) xx[j] = jar[i + j] + alpha * jv[i + j];
      cone_eval(con, d->efc_D + i, xx, &cc, force, Hc);
      c += cc;
      for (int a = 0; a < dim; a++) {
        g -= force[a] * jv[i + a];
        for (int b = 0; b < dim; b++) h += jv[i + a] * Hc[a * dim + b] * jv[i + b];
      }
      i += dim - 1;
    }
  }
  *f = c; *df = g; *ddf = h;
}

static void solve_newton(const so_model *m, so_data *d) {
  int nv = m->nv, nefc = d->nefc;
  double jar[SO_NEFCMAX], force[SO_NEFCMAX], jv[SO_NEFCMAX];
  double qacc[SO_NVMAX], grad[SO_NVMAX], search[SO_NVMAX], Ma[SO_NVMAX], H[SO_NVMAX * SO_NVMAX], Mv[SO_NVMAX];
  double scale = 1 / (m->meaninertia * (nv > 1 ? nv : 1));
  /* warm start: the better of qacc_warmstart and qacc_smooth */
  double cw = total_cost(m, d, d->qacc_warmstart, jar, force, NULL, NULL);
  double cs = total_cost(m, d, d->qacc_smooth, jar, force, NULL, NULL);
  memcpy(qacc, cw < cs ? d->qacc_warmstart : d->qacc_smooth, sizeof(double) * nv);
  int iter = 0;
  double cost = 0;
  for (; iter < m->iterations; iter++) {
    memset(H, 0, sizeof(double) * nv * nv);
    cost = total_cost(m, d, qacc, jar, force, H, Ma);
    double gnorm = 0;
    for (int i = 0; i < nv; i++) {
      double s = Ma[i] - d->qfrc_smooth[i];
      for (int r = 0; r < nefc; r++) s -= d->efc_J[r * nv + i] * force[r];
      grad[i] = s; gnorm += s * s;
    }
    gnorm = sqrt(gnorm);
    if (scale * gnorm < m->tolerance) break;
    for (int i = 0; i < nv * nv; i++) H[i] += d->M[i];
    cholesky(H, nv);
    for (int i = 0; i < nv; i++) search[i] = -grad[i];
    chol_solve(H, nv, search);
    /* exact line search on the convex 1-D cost (safeguarded Newton) */
    for (int i = 0; i < nv; i++) { double s = 0; for (int j = 0; j < nv; j++) s += d->M[i * nv + j] * search[j]; Mv[i] = s; }
    double quadGauss[3] = {0, 0, 0};
    for (int i = 0; i < nv; i++) {
      quadGauss[0] += 0.5 * (Ma[i] - d->qfrc_smooth[i]) * (qacc[i] - d->qacc_smooth[i]);
      quadGauss[1] += search[i] * (Ma[i] - d->qfrc_smooth[i]);
      quadGauss[2] += search[i] * Mv[i];
    }
    for (int r = 0; r < nefc; r++) { double s = 0; for (int k = 0; k < nv; k++) s += d->efc_J[r * nv + k] * search[k]; jv[r] = s; }
    double f0, df0, ddf0, f, df, ddf;
    line_eval(m, d, jar, jv, 0, quadGauss, &f0, &df0, &ddf0);
    if (df0 >= 0 || ddf0 <= 0) break;
    double alpha = -df0 / ddf0, lo = 0, hi = -1, dlo = df0;
    for (int ls = 0; ls < 60; ls++) {
      line_eval(m, d, jar, jv, alpha, quadGauss, &f, &df, &ddf);
      if (fabs(df) <= LS_TOLERANCE * fabs(df0)) break;
      if (df < 0) { lo = alpha; dlo = df; } else hi = alpha;
      double next = alpha - df / ddf;
      if (hi > 0 && (next <= lo || next >= hi)) next = 0.5 * (lo + hi);
      else if (hi < 0 && next <= lo) next = 2 * alpha + 1e-12;
      alpha = next;
    }
    (void)dlo;
    for (int i = 0; i < nv; i++) qacc[i] += alpha * search[i];
    double newcost = f;
    double improvement = scale * (cost - newcost);
    if (improvement < m->tolerance) { iter++; break; }
  }
  cost = total_cost(m, d, qacc, jar, force, NULL, NULL);
  d->solver_iter = iter; d->solver_cost = cost;
  memcpy(d->qacc, qacc, sizeof(double) * nv);
  memcpy(d->efc_force, force, sizeof(double) * nefc);
  for (int i = 0; i < nv; i++) { double s = 0; for (int r = 0; r < nefc; r++) s += d->efc_J[r * nv + i] * force[r]; d->qfrc_constraint[i] = s; }
}

/* ------------------------------------------------------------------------------------------ step */
void so_forward_position(const so_model *m, so_data *d) {
  kinematics(m, d);
  mass_matrix(m, d);
  d->ncon = 0;
  if (d->collide_enabled) so_collide(m, d);
  make_constraint(m, d);
}

/* [upstream] mj_implicit with integrator = implicitfast: qvel += h x with (M - h D) x = M qacc, D = d(qfrc_smooth)/d(qvel) without
 * the Coriolis terms.  In this model only the actuators depend on velocity (affine bias, biasprm[2] = +1, scene_pbr.xml:11;
 * no joint damping): D = diag(gear^2 * biasprm[2]) over the actuated dofs whose force is not clamped by forcerange
 * ([upstream] mjd_actuator_vel skips clamped actuators).  qacc_warmstart keeps the forward-dynamics qacc. */
static void implicitfast_qacc(const so_model *m, const so_data *d, double *x) {
  int nv = m->nv;
  double h = m->timestep, A[SO_NVMAX * SO_NVMAX], D[SO_NVMAX];
  for (int k = 0; k < nv; k++) D[k] = 0;
  for (int a = 0; a < m->nu; a++)
    if (!d->act_clamped[a]) D[m->jnt_dofadr[m->act_jnt[a]]] += m->act_gear[a] * m->act_gear[a] * m->act_bias[3 * a + 2];
  for (int i = 0; i < nv; i++) {
    double s = 0;
    for (int j = 0; j < nv; j++) { s += d->M[i * nv + j] * d->qacc[j]; A[i * nv + j] = d->M[i * nv + j] - (i == j ? h * D[i] : 0.0); }
    x[i] = s;
  }
  cholesky(A, nv);
  chol_solve(A, nv, x);
}

static void integrate(const so_model *m, so_data *d) { /* [upstream] mj_Euler without joint damping / mj_implicit (implicitfast) */
  double h = m->timestep;
  if (d->integrator == 1) {
    double x[SO_NVMAX];
    implicitfast_qacc(m, d, x);
    for (int k = 0; k < m->nv; k++) d->qvel[k] += h * x[k];
  } else
  for (int k = 0; k < m->nv; k++) d->qvel[k] += h * d->qacc[k];
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == SO_JNT_HINGE) d->qpos[qa] += h * d->qvel[da];
    else {
      for (int c = 0; c < 3; c++) d->qpos[qa + c] += h * d->qvel[da + c];
      double *q = d->qpos + qa + 3, w[3] = {d->qvel[da + 3], d->qvel[da + 4], d->qvel[da + 5]};
      double ang = sqrt(dot3(w, w)) * h;
      if (ang > 0) {
        double n = sqrt(dot3(w, w)), s = sin(0.5 * ang), qr[4] = {cos(0.5 * ang), w[0] / n * s, w[1] / n * s, w[2] / n * s};
        quat_mul(q, q, qr);
      }
      quat_norm(q);
    }
  }
  d->time += h;
}

void so_substep(const so_model *m, so_data *d) {
  int nv = m->nv;
  so_forward_position(m, d);
  rne_bias(m, d);
  memset(d->qfrc_actuator, 0, sizeof(double) * nv);
  for (int a = 0; a < m->nu; a++) { /* scene_pbr.xml:11 — general actuator, fixed gain, affine bias, ctrl and force clamps */
    int j = m->act_jnt[a], k = m->jnt_dofadr[j];
    double g = m->act_gear[a], len = g * d->qpos[m->jnt_qposadr[j]], vel = g * d->qvel[k];
    double c = clampd(d->ctrl[a], m->act_ctrlrange[2 * a], m->act_ctrlrange[2 * a + 1]);
    double f = m->act_gain[a] * c + m->act_bias[3 * a] + m->act_bias[3 * a + 1] * len + m->act_bias[3 * a + 2] * vel;
    d->act_clamped[a] = f <= m->act_forcerange[2 * a] || f >= m->act_forcerange[2 * a + 1];
    f = clampd(f, m->act_forcerange[2 * a], m->act_forcerange[2 * a + 1]);
    d->qfrc_actuator[k] += g * f;
  }
  double L[SO_NVMAX * SO_NVMAX];
  memcpy(L, d->M, sizeof(double) * nv * nv);
  cholesky(L, nv);
  for (int k = 0; k < nv; k++) { d->qfrc_smooth[k] = d->qfrc_actuator[k] - d->qfrc_bias[k]; d->qacc_smooth[k] = d->qfrc_smooth[k]; }
  chol_solve(L, nv, d->qacc_smooth);
  if (d->nefc > 0) solve_newton(m, d);
  else { memcpy(d->qacc, d->qacc_smooth, sizeof(double) * nv); d->solver_iter = 0; }
  memcpy(d->qacc_warmstart, d->qacc, sizeof(double) * nv);
  /* [upstream] mj_checkAcc: non-finite or huge acceleration marks the env as diverged */
  for (int k = 0; k < nv; k++) if (!isfinite(d->qacc[k]) || fabs(d->qacc[k]) > 1e10) d->diverged = 1;
  integrate(m, d);
}

double so_control_step(const so_model *m, so_data *d, const double *action, const double *offsets, int nsub) {
  for (int a = 0; a < m->nu; a++) d->ctrl[a] = action[a] + (offsets ? offsets[a] : 0); /* so100_task.py:266-287 */
  for (int s = 0; s < nsub; s++) so_substep(m, d);
  so_forward_position(m, d); /* mj_step1 refresh: reward/observations read post-step poses */
  return so_reward(m, d);
}

/* ------------------------------------------------------------------------------------------ reward */
/* oobb_utils.py:202-248 — project the 8+8 corners on the 3 axes of each box (6 axes only), strict comparisons */
static int overlap_aabb_oobb(const double *h0, const double *p1, const double *q1, const double *h1) {
  double va[8][3], vb[8][3], R[9];
  quat2mat(R, q1);
  for (int i = 0; i < 8; i++) {
    int iz = i / 4, ixy = i % 4;
    double t[3] = {(double)(ixy % 2), (double)(ixy / 2), (double)iz};
    double l[3];
    for (int c = 0; c < 3; c++) { va[i][c] = -h0[c] * (1.0 - t[c]) + h0[c] * t[c]; l[c] = -h1[c] * (1.0 - t[c]) + h1[c] * t[c]; }
    mulmv3(vb[i], R, l);
    for (int c = 0; c < 3; c++) vb[i][c] += p1[c];
  }
  for (int a = 0; a < 6; a++) {
    double ax[3];
    if (a < 3) { ax[0] = a == 0; ax[1] = a == 1; ax[2] = a == 2; }
    else { double e[3] = {a == 3, a == 4, a == 5}; mulmv3(ax, R, e); }
    double amax = -INFINITY, amin = INFINITY, bmax = -INFINITY, bmin = INFINITY;
    for (int i = 0; i < 8; i++) {
      double pa = dot3(va[i], ax), pb = dot3(vb[i], ax);
      if (pa > amax) amax = pa; if (pa < amin) amin = pa;
      if (pb > bmax) bmax = pb; if (pb < bmin) bmin = pb;
    }
    if (amax < bmin || amin > bmax) return 0;
  }
  return 1;
}
int so_overlap_oobb_oobb(const double *p0, const double *q0, const double *h0, const double *p1, const double *q1, const double *h1) {
  /* oobb_utils.py:251-273 */
  double inv[4] = {q0[0], -q0[1], -q0[2], -q0[3]}, dp[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, rp[3], rq[4];
  quat_rot(rp, dp, inv);
  quat_mul(rq, inv, q1);
  return overlap_aabb_oobb(h0, rp, rq, h1);
}

double so_reward(const so_model *m, const so_data *d) {
  if (m->nprop < 2) return 0.0; /* SO100Task.get_reward: so100_task.py:289-290 */
  int ob = m->prop_body[0], cb = m->prop_body[1];
  /* success_detector_utils.py:22-28 — linear free-joint velocity only */
  for (int p = 0; p < 2; p++) {
    int b = m->prop_body[p], dof = m->jnt_dofadr[m->body_jntadr[b]];
    double mx = 0;
    for (int c = 0; c < 3; c++) { double a = fabs(d->qvel[dof + c]); if (a > mx) mx = a; }
    if (mx >= 1e-3) return 0.0;
  }
  /* oobb_utils.py:137-148,165-172 — root BVH box at xipos/ximat */
  double q0[4], p0[3], t[3];
  mat2quat(q0, d->ximat[ob]);
  quat_rot(t, m->reward_obj_box, q0);
  for (int c = 0; c < 3; c++) p0[c] = t[c] + d->xipos[ob][c];
  /* oobb_utils.py:175-199 — container box into world */
  double q1[4], p1[3], ident[4] = {1, 0, 0, 0};
  quat_mul(q1, d->xquat[cb], ident);
  for (int k = 0; k < m->nreward_box; k++) { /* so100_hand_over.py:263-273: every overlap box must be touched */
    quat_rot(t, m->reward_box_pos + 3 * k, d->xquat[cb]);
    for (int c = 0; c < 3; c++) p1[c] = t[c] + d->xpos[cb][c];
    if (!so_overlap_oobb_oobb(p0, q0, m->reward_obj_box + 3, p1, q1, m->reward_box_half + 3 * k)) return 0.0;
  }
  return 1.0;
}

/* ------------------------------------------------------------------------------------------ ctypes access */
double *so_field(so_data *d, const char *name, int *count) {
#define FLD(n, cnt) if (strcmp(name, #n) == 0) { if (count) *count = (int)(cnt); return (double *)d->n; }
  FLD(qpos, SO_NQMAX) FLD(qvel, SO_NVMAX) FLD(ctrl, SO_NUMAX) FLD(qacc, SO_NVMAX) FLD(qacc_warmstart, SO_NVMAX)
  FLD(xpos, SO_NBMAX * 3) FLD(xquat, SO_NBMAX * 4) FLD(xmat, SO_NBMAX * 9) FLD(xipos, SO_NBMAX * 3) FLD(ximat, SO_NBMAX * 9)
  FLD(M, SO_NVMAX * SO_NVMAX) FLD(qfrc_bias, SO_NVMAX) FLD(qfrc_actuator, SO_NVMAX) FLD(qacc_smooth, SO_NVMAX) FLD(qfrc_constraint, SO_NVMAX)
  FLD(efc_J, SO_NEFCMAX * SO_NVMAX) FLD(efc_aref, SO_NEFCMAX) FLD(efc_R, SO_NEFCMAX) FLD(efc_force, SO_NEFCMAX) FLD(efc_pos, SO_NEFCMAX)
#undef FLD
  if (strcmp(name, "time") == 0) { if (count) *count = 1; return &d->time; }
  return NULL;
}
int so_info(const so_data *d, const char *name) {
  if (!strcmp(name, "ncon")) return d->ncon;
  if (!strcmp(name, "nefc")) return d->nefc;
  if (!strcmp(name, "solver_iter")) return d->solver_iter;
  if (!strcmp(name, "diverged")) return d->diverged;
  if (!strcmp(name, "ne_fric")) return d->ne_fric;
  if (!strcmp(name, "ne_limit")) return d->ne_limit;
  if (!strcmp(name, "ncon_overflow")) return d->ncon_overflow;
  return -1;
}
void so_set_collide(so_data *d, int enabled) { d->collide_enabled = enabled; }
void so_set_integrator(so_data *d, int implicitfast) { d->integrator = implicitfast ? 1 : 0; }
/* contact c -> out[0..]: dist, pos3, frame9, dim, geom1, geom2, mu, friction5, solref2, solimp5, efc_address (30 values) */
void so_get_contact(const so_data *d, int c, double *out) {
  const so_contact *k = d->contact + c;
  int o = 0;
  out[o++] = k->dist;
  for (int i = 0; i < 3; i++) out[o++] = k->pos[i];
  for (int i = 0; i < 9; i++) out[o++] = k->frame[i];
  out[o++] = k->dim; out[o++] = k->geom1; out[o++] = k->geom2; out[o++] = k->mu;
  for (int i = 0; i < 5; i++) out[o++] = k->friction[i];
  for (int i = 0; i < 2; i++) out[o++] = k->solref[i];
  for (int i = 0; i < 5; i++) out[o++] = k->solimp[i];
  out[o++] = k->efc_address;
}
