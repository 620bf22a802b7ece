/* so101 CPU oracle — TEST INFRASTRUCTURE ONLY.
 *
 * A float64, single-threaded, plain-C restatement of the computation the reference delegates to MuJoCo's
 * mj_step (through dm_control) for the SO100 scene, plus the task's reward.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product (so101_sim_b200/) never does.
 *
 * Parity pin status: the ARM dynamics are pinned by KAT-1 (reference so101_rl.ipynb:219-229, reproduced to <=1e-9
 * relative, see tests/test_oracle_kat.py).  The contact pipeline is "parity unpinned": MuJoCo is absent from the build
 * container, so collision/contact follows MuJoCo's documented pipeline without a numeric pin.
 */
#ifndef SO101_ORACLE_H
#define SO101_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SO_NQMAX 32
#define SO_NVMAX 24
#define SO_NBMAX 24
#define SO_NUMAX 16
#define SO_NCONMAX 256
#define SO_NEFCMAX (2 * SO_NVMAX + 6 * SO_NCONMAX)

enum { SO_GEOM_PLANE = 0, SO_GEOM_SPHERE = 1, SO_GEOM_CAPSULE = 2, SO_GEOM_CYLINDER = 3, SO_GEOM_BOX = 4, SO_GEOM_HULL = 5 };
enum { SO_JNT_FREE = 0, SO_JNT_HINGE = 1 };
enum { SO_ROW_FRICTION = 0, SO_ROW_LIMIT = 1, SO_ROW_CONTACT = 2 /* first row of an elliptic contact */, SO_ROW_CONTACT_FR = 3 };

typedef struct so_model {
  int nq, nv, nu, nbody, njnt, ngeom, nprop, npair, nvert;
  double timestep, gravity[3], impratio, tolerance, meaninertia;
  int iterations, elliptic;
  const int *body_parent, *body_weld, *body_geomadr, *body_geomnum, *body_jntadr, *body_jntnum;
  const double *body_pos, *body_quat, *body_ipos, *body_iquat, *body_mass, *body_inertia, *body_invweight0, *body_bcenter, *body_rbound;
  const int *jnt_type, *jnt_body, *jnt_qposadr, *jnt_dofadr, *jnt_limited;
  const double *jnt_pos, *jnt_axis, *jnt_range, *jnt_solreflimit, *jnt_solimplimit, *jnt_solreffriction, *jnt_solimpfriction, *jnt_margin;
  const int *dof_body, *dof_jnt;
  const double *dof_armature, *dof_frictionloss, *dof_damping, *dof_invweight0, *qpos0;
  const int *act_jnt;
  const double *act_gain, *act_bias, *act_ctrlrange, *act_forcerange, *act_gear;
  const int *geom_type, *geom_body, *geom_condim, *geom_priority, *geom_vertadr, *geom_vertnum, *geom_faceadr, *geom_facenum;
  const double *geom_pos, *geom_mat, *geom_size, *geom_bcenter, *geom_rbound, *geom_friction, *geom_solref, *geom_solimp, *geom_solmix,
      *geom_margin, *geom_gap;
  const double *hull_vert;
  const int *hull_face, *hull_nbradr, *hull_nbr;
  const int *bodypair, *prop_body;
  const double *reward_obj_box, *reward_box_pos, *reward_box_half;  /* box k: pos + 3k, half + 3k */
  int nreward_box;
  void *blob_copy;
} so_model;

typedef struct so_contact {
  double dist, pos[3], frame[9]; /* frame rows: normal (geom1 -> geom2), tangent1, tangent2 */
  double includemargin, friction[5], solref[2], solimp[5], mu;
  int dim, geom1, geom2, efc_address;
} so_contact;

typedef struct so_data {
  double time;
  double qpos[SO_NQMAX], qvel[SO_NVMAX], ctrl[SO_NUMAX], qacc[SO_NVMAX], qacc_warmstart[SO_NVMAX];
  /* position stage */
  double xpos[SO_NBMAX][3], xquat[SO_NBMAX][4], xmat[SO_NBMAX][9], xipos[SO_NBMAX][3], ximat[SO_NBMAX][9];
  double dof_axis[SO_NVMAX][3], dof_anchor[SO_NVMAX][3]; /* world motion axes; translational dofs: axis = direction */
  int dof_trans[SO_NVMAX];
  double M[SO_NVMAX * SO_NVMAX];
  double qfrc_bias[SO_NVMAX], qfrc_actuator[SO_NVMAX], qfrc_smooth[SO_NVMAX], qacc_smooth[SO_NVMAX], qfrc_constraint[SO_NVMAX];
  int ncon, nefc, ne_fric, ne_limit;
  so_contact contact[SO_NCONMAX];
  int efc_type[SO_NEFCMAX], efc_id[SO_NEFCMAX];
  double efc_J[SO_NEFCMAX * SO_NVMAX], efc_pos[SO_NEFCMAX], efc_margin[SO_NEFCMAX], efc_R[SO_NEFCMAX], efc_D[SO_NEFCMAX],
      efc_aref[SO_NEFCMAX], efc_vel[SO_NEFCMAX], efc_frictionloss[SO_NEFCMAX], efc_force[SO_NEFCMAX], efc_diagApprox[SO_NEFCMAX];
  int solver_iter, collide_enabled, diverged, ncon_overflow;
  int integrator;                 /* 0 = semi-implicit Euler (the reference's default), 1 = implicitfast */
  int act_clamped[SO_NUMAX];      /* actuator force sits on its forcerange limit (implicitfast skips its velocity derivative) */
  double solver_cost;
  /* collision statistics */
  long n_narrow, n_gjk_iter, n_epa_iter;
} so_data;

/* model / data life cycle */
so_model *so_model_load(const void *blob, size_t len);
void so_model_free(so_model *m);
so_data *so_data_new(const so_model *m);
void so_data_free(so_data *d);
void so_reset(const so_model *m, so_data *d);

/* physics: [upstream] mj_fwdPosition + mj_fwdVelocity ; one full mj_step */
void so_forward_position(const so_model *m, so_data *d);
void so_substep(const so_model *m, so_data *d);
/* one control step: ctrl = action + offsets (so100_task.py:266-287), nsub substeps, position refresh; returns reward */
double so_control_step(const so_model *m, so_data *d, const double *action, const double *offsets, int nsub);
/* collision (so101_collide.c) */
void so_collide(const so_model *m, so_data *d);
/* Jacobian of a world point attached to a body: jacp, jacr are 3 x nv row-major */
void so_jac(const so_model *m, const so_data *d, int body, const double point[3], double *jacp, double *jacr);

/* task: reward of SO100HandOver in overlap mode (so100_hand_over.py:238-275) */
double so_reward(const so_model *m, const so_data *d);
/* restatement of oobb_utils.overlap_oobb_oobb (oobb_utils.py:251-273): pos3, quat4(wxyz), half3 each */
void so_set_integrator(so_data *d, int implicitfast);
int so_overlap_oobb_oobb(const double *p0, const double *q0, const double *h0, const double *p1, const double *q1, const double *h1);

#ifdef __cplusplus
}
#endif
#endif
