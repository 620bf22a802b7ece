/* so101 CPU oracle — collision detection (float64, plain C).  TEST INFRASTRUCTURE ONLY.
 *
 * Restates [upstream] mj_collision for the SO100 scene (scene_pbr.xml:69-146 geoms + the YCB prop hulls,
 * so100_hand_over.py:159-199): body-pair broad phase with the static filters resolved at model-compile time
 * (tools/compile_model.py), bounding-volume mid phase, and a convex narrow phase.
 * MuJoCo >= 3.3 uses its native GJK/EPA pipeline for every mesh / cylinder pair and, with multiccd enabled
 * (so100_task.py:151), adds manifold points for flat contacts by clipping the two supporting features.  This file
 * follows that structure: boolean GJK -> EPA (penetration depth, normal geom1 -> geom2) -> support-feature clipping.
 * Box-box / capsule-box pairs go through the same generic convex routine instead of MuJoCo's analytic special cases.
 * PARITY UNPINNED: no MuJoCo build is available offline, so contact points/normals are not pinned to upstream numbers;
 * what IS pinned are geometric invariants (tests/test_oracle_collision.py) and resting heights close to KAT-1/KAT-2.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "so101_oracle.h"

#define MINVAL 1e-15
#define MAXFEAT 16  /* vertices kept per supporting feature */
#define MAXCAND 32  /* slab candidates of a primitive / small hull feature (the exact-hull path) */
#define MAXMANIFOLD 4

typedef struct {
  int type, geom;
  double pos[3], mat[9]; /* world pose of the geom frame (hulls: the body frame) */
  double size[3];
  const double *vert; int nvert;
  const int *nbradr, *nbr;  /* hull vertex adjacency (CSR over this geom's vertices, local neighbour ids) */
  int *hint;                /* hill-climbing warm start: last support vertex of this shape while its pair is processed */
  int hint_store;
  double center[3], rbound; /* world bounding sphere */
} shape;

/* Hull support: exhaustive scan for small hulls, otherwise steepest-ascent hill climbing on the hull's vertex graph,
   warm-started from the previous support vertex of the same shape ([upstream] MuJoCo's mesh support does the same for
   meshes with a vertex graph).  On a convex hull a graph-local maximum is the global one. */
#define HILLCLIMB_MIN 40

static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(double *r, const double *a, const double *b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void sub3(double *r, const double *a, const double *b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static inline void mulmv(double *r, const double *m, const double *v) {
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2], z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void mulmtv(double *r, const double *m, const double *v) {
  double x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2], z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void local2world(const shape *s, const double *l, double *w) { mulmv(w, s->mat, l); w[0] += s->pos[0]; w[1] += s->pos[1]; w[2] += s->pos[2]; }

static void make_shape(const so_model *m, const so_data *d, int g, shape *s) {
  int b = m->geom_body[g];
  s->type = m->geom_type[g]; s->geom = g;
  double t[3];
  mulmv(t, d->xmat[b], m->geom_pos + 3 * g);
  for (int c = 0; c < 3; c++) s->pos[c] = d->xpos[b][c] + t[c];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double v = 0;
    for (int k = 0; k < 3; k++) v += d->xmat[b][3 * i + k] * m->geom_mat[9 * g + 3 * k + j];
    s->mat[3 * i + j] = v;
  }
  memcpy(s->size, m->geom_size + 3 * g, sizeof s->size);
  s->vert = m->hull_vert + 3 * m->geom_vertadr[g]; s->nvert = m->geom_vertnum[g];
  s->nbradr = m->hull_nbradr + m->geom_vertadr[g]; s->nbr = m->hull_nbr;
  s->hint_store = 0; s->hint = &s->hint_store;
  mulmv(t, d->xmat[b], m->geom_bcenter + 3 * g);
  for (int c = 0; c < 3; c++) s->center[c] = d->xpos[b][c] + t[c];
  s->rbound = m->geom_rbound[g];
}

/* support point of a convex shape in world direction dir (not necessarily unit) */
static void support(const shape *s, const double *dir, double *out) {
  double dl[3], p[3];
  mulmtv(dl, s->mat, dir);
  switch (s->type) {
    case SO_GEOM_HULL: {
      int best = 0; double bv = -INFINITY;
      if (s->nvert < HILLCLIMB_MIN) {
        for (int i = 0; i < s->nvert; i++) { double v = dot3(s->vert + 3 * i, dl); if (v > bv) { bv = v; best = i; } }
      } else {
        int cur = *s->hint;
        bv = dot3(s->vert + 3 * cur, dl);
        for (;;) {
          int nxt = cur;
          for (int k = s->nbradr[cur]; k < s->nbradr[cur + 1]; k++) {
            int j = s->nbr[k];
            double v = dot3(s->vert + 3 * j, dl);
            if (v > bv) { bv = v; nxt = j; }
          }
          if (nxt == cur) break;
          cur = nxt;
        }
        best = cur;
        *s->hint = cur;
      }
      memcpy(p, s->vert + 3 * best, sizeof p);
      break;
    }
    case SO_GEOM_BOX:
      for (int c = 0; c < 3; c++) p[c] = dl[c] >= 0 ? s->size[c] : -s->size[c];
      break;
    case SO_GEOM_CYLINDER: {
      double n = sqrt(dl[0] * dl[0] + dl[1] * dl[1]);
      p[0] = n > MINVAL ? s->size[0] * dl[0] / n : 0; p[1] = n > MINVAL ? s->size[0] * dl[1] / n : 0;
      p[2] = dl[2] >= 0 ? s->size[1] : -s->size[1];
      break;
    }
    case SO_GEOM_CAPSULE: {
      double n = sqrt(dot3(dl, dl));
      for (int c = 0; c < 3; c++) p[c] = n > MINVAL ? s->size[0] * dl[c] / n : 0;
      p[2] += dl[2] >= 0 ? s->size[1] : -s->size[1];
      break;
    }
    case SO_GEOM_SPHERE: {
      double n = sqrt(dot3(dl, dl));
      for (int c = 0; c < 3; c++) p[c] = n > MINVAL ? s->size[0] * dl[c] / n : 0;
      break;
    }
    default: p[0] = p[1] = p[2] = 0;
  }
  local2world(s, p, out);
}

/* ------------------------------------------------------------------------------------------ GJK + EPA */
typedef struct { double w[3], a[3], b[3]; } mpoint; /* Minkowski difference point w = a - b with its witnesses */

static void msupport(const shape *A, const shape *B, const double *d, mpoint *p) {
  double nd[3] = {-d[0], -d[1], -d[2]};
  support(A, d, p->a); support(B, nd, p->b);
  sub3(p->w, p->a, p->b);
}

/* boolean GJK; on success the simplex (n points, newest last) encloses or touches the origin */
static int gjk_intersect(const shape *A, const shape *B, mpoint *S, int *np, long *iters) {
  double d[3];
  sub3(d, B->center, A->center);
  if (dot3(d, d) < 1e-20) { d[0] = 1; d[1] = 0; d[2] = 0; }
  int n = 0;
  for (int it = 0; it < 64; it++) {
    (*iters)++;
    mpoint p;
    msupport(A, B, d, &p);
    if (dot3(p.w, d) < 0) { *np = n; return 0; }
    S[n++] = p;
    /* nearest-simplex update */
    if (n == 1) { d[0] = -S[0].w[0]; d[1] = -S[0].w[1]; d[2] = -S[0].w[2]; }
    else if (n == 2) {
      double ab[3], ao[3] = {-S[1].w[0], -S[1].w[1], -S[1].w[2]}, t[3];
      sub3(ab, S[0].w, S[1].w);
      if (dot3(ab, ao) > 0) { cross3(t, ab, ao); cross3(d, t, ab); }
      else { S[0] = S[1]; n = 1; memcpy(d, ao, sizeof ao); }
    } else if (n == 3) {
      /* a = S[2] newest, b = S[1], c = S[0] */
      double *a = S[2].w, *b = S[1].w, *c = S[0].w, ab[3], ac[3], ao[3] = {-a[0], -a[1], -a[2]}, abc[3], t[3];
      sub3(ab, b, a); sub3(ac, c, a); cross3(abc, ab, ac);
      cross3(t, abc, ac);
      if (dot3(t, ao) > 0) {
        if (dot3(ac, ao) > 0) { S[1] = S[2]; n = 2; double u[3]; cross3(u, ac, ao); cross3(d, u, ac); } /* keep c, a */
        else goto star3;
      } else {
        cross3(t, ab, abc);
        if (dot3(t, ao) > 0) {
        star3:
          if (dot3(ab, ao) > 0) { S[0] = S[1]; S[1] = S[2]; n = 2; double u[3]; cross3(u, ab, ao); cross3(d, u, ab); } /* keep b, a */
          else { S[0] = S[2]; n = 1; memcpy(d, ao, sizeof ao); }
        } else {
          if (dot3(abc, ao) > 0) memcpy(d, abc, sizeof abc);
          else { mpoint tmp = S[0]; S[0] = S[1]; S[1] = tmp; d[0] = -abc[0]; d[1] = -abc[1]; d[2] = -abc[2]; }
        }
      }
    } else {
      /* tetrahedron: a = S[3] newest, b = S[2], c = S[1], dd = S[0] */
      double *a = S[3].w, *b = S[2].w, *c = S[1].w, *e = S[0].w, ab[3], ac[3], ad[3], ao[3] = {-a[0], -a[1], -a[2]}, abc[3], acd[3], adb[3];
      sub3(ab, b, a); sub3(ac, c, a); sub3(ad, e, a);
      cross3(abc, ab, ac); cross3(acd, ac, ad); cross3(adb, ad, ab);
      /* orient face normals outward (away from the opposite vertex) */
      if (dot3(abc, ad) > 0) { abc[0] = -abc[0]; abc[1] = -abc[1]; abc[2] = -abc[2]; }
      if (dot3(acd, ab) > 0) { acd[0] = -acd[0]; acd[1] = -acd[1]; acd[2] = -acd[2]; }
      if (dot3(adb, ac) > 0) { adb[0] = -adb[0]; adb[1] = -adb[1]; adb[2] = -adb[2]; }
      double da = dot3(abc, ao), db = dot3(acd, ao), dc = dot3(adb, ao);
      if (da > 0 && da >= db && da >= dc) { S[0] = S[1]; S[1] = S[2]; S[2] = S[3]; n = 3; memcpy(d, abc, sizeof abc); }       /* c b a */
      else if (db > 0 && db >= dc) { /* keep d c a */ S[2] = S[3]; n = 3; memcpy(d, acd, sizeof acd); }
      else if (dc > 0) { /* keep d b a */ S[1] = S[2]; S[2] = S[3]; n = 3; memcpy(d, adb, sizeof adb); }
      else { *np = 4; return 1; }
    }
    if (dot3(d, d) < 1e-30) { *np = n; return 1; } /* origin lies on the current simplex */
  }
  *np = n;
  return n == 4;
}

typedef struct { int v[3]; double n[3], d; int alive; } eface;
#define EPA_MAXV 96
#define EPA_MAXF 256
#define EPA_MAXIT 50 /* [upstream] mjOption.ccd_iterations default */

static int epa_add_face(eface *F, int *nf, const mpoint *V, int a, int b, int c, const double *inside) {
  if (*nf >= EPA_MAXF) return -1;
  eface *f = F + *nf;
  double ab[3], ac[3];
  sub3(ab, V[b].w, V[a].w); sub3(ac, V[c].w, V[a].w);
  cross3(f->n, ab, ac);
  double l = sqrt(dot3(f->n, f->n));
  if (l < 1e-30) return -1;
  for (int k = 0; k < 3; k++) f->n[k] /= l;
  double t[3];
  sub3(t, V[a].w, inside);
  if (dot3(f->n, t) < 0) { f->n[0] = -f->n[0]; f->n[1] = -f->n[1]; f->n[2] = -f->n[2]; f->v[0] = a; f->v[1] = c; f->v[2] = b; }
  else { f->v[0] = a; f->v[1] = b; f->v[2] = c; }
  f->d = dot3(f->n, V[a].w);
  f->alive = 1;
  return (*nf)++;
}

/* EPA: returns 1 and fills normal (A -> B), depth >= 0 and witness points; 0 when the penetration is degenerate */
static int epa(const shape *A, const shape *B, mpoint *S, int n, double *normal, double *depth, double *pa, double *pb, long *iters) {
  mpoint V[EPA_MAXV];
  eface F[EPA_MAXF];
  int nv = 0, nf = 0;
  for (int i = 0; i < n; i++) V[nv++] = S[i];
  /* blow the simplex up to a tetrahedron */
  if (nv == 1) return 0;
  if (nv == 2) {
    double ab[3], ax[3] = {0, 0, 0}, d[3];
    sub3(ab, V[1].w, V[0].w);
    int k = fabs(ab[0]) < fabs(ab[1]) ? (fabs(ab[0]) < fabs(ab[2]) ? 0 : 2) : (fabs(ab[1]) < fabs(ab[2]) ? 1 : 2);
    ax[k] = 1; cross3(d, ab, ax);
    msupport(A, B, d, &V[nv]);
    double t[3], cr[3]; sub3(t, V[nv].w, V[0].w); cross3(cr, ab, t);
    if (dot3(cr, cr) < 1e-24) { d[0] = -d[0]; d[1] = -d[1]; d[2] = -d[2]; msupport(A, B, d, &V[nv]); }
    nv++;
  }
  if (nv == 3) {
    double ab[3], ac[3], nn[3], t[3];
    sub3(ab, V[1].w, V[0].w); sub3(ac, V[2].w, V[0].w); cross3(nn, ab, ac);
    if (dot3(nn, nn) < 1e-30) return 0;
    msupport(A, B, nn, &V[3]);
    sub3(t, V[3].w, V[0].w);
    if (fabs(dot3(t, nn)) < 1e-12 * sqrt(dot3(nn, nn))) {
      double m[3] = {-nn[0], -nn[1], -nn[2]};
      msupport(A, B, m, &V[3]);
      sub3(t, V[3].w, V[0].w);
      if (fabs(dot3(t, nn)) < 1e-12 * sqrt(dot3(nn, nn))) return 0;
    }
    nv = 4;
  }
  double inside[3] = {0, 0, 0};
  for (int i = 0; i < 4; i++) for (int k = 0; k < 3; k++) inside[k] += 0.25 * V[i].w[k];
  if (epa_add_face(F, &nf, V, 0, 1, 2, inside) < 0 || epa_add_face(F, &nf, V, 0, 1, 3, inside) < 0 ||
      epa_add_face(F, &nf, V, 0, 2, 3, inside) < 0 || epa_add_face(F, &nf, V, 1, 2, 3, inside) < 0) return 0;
  int best = -1;
  for (int it = 0; it < EPA_MAXIT; it++) {
    (*iters)++;
    best = -1;
    double bd = INFINITY;
    for (int f = 0; f < nf; f++) if (F[f].alive && F[f].d < bd) { bd = F[f].d; best = f; }
    if (best < 0) return 0;
    if (nv >= EPA_MAXV) break;
    mpoint p;
    msupport(A, B, F[best].n, &p);
    double adv = dot3(p.w, F[best].n) - F[best].d;
    if (adv < 1e-9) break;
    /* remove faces visible from p, collect horizon */
    int he[EPA_MAXF][2], nh = 0;
    V[nv] = p;
    for (int f = 0; f < nf; f++) {
      if (!F[f].alive) continue;
      if (dot3(F[f].n, p.w) - F[f].d > 1e-12) { /* p above the face plane (n . v0 = d) */
        F[f].alive = 0;
        for (int e = 0; e < 3; e++) {
          int a = F[f].v[e], b = F[f].v[(e + 1) % 3], found = 0;
          for (int h = 0; h < nh; h++) if (he[h][0] == b && he[h][1] == a) { he[h][0] = he[nh - 1][0]; he[h][1] = he[nh - 1][1]; nh--; found = 1; break; }
          if (!found && nh < EPA_MAXF) { he[nh][0] = a; he[nh][1] = b; nh++; }
        }
      }
    }
    if (nh == 0) break;
    int failed = 0;
    for (int h = 0; h < nh; h++) if (epa_add_face(F, &nf, V, he[h][0], he[h][1], nv, inside) < 0) { failed = 1; }
    nv++;
    if (failed) break;
  }
  if (best < 0 || !F[best].alive) {
    best = -1; double bd = INFINITY;
    for (int f = 0; f < nf; f++) if (F[f].alive && F[f].d < bd) { bd = F[f].d; best = f; }
    if (best < 0) return 0;
  }
  eface *f = F + best;
  memcpy(normal, f->n, sizeof f->n);
  *depth = f->d > 0 ? f->d : 0;
  /* barycentric coordinates of the origin's projection on the closest face */
  double p[3] = {f->n[0] * f->d, f->n[1] * f->d, f->n[2] * f->d};
  const double *a = V[f->v[0]].w, *b = V[f->v[1]].w, *c = V[f->v[2]].w;
  double v0[3], v1[3], v2[3];
  sub3(v0, b, a); sub3(v1, c, a); sub3(v2, p, a);
  double d00 = dot3(v0, v0), d01 = dot3(v0, v1), d11 = dot3(v1, v1), d20 = dot3(v2, v0), d21 = dot3(v2, v1), den = d00 * d11 - d01 * d01;
  double bv = 1.0 / 3, bw = 1.0 / 3;
  if (fabs(den) > 1e-30) { bv = (d11 * d20 - d01 * d21) / den; bw = (d00 * d21 - d01 * d20) / den; }
  double bu = 1 - bv - bw;
  for (int k = 0; k < 3; k++) {
    pa[k] = bu * V[f->v[0]].a[k] + bv * V[f->v[1]].a[k] + bw * V[f->v[2]].a[k];
    pb[k] = bu * V[f->v[0]].b[k] + bv * V[f->v[1]].b[k] + bw * V[f->v[2]].b[k];
  }
  return 1;
}

/* ------------------------------------------------------------------------------------------ support features */
typedef struct { double x, y, h; } fpt; /* tangent-plane coordinates and height along the direction */

/* Large hull features (more than FEAT_EXACT slab vertices: rims and finely tessellated patches) are represented by their
   extreme points along FEAT_DIRS tangent-plane directions, counter-clockwise from -x (where the monotone chain of the
   exact path starts as well): an inscribed convex polygon found in the same pass that counts the slab, with no storage,
   sort or hull construction.  Small features (faces, edges, vertices) keep the exact 2-D hull. */
#define FEAT_EXACT 16
#define FEAT_DIRS 16
static const double FEAT_COS[8] = {1.0, 0.92387953251128674, 0.70710678118654752, 0.38268343236508977, 0.0,
                                   -0.38268343236508977, -0.70710678118654752, -0.92387953251128674};
static const double FEAT_SIN[8] = {0.0, 0.38268343236508977, 0.70710678118654752, 0.92387953251128674, 1.0,
                                   0.92387953251128674, 0.70710678118654752, 0.38268343236508977};

/* vertices of `s` whose height along dir (unit) is within delta of the maximum -> 2-D convex polygon (CCW) with heights */
static int feature(const shape *s, const double *dir, const double *t1, const double *t2, double delta, fpt *out) {
  double cand[MAXCAND][3];
  fpt P[MAXCAND];
  int nc = 0, projected = 0;
  double sp[3];
  support(s, dir, sp);
  double hmax = dot3(sp, dir);
  double w[3];
  switch (s->type) {
    case SO_GEOM_HULL: {
      /* tangent-plane coordinates are evaluated in the hull frame: x = v . (R^T t1) + pos . t1 */
      double dl[3], t1l[3], t2l[3];
      mulmtv(dl, s->mat, dir); mulmtv(t1l, s->mat, t1); mulmtv(t2l, s->mat, t2);
      double off = dot3(s->pos, dir), ox = dot3(s->pos, t1), oy = dot3(s->pos, t2);
      double emax[8], emin[8]; int imax[8], imin[8], nband = 0;
      for (int k = 0; k < 8; k++) { emax[k] = -INFINITY; emin[k] = INFINITY; imax[k] = imin[k] = 0; }
      for (int i = 0; i < s->nvert; i++) {
        const double *v = s->vert + 3 * i;
        if (!(dot3(v, dl) + off >= hmax - delta)) continue;
        nband++;
        double x = dot3(v, t1l) + ox, y = dot3(v, t2l) + oy;
        for (int k = 0; k < 8; k++) {
          double val = FEAT_COS[k] * x + FEAT_SIN[k] * y;
          if (val > emax[k]) { emax[k] = val; imax[k] = i; }
          if (val < emin[k]) { emin[k] = val; imin[k] = i; }
        }
      }
      if (nband > FEAT_EXACT) {
        /* direction j (angle pi + j * 2 pi / 16): j < 8 -> minimum along direction j, else maximum along direction j - 8 */
        int kept[FEAT_DIRS], nk = 0;
        for (int j = 0; j < FEAT_DIRS; j++) {
          int idx = j < 8 ? imin[j] : imax[j - 8], dup = 0;
          for (int q = 0; q < nk; q++) if (kept[q] == idx) dup = 1;
          if (!dup) kept[nk++] = idx;
        }
        for (int q = 0; q < nk; q++) {
          const double *v = s->vert + 3 * kept[q];
          out[q].x = dot3(v, t1l) + ox; out[q].y = dot3(v, t2l) + oy; out[q].h = dot3(v, dl) + off;
        }
        return nk;
      }
      for (int i = 0; i < s->nvert && nc < MAXCAND; i++) {
        const double *v = s->vert + 3 * i;
        double h = dot3(v, dl) + off;
        if (!(h >= hmax - delta)) continue;
        P[nc].x = dot3(v, t1l) + ox; P[nc].y = dot3(v, t2l) + oy; P[nc].h = h; nc++;
      }
      if (nc > 0) projected = 1;
      break;
    }
    case SO_GEOM_BOX:
      for (int i = 0; i < 8; i++) {
        double l[3] = {(i & 1 ? 1 : -1) * s->size[0], (i & 2 ? 1 : -1) * s->size[1], (i & 4 ? 1 : -1) * s->size[2]};
        local2world(s, l, w);
        if (dot3(w, dir) >= hmax - delta) { memcpy(cand[nc], w, sizeof w); nc++; }
      }
      break;
    case SO_GEOM_CYLINDER:
      for (int cap = -1; cap <= 1; cap += 2)
        for (int i = 0; i < 16; i++) {
          double ang = 2 * M_PI * i / 16, l[3] = {s->size[0] * cos(ang), s->size[0] * sin(ang), cap * s->size[1]};
          local2world(s, l, w);
          if (dot3(w, dir) >= hmax - delta) { memcpy(cand[nc], w, sizeof w); nc++; }
        }
      if (nc == 0) { memcpy(cand[0], sp, sizeof sp); nc = 1; }
      break;
    case SO_GEOM_CAPSULE:
      for (int e = -1; e <= 1; e += 2) {
        double l[3] = {0, 0, e * s->size[1]};
        local2world(s, l, w);
        for (int c = 0; c < 3; c++) w[c] += s->size[0] * dir[c];
        if (dot3(w, dir) >= hmax - delta) { memcpy(cand[nc], w, sizeof w); nc++; }
      }
      break;
    default: memcpy(cand[0], sp, sizeof sp); nc = 1;
  }
  if (nc == 0) { memcpy(cand[0], sp, sizeof sp); nc = 1; }
  /* project and take the 2-D convex hull (Andrew's monotone chain) */
  if (!projected)
    for (int i = 0; i < nc; i++) { P[i].x = dot3(cand[i], t1); P[i].y = dot3(cand[i], t2); P[i].h = dot3(cand[i], dir); }
  for (int i = 1; i < nc; i++) { /* insertion sort by (x, y) */
    fpt k = P[i]; int j = i - 1;
    while (j >= 0 && (P[j].x > k.x || (P[j].x == k.x && P[j].y > k.y))) { P[j + 1] = P[j]; j--; }
    P[j + 1] = k;
  }
  if (nc <= 2) { for (int i = 0; i < nc; i++) out[i] = P[i]; return nc; }
  fpt H[MAXCAND * 2 + 2];
  int k = 0;
  for (int i = 0; i < nc; i++) {
    while (k >= 2 && (H[k - 1].x - H[k - 2].x) * (P[i].y - H[k - 2].y) - (H[k - 1].y - H[k - 2].y) * (P[i].x - H[k - 2].x) <= 1e-14) k--;
    H[k++] = P[i];
  }
  for (int i = nc - 2, t = k + 1; i >= 0; i--) {
    while (k >= t && (H[k - 1].x - H[k - 2].x) * (P[i].y - H[k - 2].y) - (H[k - 1].y - H[k - 2].y) * (P[i].x - H[k - 2].x) <= 1e-14) k--;
    H[k++] = P[i];
  }
  k--; /* last point equals the first */
  if (k > MAXFEAT) k = MAXFEAT;
  for (int i = 0; i < k; i++) out[i] = H[i];
  return k;
}

/* height of a feature at tangent-plane location (x, y): plane through 3 spread vertices, line for 2, constant for 1 */
static double feature_height(const fpt *P, int n, double x, double y) {
  if (n == 1) return P[0].h;
  int i1 = 1; double best = -1;
  for (int i = 1; i < n; i++) { double dx = P[i].x - P[0].x, dy = P[i].y - P[0].y, l = dx * dx + dy * dy; if (l > best) { best = l; i1 = i; } }
  double ex = P[i1].x - P[0].x, ey = P[i1].y - P[0].y, eh = P[i1].h - P[0].h, el = ex * ex + ey * ey;
  if (n == 2 || el < 1e-20) {
    if (el < 1e-20) return P[0].h;
    double t = ((x - P[0].x) * ex + (y - P[0].y) * ey) / el;
    return P[0].h + t * eh;
  }
  int i2 = -1; best = 0;
  for (int i = 1; i < n; i++) { double a = fabs(ex * (P[i].y - P[0].y) - ey * (P[i].x - P[0].x)); if (a > best) { best = a; i2 = i; } }
  if (i2 < 0 || best < 1e-12 * el) { double t = ((x - P[0].x) * ex + (y - P[0].y) * ey) / el; return P[0].h + t * eh; }
  double fx = P[i2].x - P[0].x, fy = P[i2].y - P[0].y, fh = P[i2].h - P[0].h, det = ex * fy - ey * fx;
  double px = x - P[0].x, py = y - P[0].y, u = (px * fy - py * fx) / det, v = (ex * py - ey * px) / det;
  return P[0].h + u * eh + v * fh;
}

/* clip convex polygon `subj` (CCW, n >= 1) by convex polygon `clip` (CCW, m >= 3): Sutherland-Hodgman */
static int clip_poly(const fpt *subj, int n, const fpt *clip, int m, fpt *out) {
  fpt bufA[MAXFEAT * 2 + 8], bufB[MAXFEAT * 2 + 8];
  int na = n;
  memcpy(bufA, subj, sizeof(fpt) * n);
  fpt *in = bufA, *res = bufB;
  for (int e = 0; e < m && na > 0; e++) {
    double ax = clip[e].x, ay = clip[e].y, bx = clip[(e + 1) % m].x, by = clip[(e + 1) % m].y;
    double ex = bx - ax, ey = by - ay, tol = 1e-12;
    int nr = 0;
    for (int i = 0; i < na; i++) {
      fpt P = in[i], Q = in[(i + 1) % na];
      double sp = ex * (P.y - ay) - ey * (P.x - ax), sq = ex * (Q.y - ay) - ey * (Q.x - ax);
      int pin = sp >= -tol, qin = sq >= -tol;
      if (pin) res[nr++] = P;
      if (na > 1 && pin != qin) {
        double t = sp / (sp - sq);
        fpt I = {P.x + t * (Q.x - P.x), P.y + t * (Q.y - P.y), P.h + t * (Q.h - P.h)};
        res[nr++] = I;
      }
      if (na == 1) break;
      if (na == 2 && i == 0 && 0) break;
    }
    if (na == 2) { /* a segment visits each end once: drop the wrap-around duplicate handling */
      nr = 0;
      fpt P = in[0], Q = in[1];
      double sp = ex * (P.y - ay) - ey * (P.x - ax), sq = ex * (Q.y - ay) - ey * (Q.x - ax);
      int pin = sp >= -tol, qin = sq >= -tol;
      if (pin && qin) { res[nr++] = P; res[nr++] = Q; }
      else if (pin || qin) {
        double t = sp / (sp - sq);
        fpt I = {P.x + t * (Q.x - P.x), P.y + t * (Q.y - P.y), P.h + t * (Q.h - P.h)};
        if (pin) { res[nr++] = P; res[nr++] = I; } else { res[nr++] = I; res[nr++] = Q; }
      }
    }
    fpt *tmp = in; in = res; res = tmp;
    na = nr;
    if (na > MAXFEAT * 2) na = MAXFEAT * 2;
  }
  memcpy(out, in, sizeof(fpt) * na);
  return na;
}

/* ------------------------------------------------------------------------------------------ contact emission */
static void mix_params(const so_model *m, int g1, int g2, so_contact *c) { /* [upstream] mj_contactParam */
  c->dim = m->geom_condim[g1] > m->geom_condim[g2] ? m->geom_condim[g1] : m->geom_condim[g2];
  double f[3];
  int p1 = m->geom_priority[g1], p2 = m->geom_priority[g2];
  double mix;
  if (p1 == p2) {
    for (int k = 0; k < 3; k++) f[k] = fmax(m->geom_friction[3 * g1 + k], m->geom_friction[3 * g2 + k]);
    double s1 = m->geom_solmix[g1], s2 = m->geom_solmix[g2];
    if (s1 >= MINVAL && s2 >= MINVAL) mix = s1 / (s1 + s2);
    else if (s1 < MINVAL && s2 < MINVAL) mix = 0.5;
    else mix = s1 < MINVAL ? 0.0 : 1.0;
  } else {
    int g = p1 > p2 ? g1 : g2;
    for (int k = 0; k < 3; k++) f[k] = m->geom_friction[3 * g + k];
    mix = p1 > p2 ? 1.0 : 0.0;
  }
  const double *r1 = m->geom_solref + 2 * g1, *r2 = m->geom_solref + 2 * g2;
  if (r1[0] > 0 && r2[0] > 0) for (int k = 0; k < 2; k++) c->solref[k] = mix * r1[k] + (1 - mix) * r2[k];
  else for (int k = 0; k < 2; k++) c->solref[k] = fmin(r1[k], r2[k]);
  for (int k = 0; k < 5; k++) c->solimp[k] = mix * m->geom_solimp[5 * g1 + k] + (1 - mix) * m->geom_solimp[5 * g2 + k];
  c->friction[0] = c->friction[1] = f[0]; c->friction[2] = f[1]; c->friction[3] = c->friction[4] = f[2];
  double margin = fmax(m->geom_margin[g1], m->geom_margin[g2]), gap = fmax(m->geom_gap[g1], m->geom_gap[g2]);
  c->includemargin = margin - gap;
}

static void frame_from_normal(const double *n, double *frame) { /* [upstream] mju_makeFrame */
  double *x = frame, *y = frame + 3, *z = frame + 6;
  double l = sqrt(dot3(n, n));
  for (int c = 0; c < 3; c++) x[c] = n[c] / l;
  y[0] = y[1] = y[2] = 0;
  if (x[1] < 0.5 && x[1] > -0.5) y[1] = 1; else y[2] = 1;
  double dd = dot3(x, y);
  for (int c = 0; c < 3; c++) y[c] -= dd * x[c];
  l = sqrt(dot3(y, y));
  for (int c = 0; c < 3; c++) y[c] /= l;
  cross3(z, x, y);
}

static void emit(const so_model *m, so_data *d, int g1, int g2, const double *frame, const double *pos, double dist) {
  if (d->ncon >= SO_NCONMAX) { d->ncon_overflow++; return; }
  so_contact *c = d->contact + d->ncon++;
  memset(c, 0, sizeof *c);
  c->geom1 = g1; c->geom2 = g2; c->dist = dist;
  memcpy(c->pos, pos, sizeof(double) * 3); memcpy(c->frame, frame, sizeof(double) * 9);
  mix_params(m, g1, g2, c);
}

/* keep at most MAXMANIFOLD points: deepest, farthest from it, farthest from that line on each side */
static int reduce_manifold(fpt *P, double *dist, int n) {
  if (n <= MAXMANIFOLD) return n;
  /* Flat contacts make all candidates tie to round-off; every comparison therefore carries a tolerance (0.1 um on depth,
     1e-4 relative on lengths/areas) so that the first candidate in polygon order wins a tie in any arithmetic. */
  int sel[4] = {0, -1, -1, -1};
  for (int i = 1; i < n; i++) if (dist[i] < dist[sel[0]] - 1e-7) sel[0] = i;
  double best = -1;
  for (int i = 0; i < n; i++) { double dx = P[i].x - P[sel[0]].x, dy = P[i].y - P[sel[0]].y, l = dx * dx + dy * dy; if (l > best * 1.0001 + 1e-12) { best = l; sel[1] = i; } }
  double ex = P[sel[1]].x - P[sel[0]].x, ey = P[sel[1]].y - P[sel[0]].y, bp = 0, bn = 0;
  for (int i = 0; i < n; i++) {
    if (i == sel[0] || i == sel[1]) continue; /* their cross product is 0 up to round-off (FMA contraction makes it +-eps) */
    double s = ex * (P[i].y - P[sel[0]].y) - ey * (P[i].x - P[sel[0]].x);
    if (s > bp * 1.0001 + 1e-12) { bp = s; sel[2] = i; }
    if (s < bn * 1.0001 - 1e-12) { bn = s; sel[3] = i; }
  }
  fpt Q[4]; double qd[4]; int k = 0;
  for (int i = 0; i < 4; i++) if (sel[i] >= 0) { Q[k] = P[sel[i]]; qd[k] = dist[sel[i]]; k++; }
  memcpy(P, Q, sizeof(fpt) * k); memcpy(dist, qd, sizeof(double) * k);
  return k;
}

/* manifold from the two supporting features along normal n (A -> B).  Returns number of contacts emitted. */
static int manifold(const so_model *m, so_data *d, const shape *A, const shape *B, const double *n, double depth) {
  double frame[9];
  frame_from_normal(n, frame);
  const double *t1 = frame + 3, *t2 = frame + 6;
  double nn[3] = {-frame[0], -frame[1], -frame[2]};
  double delta = depth + 1e-7;
  fpt FA[MAXFEAT], FB[MAXFEAT], R[MAXFEAT * 2 + 8];
  int na = feature(A, frame, t1, t2, delta, FA);
  int nb = feature(B, nn, t1, t2, delta, FB);
  /* heights of B's feature were measured along -n: convert to heights along +n */
  for (int i = 0; i < nb; i++) FB[i].h = -FB[i].h;
  /* FB is CCW for (t1,t2) seen along -n; hull was built in the same (t1,t2) coordinates, so orientation is already CCW */
  const char *dbg = getenv("SO101_ORACLE_DBG");
  if (dbg && atoi(dbg) == d->solver_iter + 1000000 * 0 && 0) {}
  int dbgon = dbg && A->geom == atoi(dbg) && B->geom == atoi(strchr(dbg, ',') ? strchr(dbg, ',') + 1 : "-1");
  if (dbgon) {
    printf("ORACLE MANIFOLD g=(%d,%d) n=(%.17g,%.17g,%.17g) depth=%.17g na=%d nb=%d\n", A->geom, B->geom, n[0], n[1], n[2], depth, na, nb);
    for (int i = 0; i < na; i++) printf("  FA[%d]=(%.17g,%.17g,%.17g)\n", i, FA[i].x, FA[i].y, FA[i].h);
    for (int i = 0; i < nb; i++) printf("  FB[%d]=(%.17g,%.17g,%.17g)\n", i, FB[i].x, FB[i].y, -FB[i].h);
  }
  int nr = 0;
  if (na >= 3 && nb >= 3) {
    nr = clip_poly(FA, na, FB, nb, R);
  } else if (na >= 3 && nb == 2) {
    nr = clip_poly(FB, nb, FA, na, R);
  } else if (nb >= 3 && na <= 2) {
    nr = clip_poly(FA, na, FB, nb, R);
  } else if (na >= 3 && nb == 1) {
    nr = clip_poly(FB, nb, FA, na, R);
  } else nr = 0;
  double dist[MAXFEAT * 2 + 8];
  int k = 0;
  for (int i = 0; i < nr; i++) {
    double ha = feature_height(FA, na, R[i].x, R[i].y), hb = feature_height(FB, nb, R[i].x, R[i].y);
    double di = hb - ha; /* gap along n between B's and A's surfaces: negative = penetration */
    if (di < 0) { R[k] = R[i]; R[k].h = 0.5 * (ha + hb); dist[k] = di; k++; }
  }
  /* drop near-duplicate points */
  int u = 0;
  for (int i = 0; i < k; i++) {
    int dup = 0;
    for (int j = 0; j < u; j++) if (fabs(R[i].x - R[j].x) + fabs(R[i].y - R[j].y) < 1e-7) { dup = 1; if (dist[i] < dist[j]) { R[j] = R[i]; dist[j] = dist[i]; } break; }
    if (!dup) { R[u] = R[i]; dist[u] = dist[i]; u++; }
  }
  if (dbgon) { printf("  nr=%d k=%d u=%d\n", nr, k, u); for (int i = 0; i < u; i++) printf("  R[%d]=(%.17g,%.17g) d=%.17g\n", i, R[i].x, R[i].y, dist[i]); }
  u = reduce_manifold(R, dist, u);
  for (int i = 0; i < u; i++) {
    double pos[3];
    for (int c = 0; c < 3; c++) pos[c] = R[i].x * t1[c] + R[i].y * t2[c] + R[i].h * frame[c];
    emit(m, d, A->geom, B->geom, frame, pos, dist[i]);
  }
  return u;
}

static void collide_convex(const so_model *m, so_data *d, const shape *A, const shape *B) {
  mpoint S[4];
  int n = 0;
  d->n_narrow++;
  if (!gjk_intersect(A, B, S, &n, &d->n_gjk_iter)) return;
  double normal[3], depth, pa[3], pb[3];
  if (!epa(A, B, S, n, normal, &depth, pa, pb, &d->n_epa_iter)) return;
  if (depth <= 0) return;
  if (manifold(m, d, A, B, normal, depth) > 0) return;
  double frame[9], pos[3];
  frame_from_normal(normal, frame);
  for (int c = 0; c < 3; c++) pos[c] = 0.5 * (pa[c] + pb[c]);
  emit(m, d, A->geom, B->geom, frame, pos, -depth);
}

static void collide_plane(const so_model *m, so_data *d, const shape *P, const shape *B) {
  /* plane normal = local z; normal geom1 (plane) -> geom2 */
  double n[3] = {P->mat[2], P->mat[5], P->mat[8]}, nn[3] = {-n[0], -n[1], -n[2]}, sp[3];
  support(B, nn, sp);
  double off = dot3(n, P->pos), depth = off - dot3(sp, n);
  if (depth <= 0) return;
  d->n_narrow++;
  double frame[9];
  frame_from_normal(n, frame);
  const double *t1 = frame + 3, *t2 = frame + 6;
  fpt FB[MAXFEAT];
  int nb = feature(B, nn, t1, t2, depth + 1e-7, FB);
  double dist[MAXFEAT];
  for (int i = 0; i < nb; i++) { FB[i].h = -FB[i].h; dist[i] = FB[i].h - off; }
  nb = reduce_manifold(FB, dist, nb);
  for (int i = 0; i < nb; i++) {
    if (dist[i] >= 0) continue;
    double pos[3];
    for (int c = 0; c < 3; c++) pos[c] = FB[i].x * t1[c] + FB[i].y * t2[c] + (FB[i].h - 0.5 * dist[i]) * frame[c];
    emit(m, d, P->geom, B->geom, frame, pos, dist[i]);
  }
}

/* ------------------------------------------------------------------------------------------ broad / mid phase */
static int sphere_vs_obb(const double *c, double r, const shape *box, const double *half) {
  double l[3], t[3];
  sub3(t, c, box->pos); mulmtv(l, box->mat, t);
  double d2 = 0;
  for (int k = 0; k < 3; k++) { double e = fabs(l[k]) - half[k]; if (e > 0) d2 += e * e; }
  return d2 <= r * r;
}
static void bound_half(const shape *s, double *half) {
  switch (s->type) {
    case SO_GEOM_BOX: memcpy(half, s->size, sizeof(double) * 3); break;
    case SO_GEOM_CYLINDER: half[0] = half[1] = s->size[0]; half[2] = s->size[1]; break;
    case SO_GEOM_CAPSULE: half[0] = half[1] = s->size[0]; half[2] = s->size[0] + s->size[1]; break;
    default: half[0] = half[1] = half[2] = s->rbound;
  }
}

void so_collide(const so_model *m, so_data *d) {
  d->ncon = 0;
  for (int p = 0; p < m->npair; p++) {
    int b1 = m->bodypair[2 * p], b2 = m->bodypair[2 * p + 1];
    /* body-level bounding spheres (the world body holds only the floor plane: always descend) */
    if (b1 != 0) {
      double c1[3], c2[3], t[3];
      mulmv(t, d->xmat[b1], m->body_bcenter + 3 * b1); for (int c = 0; c < 3; c++) c1[c] = d->xpos[b1][c] + t[c];
      mulmv(t, d->xmat[b2], m->body_bcenter + 3 * b2); for (int c = 0; c < 3; c++) c2[c] = d->xpos[b2][c] + t[c];
      sub3(t, c1, c2);
      double r = m->body_rbound[b1] + m->body_rbound[b2];
      if (dot3(t, t) > r * r) continue;
    }
    for (int g1 = m->body_geomadr[b1]; g1 < m->body_geomadr[b1] + m->body_geomnum[b1]; g1++) {
      shape A;
      make_shape(m, d, g1, &A);
      for (int g2 = m->body_geomadr[b2]; g2 < m->body_geomadr[b2] + m->body_geomnum[b2]; g2++) {
        shape B;
        make_shape(m, d, g2, &B);
        A.hint_store = 0; /* hill-climbing warm starts are per pair */
        if (A.type == SO_GEOM_PLANE) {
          double n[3] = {A.mat[2], A.mat[5], A.mat[8]};
          if (dot3(n, B.center) - dot3(n, A.pos) - B.rbound > 0) continue;
          collide_plane(m, d, &A, &B);
          continue;
        }
        double t[3];
        sub3(t, A.center, B.center);
        double r = A.rbound + B.rbound;
        if (dot3(t, t) > r * r) continue;
        double half[3];
        if (A.type != SO_GEOM_HULL) { bound_half(&A, half); if (!sphere_vs_obb(B.center, B.rbound, &A, half)) continue; }
        if (B.type != SO_GEOM_HULL) { bound_half(&B, half); if (!sphere_vs_obb(A.center, A.rbound, &B, half)) continue; }
        collide_convex(m, d, &A, &B);
      }
    }
  }
}
